// b200flac_metadata.cu -- metadata finalisation after an encode, host only (no device code).
//
// What FlacAudio.from_pcm does in Python once encode_flac has returned its (byte offset, PCM frames)
// list (audiotools/flac.py:1811-1832): build a SEEKTABLE from the list (flac.py:1847-1876), add it to
// the file's metadata in FlacMetaData.add_block's preferred order (flac.py:53-75), optionally tag the
// VORBIS_COMMENT with WAVEFORMATEXTENSIBLE_CHANNEL_MASK (flac.py:1827-1832), and write the metadata
// back with update_metadata's rule (flac.py:1369-1462): shrink the PADDING blocks so that the frames
// do not move when the growth fits in them, else rewrite the file.
// this removes a Python pass per file from the caller.
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/b200flac.h"

extern "C" void b200flac_internal_set_error(const char* msg);

namespace {

struct Block {
    uint32_t id;
    std::string payload;
};

const uint32_t kPreferredOrder[7] = {0, 3, 5, 4, 6, 2, 1};   // flac.py:59-65

void fail(const char* what, const char* path)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "%s: %s (%s)", what, path, errno ? strerror(errno) : "invalid FLAC file");
    b200flac_internal_set_error(buf);
}

// FlacMetaData.add_block: before the first block that comes later in the preferred order, else last
void add_block(std::vector<Block>& blocks, const Block& b)
{
    size_t pos = 0;
    while (pos < 7 && kPreferredOrder[pos] != b.id) pos++;
    for (size_t i = 0; i < blocks.size(); i++) {
        for (size_t j = pos + 1; j < 7; j++) {
            if (blocks[i].id == kPreferredOrder[j]) {
                blocks.insert(blocks.begin() + (long)i, b);
                return;
            }
        }
    }
    blocks.push_back(b);
}

size_t total_size(const std::vector<Block>& blocks)
{
    size_t n = 0;
    for (const Block& b : blocks) n += 4 + b.payload.size();
    return n;
}

std::string build(const std::vector<Block>& blocks)
{
    std::string out;
    out.reserve(total_size(blocks));
    for (size_t i = 0; i < blocks.size(); i++) {
        const uint32_t len = (uint32_t)blocks[i].payload.size();
        const unsigned char h[4] = {(unsigned char)((i + 1 == blocks.size() ? 0x80 : 0) | blocks[i].id),
                                    (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len};
        out.append((const char*)h, 4);
        out.append(blocks[i].payload);
    }
    return out;
}

void put_be(std::string& s, uint64_t v, int bytes)
{
    for (int i = bytes - 1; i >= 0; i--) s.push_back((char)(v >> (8 * i)));
}

}  // namespace

extern "C" int b200flac_finalize_metadata(const char* filename, const uint64_t* frame_offsets,
                                          const uint32_t* frame_pcm_frames, uint64_t n_frames,
                                          uint32_t seekpoint_interval, uint32_t channel_mask)
{
    errno = 0;
    FILE* f = fopen(filename, "r+b");
    if (!f) { fail("cannot open", filename); return 1; }
    unsigned char magic[4];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "fLaC", 4) != 0) { errno = 0; fail("not a FLAC stream", filename); fclose(f); return 1; }
    std::vector<Block> blocks;
    for (;;) {
        unsigned char h[4];
        if (fread(h, 1, 4, f) != 4) { errno = 0; fail("truncated metadata", filename); fclose(f); return 1; }
        Block b;
        b.id = h[0] & 0x7F;
        const uint32_t len = ((uint32_t)h[1] << 16) | ((uint32_t)h[2] << 8) | h[3];
        b.payload.resize(len);
        if (len && fread(&b.payload[0], 1, len, f) != len) { errno = 0; fail("truncated metadata", filename); fclose(f); return 1; }
        blocks.push_back(b);
        if (h[0] & 0x80) break;
    }
    const size_t old_size = total_size(blocks);
    if (blocks.empty() || blocks[0].id != 0 || blocks[0].payload.size() != 34) {
        errno = 0; fail("STREAMINFO missing", filename); fclose(f); return 1;
    }
    const unsigned char* si = (const unsigned char*)blocks[0].payload.data();
    if (seekpoint_interval == 0) {
        const uint32_t rate = ((uint32_t)si[10] << 12) | ((uint32_t)si[11] << 4) | (si[12] >> 4);
        seekpoint_interval = rate * 10;                       // flac.py:1857-1858
    }
    const uint64_t total_frames = ((uint64_t)(si[13] & 0x0F) << 32) | ((uint64_t)si[14] << 24) | ((uint64_t)si[15] << 16) |
                                  ((uint64_t)si[16] << 8) | si[17];

    // ---- SEEKTABLE, flac.py:1860-1876: one point every interval, at the frame that holds the sample ----
    Block seek;
    seek.id = 3;
    if (seekpoint_interval) {
        uint64_t frame = 0, frame_start = 0;                  // frame index and its first sample
        for (uint64_t pcm = 0; pcm < total_frames; pcm += seekpoint_interval) {
            // bisect_right(sample_offsets, pcm) - 1: the last frame whose first sample is <= pcm
            // (zero-length frames share a first sample; the last of them wins, as in the dict of flac.py:1864)
            while (frame + 1 < n_frames && frame_start + frame_pcm_frames[frame] <= pcm) {
                frame_start += frame_pcm_frames[frame];
                frame++;
            }
            if (n_frames == 0) break;
            put_be(seek.payload, frame_start, 8);
            put_be(seek.payload, frame_offsets[frame], 8);
            put_be(seek.payload, frame_pcm_frames[frame], 2);
        }
    }
    add_block(blocks, seek);

    // ---- WAVEFORMATEXTENSIBLE_CHANNEL_MASK, flac.py:1827-1832 ----
    if (channel_mask) {
        for (Block& b : blocks) {
            if (b.id != 4) continue;
            const unsigned char* p = (const unsigned char*)b.payload.data();
            if (b.payload.size() < 8) break;
            const uint32_t vlen = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
            if (b.payload.size() < 8 + (size_t)vlen) break;
            uint32_t count = p[4 + vlen] | (p[5 + vlen] << 8) | (p[6 + vlen] << 16) | ((uint32_t)p[7 + vlen] << 24);
            char tag[64];
            const int tl = snprintf(tag, sizeof(tag), "WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x%.4X", channel_mask);
            count++;
            std::string np = b.payload.substr(0, 4 + vlen);
            for (int i = 0; i < 4; i++) np.push_back((char)(count >> (8 * i)));
            np.append(b.payload.substr(8 + vlen));
            for (int i = 0; i < 4; i++) np.push_back((char)((uint32_t)tl >> (8 * i)));
            np.append(tag, (size_t)tl);
            b.payload.swap(np);
            break;
        }
    }

    // ---- update_metadata, flac.py:1386-1462 ----
    size_t total_padding = 0;
    bool has_padding = false;
    for (const Block& b : blocks) if (b.id == 1) { has_padding = true; total_padding += b.payload.size(); }
    long long delta = (long long)total_size(blocks) - (long long)old_size;
    if (has_padding && delta <= (long long)total_padding) {
        for (Block& b : blocks) {
            if (b.id != 1) continue;
            if (delta > 0) {
                const size_t take = (size_t)delta <= b.payload.size() ? (size_t)delta : b.payload.size();
                b.payload.resize(b.payload.size() - take);
                delta -= (long long)take;
            } else if (delta < 0) {
                b.payload.resize(b.payload.size() + (size_t)(-delta), '\0');
                delta = 0;
            } else break;
        }
        const std::string img = build(blocks);
        if (fseek(f, 4, SEEK_SET) != 0 || fwrite(img.data(), 1, img.size(), f) != img.size()) {
            fail("cannot rewrite metadata", filename); fclose(f); return 1;
        }
        if (fclose(f) != 0) { fail("cannot rewrite metadata", filename); return 1; }
        return 0;
    }
    // the growth does not fit the padding: the frames have to move (flac.py:1430-1462)
    const std::string img = build(blocks);
    std::string tmp = std::string(filename) + ".tmp";
    FILE* g = fopen(tmp.c_str(), "wb");
    if (!g) { fail("cannot create", tmp.c_str()); fclose(f); return 1; }
    bool ok = fwrite("fLaC", 1, 4, g) == 4 && fwrite(img.data(), 1, img.size(), g) == img.size();
    if (ok && fseek(f, (long)(4 + old_size), SEEK_SET) != 0) ok = false;
    std::vector<char> buf(1 << 20);
    while (ok) {
        const size_t r = fread(buf.data(), 1, buf.size(), f);
        if (r == 0) break;
        if (fwrite(buf.data(), 1, r, g) != r) ok = false;
    }
    fclose(f);
    if (fclose(g) != 0) ok = false;
    if (!ok || rename(tmp.c_str(), filename) != 0) { fail("cannot rewrite", filename); remove(tmp.c_str()); return 1; }
    return 0;
}
