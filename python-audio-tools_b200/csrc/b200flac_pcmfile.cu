// b200flac_pcmfile.cu -- file-backed PCM sources of the C ABI (include/b200flac.h).
//
// Host-only restatement of the two container parsers that stand in front of the FLAC encoder in
// the reference:
//   WaveReader.__init__    audiotools/wav.py:424-502     parse_fmt   wav.py:288-354
//   AiffReader.__init__    audiotools/aiff.py:353-432    parse_comm  aiff.py:327-347
//                                                        parse_ieee_extended aiff.py:25-36
// and the one-call file-to-file forms on top of the stream layer.  Error texts are the reference's
// (audiotools/text.py:530-544, 621-634); the return code says which exception class it raises
// (1 = ValueError, 2 = IOError).
#include <errno.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200flac.h"

extern "C" void b200flac_internal_set_error(const char* msg); // b200flac_encoder.cu

namespace {

struct Rd {
    FILE* f;
    uint64_t pos, size;
    // file.read(n): fewer bytes at the end of the file, never an error
    size_t read(uint8_t* dst, size_t n)
    {
        const size_t got = fread(dst, 1, n, f);
        pos += got;
        return got;
    }
    // file.read(n) whose result is thrown away
    void skip(uint64_t n)
    {
        pos = (n > size - pos) ? size : pos + n;
        fseeko(f, (off_t)pos, SEEK_SET);
    }
};

int fail(int rc, const char* msg)
{
    b200flac_internal_set_error(msg);
    return rc;
}

bool open_rd(const char* path, Rd* r)
{
    r->f = fopen(path, "rb");
    if (!r->f) {
        char msg[400];
        snprintf(msg, sizeof(msg), "[Errno %d] %s: '%.300s'", errno, strerror(errno), path);
        b200flac_internal_set_error(msg);
        return false;
    }
    fseeko(r->f, 0, SEEK_END);
    r->size = (uint64_t)ftello(r->f);
    fseeko(r->f, 0, SEEK_SET);
    r->pos = 0;
    return true;
}

// WaveAudio.PRINTABLE_ASCII / AiffAudio.PRINTABLE_ASCII: wav.py:587, aiff.py:498
bool printable(const uint8_t id[4])
{
    for (int i = 0; i < 4; i++)
        if (id[i] < 0x20 || id[i] > 0x7E) return false;
    return true;
}

uint32_t le32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint32_t le16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
uint32_t be32(const uint8_t* p) { return (uint32_t)p[3] | ((uint32_t)p[2] << 8) | ((uint32_t)p[1] << 16) | ((uint32_t)p[0] << 24); }
uint32_t be16(const uint8_t* p) { return (uint32_t)p[1] | ((uint32_t)p[0] << 8); }

// parse_fmt, wav.py:288-354.  Reads 16 bytes (plain PCM) or 40 (WAVEFORMATEXTENSIBLE) and leaves the
// rest of the chunk where it is -- as the reference does.
int parse_fmt(Rd* r, b200flac_pcm_source* s)
{
    uint8_t b[40];
    if (r->read(b, 16) < 16) return fail(2, "I/O error reading stream");
    const uint32_t compression = le16(b);
    s->channels = le16(b + 2);
    s->sample_rate = le32(b + 4);
    s->bits_per_sample = le16(b + 14);
    if (compression == 1) {
        // a multi-channel WAVE that is not WAVEFORMATEXTENSIBLE is assumed to follow SMPTE/ITU-R order
        static const uint32_t by_count[7] = {0, 0x4, 0x3, 0x7, 0x33, 0x37, 0x3F};
        s->channel_mask = s->channels <= 6 ? by_count[s->channels] : 0;
        return 0;
    }
    if (compression == 0xFFFE) {
        static const uint8_t pcm_guid[16] = {0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x10, 0x00,
                                             0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71};
        if (r->read(b + 16, 24) < 24) return fail(2, "I/O error reading stream");
        s->channel_mask = le32(b + 20);
        if (memcmp(b + 24, pcm_guid, 16) != 0) return fail(1, "invalid WAVE sub-format");
        return 0;
    }
    return fail(1, "unsupported WAVE compression");
}

// parse_ieee_extended, aiff.py:25-36, then int() as parse_comm applies it (aiff.py:339)
int parse_rate(const uint8_t* p, uint32_t* rate)
{
    const uint32_t se = be16(p);
    const uint64_t mant = ((uint64_t)be32(p + 2) << 32) | be32(p + 6);
    const uint32_t exponent = se & 0x7FFF;
    double v;
    if (exponent == 0 && mant == 0) v = 0.0;
    else if (exponent == 0x7FFF) v = 1.79769313486231e+308;
    else {
        v = ldexp((double)mant, (int)exponent - 16383 - 63);
        if (se & 0x8000) v = -v;
    }
    v = trunc(v);
    if (!(v >= 0.0 && v <= 4294967295.0)) return fail(1, "sample rate out of range");
    *rate = (uint32_t)v;
    return 0;
}

int encode_container(int aiff, const char* flac_filename, const char* in_filename, const b200flac_params* params,
                     uint32_t padding_size, const char* version, const int* devices, int n_devices,
                     b200flac_pcm_source* src_out, uint64_t** frame_offsets, uint32_t** frame_pcm_frames,
                     uint64_t* n_frames)
{
    if (!flac_filename || !in_filename || !params) return fail(3, "filename/params is NULL");
    b200flac_pcm_source src;
    const int rc = aiff ? b200flac_aiff_probe(in_filename, &src) : b200flac_wave_probe(in_filename, &src);
    if (rc) return rc;
    if (src_out) *src_out = src;
    b200flac_params p = *params;
    p.sample_rate = src.sample_rate;
    p.channels = src.channels;
    p.bits_per_sample = src.bits_per_sample;
    b200flac_stream* s = b200flac_stream_open(flac_filename, &p, padding_size, version, devices, n_devices);
    if (!s) return 3;
    const int wrc = b200flac_stream_write_file(s, in_filename, src.data_offset, src.total_pcm_frames, src.flags);
    if (wrc) {
        b200flac_stream_close(s, 1, nullptr, nullptr, nullptr);
        if (wrc == 2) return fail(2, aiff ? "premature end of SSND chunk" : "premature end of data chunk");
        return 3;
    }
    return b200flac_stream_close(s, 0, frame_offsets, frame_pcm_frames, n_frames) ? 3 : 0;
}

} // namespace

extern "C" int b200flac_wave_probe(const char* path, b200flac_pcm_source* src)
{
    if (!path || !src) return fail(1, "path/src is NULL");
    memset(src, 0, sizeof(*src));
    Rd r;
    if (!open_rd(path, &r)) return 2;
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{r.f};
    uint8_t h[12];
    if (r.read(h, 12) < 12) return fail(1, "invalid RIFF WAVE file");
    if (memcmp(h, "RIFF", 4) != 0) return fail(1, "not a RIFF WAVE file");
    if (memcmp(h + 8, "WAVE", 4) != 0) return fail(1, "invalid RIFF WAVE file");
    int64_t total = (int64_t)le32(h + 4) - 4;
    bool fmt_read = false;
    uint32_t frame_bytes = 0;
    while (total > 0) {
        if (r.read(h, 8) < 8) return fail(1, "invalid RIFF WAVE file");
        if (!printable(h)) return fail(1, "invalid RIFF WAVE chunk ID");
        total -= 8;
        const uint32_t chunk_size = le32(h + 4);
        if (memcmp(h, "fmt ", 4) == 0) {
            const int rc = parse_fmt(&r, src);
            if (rc) return rc;
            frame_bytes = (src->bits_per_sample / 8) * src->channels;
            fmt_read = true;
        } else if (memcmp(h, "data", 4) == 0) {
            if (!fmt_read) return fail(1, "data chunk found before fmt");
            if (frame_bytes == 0) return fail(1, "integer division or modulo by zero");
            src->total_pcm_frames = chunk_size / frame_bytes;
            src->data_offset = r.pos;
            src->flags = src->bits_per_sample == 8 ? B200FLAC_PCM_UNSIGNED : 0u; // wav.py:527
            return 0;
        } else {
            r.skip(chunk_size);
        }
        if (chunk_size % 2) {
            uint8_t pad;
            if (r.read(&pad, 1) < 1) return fail(1, "invalid RIFF WAVE chunk ID");
            total -= (int64_t)chunk_size + 1;
        } else {
            total -= chunk_size;
        }
    }
    return fail(1, "data chunk not found");
}

extern "C" int b200flac_aiff_probe(const char* path, b200flac_pcm_source* src)
{
    if (!path || !src) return fail(1, "path/src is NULL");
    memset(src, 0, sizeof(*src));
    Rd r;
    if (!open_rd(path, &r)) return 2;
    struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{r.f};
    uint8_t h[18];
    if (r.read(h, 12) < 12) return fail(1, "invalid AIFF file");
    if (memcmp(h, "FORM", 4) != 0) return fail(1, "not an AIFF file");
    if (memcmp(h + 8, "AIFF", 4) != 0) return fail(1, "invalid AIFF file");
    int64_t total = (int64_t)be32(h + 4) - 4;
    bool comm_read = false;
    while (total > 0) {
        if (r.read(h, 8) < 8) return fail(1, "invalid AIFF file");
        if (!printable(h)) return fail(1, "invalid AIFF chunk ID");
        total -= 8;
        const uint32_t chunk_size = be32(h + 4);
        if (memcmp(h, "COMM", 4) == 0) {
            if (r.read(h, 18) < 18) return fail(2, "I/O error reading stream");
            src->channels = be16(h);
            src->total_pcm_frames = be32(h + 2);
            src->bits_per_sample = be16(h + 6);
            const int rc = parse_rate(h + 8, &src->sample_rate);
            if (rc) return rc;
            if (src->channels == 2) src->channel_mask = 0x3;        // ChannelMask.from_channels,
            else if (src->channels == 1) src->channel_mask = 0x4;   // audiotools/__init__.py:2049-2060
            else if (src->channels == 0) return fail(1, "ambiguous channel assignment");
            else src->channel_mask = 0;
            comm_read = true;
        } else if (memcmp(h, "SSND", 4) == 0) {
            if (!comm_read) return fail(1, "SSND chunk found before fmt");
            r.skip(8); // "offset" and "block_size", aiff.py:415
            src->data_offset = r.pos;
            src->flags = B200FLAC_PCM_BIG_ENDIAN;
            return 0;
        } else {
            r.skip(chunk_size);
        }
        if (chunk_size % 2) {
            uint8_t pad;
            if (r.read(&pad, 1) < 1) return fail(1, "invalid AIFF chunk");
            total -= (int64_t)chunk_size + 1;
        } else {
            total -= chunk_size;
        }
    }
    return fail(1, "SSND chunk not found");
}

extern "C" int b200flac_encode_wave(const char* flac_filename, const char* wave_filename, const b200flac_params* params,
                                    uint32_t padding_size, const char* version, const int* devices, int n_devices,
                                    b200flac_pcm_source* src, uint64_t** frame_offsets, uint32_t** frame_pcm_frames,
                                    uint64_t* n_frames)
{
    return encode_container(0, flac_filename, wave_filename, params, padding_size, version, devices, n_devices, src,
                            frame_offsets, frame_pcm_frames, n_frames);
}

extern "C" int b200flac_encode_aiff(const char* flac_filename, const char* aiff_filename, const b200flac_params* params,
                                    uint32_t padding_size, const char* version, const int* devices, int n_devices,
                                    b200flac_pcm_source* src, uint64_t** frame_offsets, uint32_t** frame_pcm_frames,
                                    uint64_t* n_frames)
{
    return encode_container(1, flac_filename, aiff_filename, params, padding_size, version, devices, n_devices, src,
                            frame_offsets, frame_pcm_frames, n_frames);
}
