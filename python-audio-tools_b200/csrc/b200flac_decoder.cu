// b200flac_decoder.cu -- frame-parallel FLAC decode / verify on the GPU.
//
// Reference being restated (src/decoders/flac.c of widgital/python-audio-tools):
//   flacdec_read_metadata          :569-708   STREAMINFO (host, here: parse_streaminfo)
//   flacdec_read_frame_header      :710-852   + read_utf8 :1313-1323     (dec_parse_header, host and device)
//   flacdec_read_subframe(_header) :854-950, subframe bits per sample :952-965
//   flacdec_read_constant/verbatim/fixed/lpc_subframe :967-1133
//   flacdec_read_residual          :1135-1210
//   flacdec_decorrelate_channels   :1213-1270
//   frame loop, CRC-16, MD5        :174-286 (FlacDecoder_read), :1395-1501 (standalone flacdec)
//
// The reference walks the stream serially because a frame's length is only known once its last
// residual has been read.  Here every byte position that holds a header which is valid for this
// stream (sync code, field codes, CRC-8, agreement with STREAMINFO) is a *candidate*; one thread
// decodes each candidate to the end of its frame, speculatively and all at once, and reports where
// the frame ended and whether its CRC-16 holds.  The host then follows the chain first frame ->
// end -> next frame -> ... exactly as the reference's loop does (a position without a valid
// candidate is diagnosed by parsing that header on the host, which yields the reference's error),
// and a last kernel undoes the channel decorrelation of the frames on the chain and writes the
// interleaved PCM at their final positions.  False candidates (a header look-alike inside
// compressed data: about one per 16 MB) cost one wasted thread each.
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "../../include/b200flac.h"

extern "C" void b200flac_internal_set_error(const char* msg);                         // b200flac_encoder.cu
extern "C" void b200flac_internal_md5(const uint8_t* p, size_t n, uint8_t out[16]);   // b200flac_stream.cu

typedef unsigned char u8;
typedef unsigned int u32;
typedef unsigned long long u64;

// flac_status, src/decoders/flac.h:68-81, plus the two conditions the reference reports through
// other channels (EOF via br_abort -> "EOF reading frame"; the frame CRC-16)
enum {
    DS_OK = 0, DS_ERROR, DS_INVALID_SYNC_CODE, DS_INVALID_RESERVED_BIT, DS_INVALID_BITS_PER_SAMPLE,
    DS_INVALID_SAMPLE_RATE, DS_INVALID_FRAME_CRC, DS_SAMPLE_RATE_MISMATCH, DS_CHANNEL_COUNT_MISMATCH,
    DS_BITS_PER_SAMPLE_MISMATCH, DS_MAXIMUM_BLOCK_SIZE_EXCEEDED, DS_INVALID_CODING_METHOD,
    DS_INVALID_FIXED_ORDER, DS_INVALID_SUBFRAME_TYPE,
    DS_EOF = 32, DS_FRAME_CRC16 = 33, DS_MALFORMED = 34, DS_UNDEFINED = 35
};

static const char* ds_strerror(u32 s)
{
    switch (s) { // FlacDecoder_strerror, src/decoders/flac.c:1273-1311
    case DS_OK: return "No Error";
    case DS_INVALID_SYNC_CODE: return "invalid sync code";
    case DS_INVALID_RESERVED_BIT: return "invalid reserved bit";
    case DS_INVALID_BITS_PER_SAMPLE: return "invalid bits per sample";
    case DS_INVALID_SAMPLE_RATE: return "invalid sample rate";
    case DS_INVALID_FRAME_CRC: return "invalid checksum in frame header";
    case DS_SAMPLE_RATE_MISMATCH: return "frame sample rate does not match STREAMINFO sample rate";
    case DS_CHANNEL_COUNT_MISMATCH: return "frame channel count does not match STREAMINFO channel count";
    case DS_BITS_PER_SAMPLE_MISMATCH: return "frame bits-per-sample does not match STREAMINFO bits per sample";
    case DS_MAXIMUM_BLOCK_SIZE_EXCEEDED: return "frame block size exceeds STREAMINFO's maximum block size";
    case DS_INVALID_CODING_METHOD: return "invalid residual partition coding method";
    case DS_INVALID_FIXED_ORDER: return "invalid FIXED subframe order";
    case DS_INVALID_SUBFRAME_TYPE: return "invalid subframe type";
    case DS_EOF: return "EOF reading frame";                      // flac.c:262, :1470
    case DS_FRAME_CRC16: return "invalid checksum in frame";      // flac.c:252, :1459
    // a residual block whose partitions do not add up to the block size: the reference reads what the
    // partition arithmetic says (flac.c:1157-1163), indexes past the residuals it got, and then fails the
    // frame's CRC-16 -- that is the error its caller sees
    case DS_MALFORMED: return "invalid checksum in frame";
    // more wasted bits than bits per sample: the reference's read count underflows (flac.c:872-873); no defined
    // behaviour to follow, reported like any other damaged frame
    case DS_UNDEFINED: return "invalid checksum in frame";
    default: return "Error";
    }
}

struct DecStream {
    u32 sample_rate, channels, bits_per_sample, max_block_size;
};

struct DecHeader {
    u32 block_size, assignment, length; // length: header bytes including the CRC-8
};

struct bf_dec_cand {
    u64 pos;          // byte offset of the sync code inside the frame region
    u64 end;          // byte offset just past the frame's CRC-16
    u32 block_size;
    u32 status;       // DS_*
    u32 assignment;
    u8 wasted[B200FLAC_MAX_CHANNELS];
    u32 pad;
};

struct bf_dec_emit {  // one frame on the chain
    u64 pcm_frame;    // index of its first PCM frame in the output
    u32 cand;
    u32 n;            // PCM frames to write
};

// ---- frame header, host and device (flacdec_read_frame_header, flac.c:710-852) -------------------
__host__ __device__ inline u32 dec_crc8(const u8* p, u32 n)
{
    u32 crc = 0;
    for (u32 i = 0; i < n; i++) {
        crc ^= p[i];
        for (int k = 0; k < 8; k++) crc = (crc & 0x80) ? ((crc << 1) ^ 0x07) & 0xFF : (crc << 1) & 0xFF;
    }
    return crc;
}

__host__ __device__ inline u32 dec_parse_header(const u8* b, u64 avail, const DecStream& S, DecHeader* h)
{
    if (avail < 2) return DS_EOF;
    if (b[0] != 0xFF || (b[1] & 0xFC) != 0xF8) return DS_INVALID_SYNC_CODE;
    if (b[1] & 0x02) return DS_INVALID_RESERVED_BIT;
    if (avail < 5) return DS_EOF;
    const u32 bs_bits = b[2] >> 4, sr_bits = b[2] & 15;
    h->assignment = b[3] >> 4;
    const u32 channel_count = (h->assignment >= 8 && h->assignment <= 10) ? 2 : h->assignment + 1;
    u32 bps;
    switch ((b[3] >> 1) & 7) {
    case 0: bps = S.bits_per_sample; break;
    case 1: bps = 8; break;
    case 2: bps = 12; break;
    case 4: bps = 16; break;
    case 5: bps = 20; break;
    case 6: bps = 24; break;
    default: return DS_INVALID_BITS_PER_SAMPLE;
    }
    // read_utf8 (flac.c:1313-1323): leading one bits = byte count; only the length matters here
    u32 ones = 0;
    while (ones < 8 && (b[4] & (0x80u >> ones))) ones++;
    if (ones == 8) return DS_INVALID_FRAME_CRC; // the reference's read(7 - 8) is undefined; such a header never checks
    u32 n = 5 + (ones > 1 ? ones - 1 : 0);
    if (avail < n + 5) return DS_EOF;
    switch (bs_bits) {
    case 0: h->block_size = S.max_block_size; break;
    case 1: h->block_size = 192; break;
    case 2: case 3: case 4: case 5: h->block_size = 576u << (bs_bits - 2); break;
    case 6: h->block_size = (u32)b[n] + 1; n += 1; break;
    case 7: h->block_size = (((u32)b[n] << 8) | b[n + 1]) + 1; n += 2; break;
    default: h->block_size = 256u << (bs_bits - 8); break;
    }
    u32 rate;
    switch (sr_bits) {
    case 0: rate = S.sample_rate; break;
    case 1: rate = 88200; break;
    case 2: rate = 176400; break;
    case 3: rate = 192000; break;
    case 4: rate = 8000; break;
    case 5: rate = 16000; break;
    case 6: rate = 22050; break;
    case 7: rate = 24000; break;
    case 8: rate = 32000; break;
    case 9: rate = 44100; break;
    case 10: rate = 48000; break;
    case 11: rate = 96000; break;
    case 12: rate = (u32)b[n] * 1000; n += 1; break;
    case 13: rate = ((u32)b[n] << 8) | b[n + 1]; n += 2; break;
    case 14: rate = (((u32)b[n] << 8) | b[n + 1]) * 10; n += 2; break;
    default: return DS_INVALID_SAMPLE_RATE;
    }
    n += 1; // the CRC-8 byte
    if (dec_crc8(b, n) != 0) return DS_INVALID_FRAME_CRC;
    h->length = n;
    if (S.sample_rate != rate) return DS_SAMPLE_RATE_MISMATCH;
    if (S.channels != channel_count) return DS_CHANNEL_COUNT_MISMATCH;
    if (S.bits_per_sample != bps) return DS_BITS_PER_SAMPLE_MISMATCH;
    if (h->block_size > S.max_block_size) return DS_MAXIMUM_BLOCK_SIZE_EXCEEDED;
    return DS_OK;
}

// ---- kernel 1: candidate scan ---------------------------------------------------------------------
__global__ void k_dec_scan(const u8* __restrict__ data, u64 n_bytes, DecStream S, bf_dec_cand* __restrict__ cands,
                           u32 cap, u32* __restrict__ count)
{
    // four byte positions per thread from two aligned 32-bit loads (the buffer is padded past n_bytes)
    const u32* __restrict__ words = (const u32*)data;
    const u64 n_words = (n_bytes + 3) / 4, stride = (u64)gridDim.x * blockDim.x;
    for (u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x; t < n_words; t += stride) {
        const u32 a = words[t];
        // no 0xFF byte in this word: no sync code starts here (zero-byte test on the complement)
        if ((((~a) - 0x01010101u) & a & 0x80808080u) == 0) continue;
        const u64 v = (u64)a | ((u64)words[t + 1] << 32);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (((v >> (8 * j)) & 0xFF) != 0xFF || ((v >> (8 * j + 8)) & 0xFE) != 0xF8) continue;
            const u64 p = t * 4 + j;
            if (p + 1 >= n_bytes) continue;
            DecHeader h;
            if (dec_parse_header(data + p, n_bytes - p, S, &h) != DS_OK) continue;
            const u32 idx = atomicAdd(count, 1u);
            if (idx < cap) {
                cands[idx].pos = p;
                cands[idx].status = DS_ERROR;
            }
        }
    }
}

// ---- kernel 2: one thread decodes one candidate frame ------------------------------------------------
// MSB-first bit reader over global memory (the reference's BitstreamReader, big-endian): 64-bit window,
// refilled with aligned 32-bit loads.  The buffer is padded with zeros, so reads past the end are safe;
// they raise `eof`, which the caller reports as the reference's "EOF reading frame".
struct DecBits {
    const u8* base;
    u64 n_bytes, next, buf; // next: byte offset (a multiple of 4) of `ahead`
    u32 ahead;              // the word at `next`, as loaded (little-endian), fetched one refill EARLY: the lanes of a warp
                            // walk 32 different frames, so a refill load usually misses L1 for some lane;
                            // issued a refill ahead (about three samples), it has landed when it is needed
    int avail;
    bool eof;

    // (the word is kept as loaded and byte-swapped only when it is consumed: nothing may read the
    // load's register before the next refill, or the warp waits for the load right here)
    __device__ __forceinline__ u32 word(u64 off) const
    {
        u32 w = 0;
        if (off < n_bytes + 32) w = *(const u32*)(base + off);
        return w;
    }
    // keeps at least 32 valid bits in the window.  Every consumer takes at most `avail` bits, so one word
    // always restores the invariant -- a plain `if`, not a loop (as a loop the compiler unrolled it into
    // some fifty instructions of trip-count arithmetic per sample)
    __device__ __forceinline__ void refill()
    {
        if (avail < 32) {
            if (next >= n_bytes + 16) eof = true;
            buf |= (u64)__byte_perm(ahead, 0, 0x0123) << (32 - avail);
            avail += 32;
            next += 4;
            ahead = word(next);
            // each lane streams through its own frame once: at every 128-byte line it asks for the line two
            // ahead, so that the word fetched above comes from cache rather than from DRAM
            if ((next & 127) == 0 && next + 512 < n_bytes) asm volatile("prefetch.global.L1 [%0];" ::"l"(base + next + 256));
        }
    }
    __device__ __forceinline__ void init(const u8* b, u64 n, u64 pos)
    {
        base = b; n_bytes = n; eof = false;
        // the first word holds pos & 3 bytes that precede the position
        const u32 skip = (u32)(pos & 3);
        buf = ((u64)__byte_perm(word(pos & ~3ull), 0, 0x0123) << 32) << (8 * skip);
        avail = 32 - 8 * (int)skip;
        next = (pos & ~3ull) + 4;
        ahead = word(next);
        refill();
    }
    __device__ __forceinline__ u32 read(u32 n) // 0..32 bits
    {
        if (n == 0) return 0;
        const u32 v = (u32)(buf >> (64 - n));
        buf <<= n;
        avail -= (int)n;
        refill();
        return v;
    }
    __device__ __forceinline__ int read_signed(u32 n) // two's complement, 1..32 bits (bitstream.c read_signed)
    {
        if (n == 0) return 0;
        const u32 v = read(n);
        return (int)(v << (32 - n)) >> (32 - n);
    }
    __device__ __forceinline__ u32 unary1() // zeros before the next one bit (read_unary(bs, 1))
    {
        u32 cnt = 0;
        for (;;) {
            if (buf) {
                const int z = __clzll((long long)buf);
                cnt += (u32)z;
                buf = (buf << z) << 1;
                avail -= z + 1;
                refill();
                return cnt;
            }
            cnt += (u32)avail;
            avail = 0;
            if (eof) return cnt;
            refill();
        }
    }
    // one Rice code, (unary << k) | k low bits (flac.c:1190-1193).  The window always holds at least 32
    // valid bits here, and nearly every code fits in them: then the zeros, the stop bit and the low bits
    // come out of ONE 32-bit word with one consume and one refill; longer codes take the two general reads
    __device__ __forceinline__ u32 rice(u32 k)
    {
        const u32 top = (u32)(buf >> 32);
        const u32 z = (u32)__clz((int)top);
        const u32 len = z + 1 + k;
        if (len <= 32) {
            const u32 low = k ? (((top << z) << 1) >> (32 - k)) : 0u;
            buf <<= len;
            avail -= (int)len;
            refill();
            return (z << k) | low;
        }
        const u32 msb = unary1();
        return (msb << k) | read(k);
    }
    __device__ __forceinline__ u64 bits_consumed() const { return next * 8 - (u64)avail; }
};

#define DEC_FAST_ORDER 12

// One subframe into row[0..n): flacdec_read_subframe (flac.c:854-916) with the residual decode
// (flac.c:1135-1210) and the predictor (flac.c:1025-1062, 1118-1130) fused sample by sample; the
// wasted bits are re-inserted by the emit kernel.  FIXED orders are the LPC recurrence with binomial
// coefficients and shift 0 (the same value modulo 2^32 as the reference's int expressions).
__device__ u32 dec_subframe(DecBits& rd, u32 n, u32 bps, int* __restrict__ row, u32* wasted_out)
{
    rd.read(1);
    const u32 type = rd.read(6);
    u32 wasted = 0;
    if (rd.read(1)) wasted = rd.unary1() + 1;
    *wasted_out = wasted;
    if (wasted >= bps) return DS_UNDEFINED;
    bps -= wasted;
    if (rd.eof) return DS_EOF;

    if (type == 0) { // CONSTANT
        const int v = rd.read_signed(bps);
        for (u32 i = 0; i < n; i++) row[i] = v;
        return DS_OK;
    }
    if (type == 1) { // VERBATIM
        for (u32 i = 0; i < n; i++) row[i] = rd.read_signed(bps);
        return rd.eof ? DS_EOF : DS_OK;
    }
    u32 order, shift = 0;
    int q[B200FLAC_MAX_LPC_ORDER];
    const bool fixed = (type & 0x38) == 0x08;
    if (fixed) {
        // orders 5..7 are only rejected after the warm-up samples and the residual have been read
        // (the switch of flac.c:1030-1062), so an error inside the residual comes first
        order = type & 7;
    } else if (type & 0x20) {
        order = (type & 0x1F) + 1;
    } else {
        return DS_INVALID_SUBFRAME_TYPE;
    }
    for (u32 i = 0; i < order; i++) { // warm-up samples (read even when the block is shorter than the order)
        const int v = rd.read_signed(bps);
        if (i < n) row[i] = v;
    }
#pragma unroll
    for (int j = 0; j < B200FLAC_MAX_LPC_ORDER; j++) q[j] = 0;
    if (fixed) {
        const int c1[8] = {0, 1, 2, 3, 4, 0, 0, 0}, c2[8] = {0, 0, -1, -3, -6, 0, 0, 0}, c3[8] = {0, 0, 0, 1, 4, 0, 0, 0},
                  c4[8] = {0, 0, 0, 0, -1, 0, 0, 0};
        q[0] = c1[order]; q[1] = c2[order]; q[2] = c3[order]; q[3] = c4[order];
    } else {
        const u32 precision = rd.read(4) + 1;
        const int sh = rd.read_signed(5);
        shift = sh > 0 ? (u32)sh : 0u; // MAX(qlp_shift_needed, 0), flac.c:1102
        for (u32 j = 0; j < order; j++) q[j] = rd.read_signed(precision);
    }

    // residual header, flac.c:1140-1143
    const u32 method = rd.read(2);
    if (method > 1) return DS_INVALID_CODING_METHOD;
    const u32 po = rd.read(4);
    const u32 plen = n >> po;
    if (rd.eof) return DS_EOF;
    const u32 kbits = method ? 5u : 4u, kesc = method ? 31u : 15u, n_parts = 1u << po;
    if (((u64)plen << po) != n || plen < order) {
        // The partitions do not add up to n - order residuals.  The reference reads what its partition
        // arithmetic says (flac.c:1150-1206: max(plen - order, 0) in the first partition, plen in the others),
        // then builds samples from residuals it never read; what its caller sees is decided by the reads
        // alone -- EOF, a later subframe's error, or the frame's CRC-16 -- so the same bits are consumed here,
        // the values dropped, and the frame is marked for the CRC verdict.
        for (u32 part = 0; part < n_parts; part++) {
            u32 count = part == 0 ? (plen > order ? plen - order : 0u) : plen;
            const u32 k = rd.read(kbits);
            const u32 escape = (k == kesc) ? rd.read(5) : 0u;
            for (; count && !rd.eof; count--) {
                if (!escape) { rd.unary1(); rd.read(k); }
                else rd.read(escape);
            }
            if (rd.eof) return DS_EOF;
        }
        if (fixed && order > 4) return DS_INVALID_FIXED_ORDER;
        return DS_MALFORMED;
    }

    // recent samples, most recent first, for orders up to DEC_FAST_ORDER; longer predictors read the row
    int h[DEC_FAST_ORDER], qf[DEC_FAST_ORDER];
#pragma unroll
    for (int j = 0; j < DEC_FAST_ORDER; j++) {
        qf[j] = q[j];
        h[j] = ((u32)j < order) ? row[order - 1 - j] : 0;
    }
    // ONE loop over the samples of the subframe, whatever the partition order: the lanes of a warp decode
    // different frames, and nested partition/sample loops would leave lanes with different partition
    // orders waiting for each other at every partition boundary.  The partition header (flac.c:1157-1186)
    // is a short predicated detour inside the sample loop instead.
    // The loop exists twice, chosen once per subframe: predictors of up to DEC_FAST_ORDER taps run from the
    // register history, longer ones from the row.  After the end of the data the reader hands out zeros and
    // every loop here is bounded by the block size, so EOF is only looked at when the subframe is done.
    u32 part = 0, left = 0, k = 0, escape = 0;
    auto samples = [&](auto fast_tag) {
        constexpr bool FAST = decltype(fast_tag)::value;
#pragma unroll 4
        for (u32 i = order; i < n; i++) {
            while (left == 0) {
                left = part == 0 ? plen - order : plen;
                part++;
                k = rd.read(kbits);
                escape = (k == kesc) ? rd.read(5) : 0u;
            }
            left--;
            int r;
            if (!escape) {
                const u32 v = rd.rice(k);
                r = (int)(v >> 1) ^ -(int)(v & 1);
            } else {
                r = rd.read_signed(escape);
            }
            long long acc = 0;
            if (FAST) {
#pragma unroll
                for (int j = 0; j < DEC_FAST_ORDER; j++) acc += (long long)qf[j] * (long long)h[j];
            } else {
                for (u32 j = 0; j < order; j++) acc += (long long)q[j] * (long long)row[i - 1 - j];
            }
            const int s = (int)(acc >> shift) + r;
            row[i] = s;
            if (FAST) {
#pragma unroll
                for (int j = DEC_FAST_ORDER - 1; j > 0; j--) h[j] = h[j - 1];
                h[0] = s;
            }
        }
    };
    // (measured: a narrower copy of the loop for FIXED and short predictors doubles the kernel's time -- the lanes
    // of a warp then sit in different copies and run one after the other; one body for every order up to 12 it is)
    if (order <= DEC_FAST_ORDER) samples(std::true_type());
    else samples(std::false_type());
    for (; part < n_parts; part++) { // headers of partitions without residuals (order == partition length)
        if (rd.read(kbits) == kesc) rd.read(5);
    }
    if (rd.eof) return DS_EOF;
    if (fixed && order > 4) return DS_INVALID_FIXED_ORDER;
    return DS_OK;
}

// CRC-16 of src/common/flac_crc.c:62-100 over [pos, pos + n): the aligned middle sixteen bytes per step --
// the next 16-byte load is issued before the current one is folded in, four bytes per table step
// (tab[k][b] = CRC of byte b followed by k zero bytes) -- head and tail bytes singly
__device__ __forceinline__ u32 dec_crc16_word(u32 crc, u32 w, const unsigned short (*tab)[256])
{
    return tab[3][((crc >> 8) ^ w) & 0xFF] ^ tab[2][(crc ^ (w >> 8)) & 0xFF] ^ tab[1][(w >> 16) & 0xFF] ^ tab[0][w >> 24];
}

__device__ u32 dec_crc16(const u8* __restrict__ data, u64 pos, u64 n, const unsigned short (*tab)[256])
{
    u32 crc = 0;
    const u8* p = data + pos;
    for (; n && ((size_t)p & 15); p++, n--) crc = ((crc << 8) ^ tab[0][((crc >> 8) ^ *p) & 0xFF]) & 0xFFFF;
    if (n >= 16) {
        uint4 cur = *(const uint4*)p;
        for (;;) {
            n -= 16;
            p += 16;
            const bool more = n >= 16;
            uint4 nx = cur;
            if (more) nx = *(const uint4*)p;
            crc = dec_crc16_word(crc, cur.x, tab);
            crc = dec_crc16_word(crc, cur.y, tab);
            crc = dec_crc16_word(crc, cur.z, tab);
            crc = dec_crc16_word(crc, cur.w, tab);
            if (!more) break;
            cur = nx;
        }
    }
    for (; n; p++, n--) crc = ((crc << 8) ^ tab[0][((crc >> 8) ^ *p) & 0xFF]) & 0xFFFF;
    return crc;
}

__global__ void __launch_bounds__(32) k_dec_frames(const u8* __restrict__ data, u64 n_bytes, DecStream S,
                                                   bf_dec_cand* __restrict__ cands, u32 n_cands,
                                                   int* __restrict__ scratch, u32 row_stride)
{
    __shared__ unsigned short tab[4][256];
    for (u32 b = threadIdx.x; b < 256; b += blockDim.x) {
        u32 crc = b << 8;
        for (int k = 0; k < 8; k++) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x8005) & 0xFFFF : (crc << 1) & 0xFFFF;
        tab[0][b] = (unsigned short)crc;
    }
    __syncthreads();
    for (int k = 1; k < 4; k++) {
        for (u32 b = threadIdx.x; b < 256; b += blockDim.x) {
            const u32 c = tab[k - 1][b];
            tab[k][b] = (unsigned short)(((c << 8) ^ tab[0][c >> 8]) & 0xFFFF);
        }
        __syncthreads();
    }
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cands) return;
    bf_dec_cand cd = cands[c];
    DecHeader h;
    u32 status = dec_parse_header(data + cd.pos, n_bytes - cd.pos, S, &h);
    cd.block_size = h.block_size;
    cd.assignment = h.assignment;
    DecBits rd;
    if (status == DS_OK) {
        rd.init(data, n_bytes, cd.pos + h.length);
        bool malformed = false; // a subframe whose residual partitions do not add up: verdict by the CRC-16 below
        for (u32 ch = 0; ch < S.channels && status == DS_OK; ch++) {
            // flacdec_subframe_bits_per_sample, flac.c:952-965
            const bool side = (h.assignment == 8 && ch == 1) || (h.assignment == 9 && ch == 0) || (h.assignment == 10 && ch == 1);
            u32 w = 0;
            status = dec_subframe(rd, h.block_size, S.bits_per_sample + (side ? 1u : 0u),
                                  scratch + ((u64)c * S.channels + ch) * row_stride, &w);
            cd.wasted[ch] = (u8)w;
            if (status == DS_MALFORMED) { malformed = true; status = DS_OK; }
        }
        if (status == DS_OK && malformed) status = DS_MALFORMED + 1000; // resolved after the CRC-16
    }
    const bool soft = status == DS_MALFORMED + 1000;
    if (soft) status = DS_OK;
    if (status == DS_OK) {
        if (rd.eof) status = DS_EOF;
        const u64 end = (rd.bits_consumed() + 7) / 8 + 2; // byte_align, then the CRC-16 (flac.c:247-249)
        if (status == DS_OK && end > n_bytes) status = DS_EOF;
        if (status == DS_OK) {
            cd.end = end;
            if (dec_crc16(data, cd.pos, end - cd.pos, tab) != 0) status = DS_FRAME_CRC16;
            else if (soft) status = DS_MALFORMED;
        }
    }
    cd.status = status;
    cands[c] = cd;
}

// ---- kernel 3: decorrelate, re-insert wasted bits, interleave, pack (flac.c:1213-1270, 905-913) ------
__global__ void k_dec_emit(const bf_dec_cand* __restrict__ cands, const bf_dec_emit* __restrict__ emits,
                           const int* __restrict__ scratch, u32 row_stride, DecStream S, u8* __restrict__ out)
{
    const bf_dec_emit e = emits[blockIdx.x];
    const bf_dec_cand cd = cands[e.cand];
    const u32 C = S.channels, B = S.bits_per_sample / 8;
    const int* rows = scratch + (u64)e.cand * C * row_stride;
    u8* dst = out + e.pcm_frame * C * B;
    if (C == 2) {
        const u32 w0 = cd.wasted[0], w1 = cd.wasted[1];
        for (u32 i = threadIdx.x; i < e.n; i += blockDim.x) {
            const int a = (int)((u32)rows[i] << w0), b = (int)((u32)rows[row_stride + i] << w1);
            int l, r;
            if (cd.assignment == 8) { l = a; r = a - b; }
            else if (cd.assignment == 9) { l = a + b; r = b; }
            else if (cd.assignment == 10) {
                const long long mid = ((long long)a << 1) | (b & 1);
                l = (int)((mid + b) >> 1);
                r = (int)((mid - b) >> 1);
            } else { l = a; r = b; }
            if (B == 2) {
                ((u32*)dst)[i] = ((u32)l & 0xFFFF) | ((u32)r << 16);
            } else {
                u8* o = dst + (u64)i * 2 * B;
                for (u32 k = 0; k < B; k++) { o[k] = (u8)((u32)l >> (8 * k)); o[B + k] = (u8)((u32)r >> (8 * k)); }
            }
        }
        return;
    }
    for (u32 t = threadIdx.x; t < e.n * C; t += blockDim.x) {
        const u32 i = t / C, ch = t - i * C;
        const u32 v = (u32)rows[(u64)ch * row_stride + i] << cd.wasted[ch];
        u8* o = dst + (u64)t * B;
        for (u32 k = 0; k < B; k++) o[k] = (u8)(v >> (8 * k));
    }
}

// ---- the frame chain on the device ----------------------------------------------------------------------
// first frame -> its end -> the frame that starts there -> ...: a linked list through the candidates.
// Positions are looked up through a hash table, every candidate learns its predecessor, and pointer
// jumping (log2 rounds) gives each candidate the root of its list, the PCM frames and the FLAC frames in
// front of it.  The candidates whose root is the candidate at offset 0 are the stream.  Anything unusual
// -- a damaged or missing frame on the way, two candidates ending at the same place, a chain that does
// not add up to STREAMINFO's total -- is left to the host walk below, which also words the error.
#define DC_NONE 0xFFFFFFFFu
struct bf_dec_chain_result { u32 bad, n_emit; u64 covered; };

__device__ __forceinline__ u32 dc_slot(u64 pos, u32 bits) { return (u32)((pos * 0x9E3779B97F4A7C15ull) >> (64 - bits)); }

__global__ void k_chain_insert(const bf_dec_cand* __restrict__ cands, u32 count, u32* __restrict__ table, u32 bits)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const u32 mask = (1u << bits) - 1;
    for (u32 h = dc_slot(cands[i].pos, bits);; h = (h + 1) & mask)
        if (atomicCAS(&table[h], 0u, i + 1) == 0u) return;
}

__global__ void k_chain_link(const bf_dec_cand* __restrict__ cands, u32 count, const u32* __restrict__ table, u32 bits,
                             u32* __restrict__ prev, bf_dec_chain_result* __restrict__ res)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count || cands[i].status != DS_OK) return;
    const u32 mask = (1u << bits) - 1;
    const u64 end = cands[i].end;
    for (u32 h = dc_slot(end, bits); table[h]; h = (h + 1) & mask) {
        const u32 j = table[h] - 1;
        if (cands[j].pos == end) {
            if (atomicCAS(&prev[j], DC_NONE, i) != DC_NONE) atomicOr(&res->bad, 1u); // two frames end here
            return;
        }
    }
}

__global__ void k_chain_init(const bf_dec_cand* __restrict__ cands, u32 count, const u32* __restrict__ prev,
                             u32* __restrict__ up, u64* __restrict__ psum, u32* __restrict__ pcnt)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const u32 p = prev[i];
    up[i] = p == DC_NONE ? i : p;                       // a root points at itself
    psum[i] = p == DC_NONE ? 0ull : (u64)cands[p].block_size;
    pcnt[i] = p == DC_NONE ? 0u : 1u;
}

__global__ void k_chain_jump(u32 count, const u32* __restrict__ up0, const u64* __restrict__ psum0, const u32* __restrict__ pcnt0,
                             u32* __restrict__ up1, u64* __restrict__ psum1, u32* __restrict__ pcnt1)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const u32 a = up0[i];
    psum1[i] = psum0[i] + (a == i ? 0ull : psum0[a]);
    pcnt1[i] = pcnt0[i] + (a == i ? 0u : pcnt0[a]);
    up1[i] = up0[a];
}

__global__ void k_chain_emit_list(const bf_dec_cand* __restrict__ cands, u32 count, const u32* __restrict__ table, u32 bits,
                                  const u32* __restrict__ up, const u64* __restrict__ psum, const u32* __restrict__ pcnt,
                                  u64 total, bf_dec_emit* __restrict__ emits, u64* __restrict__ offs, u32* __restrict__ lens,
                                  bf_dec_chain_result* __restrict__ res)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    // the head: the candidate at offset 0
    u32 head = DC_NONE;
    const u32 mask = (1u << bits) - 1;
    for (u32 h = dc_slot(0, bits); table[h]; h = (h + 1) & mask)
        if (cands[table[h] - 1].pos == 0) { head = table[h] - 1; break; }
    if (head == DC_NONE) { if (i == 0) atomicOr(&res->bad, 2u); return; }
    if (up[i] != head || psum[i] >= total) return;      // not part of the stream (or past its end)
    const bf_dec_cand cd = cands[i];
    if (cd.status != DS_OK || cd.block_size > total - psum[i]) { atomicOr(&res->bad, 4u); return; }
    const u32 k = pcnt[i];
    bf_dec_emit e;
    e.pcm_frame = psum[i]; e.cand = i; e.n = cd.block_size;
    emits[k] = e;
    offs[k] = cd.pos;
    lens[k] = cd.block_size;
    atomicAdd(&res->n_emit, 1u);
    atomicMax((unsigned long long*)&res->covered, (unsigned long long)(psum[i] + cd.block_size));
}

// ---- host ------------------------------------------------------------------------------------------
static int dfail(int rc, const char* msg)
{
    b200flac_internal_set_error(msg);
    return rc;
}

#define DCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char m_[256]; \
        snprintf(m_, sizeof(m_), "%s failed: %s", #call, cudaGetErrorString(e_)); b200flac_internal_set_error(m_); \
        rc = 3; goto done; } } while (0)

// flacdec_read_metadata, src/decoders/flac.c:569-708: "fLaC", then blocks until the last-block flag;
// STREAMINFO gives the stream parameters, everything else is skipped here
extern "C" int b200flac_read_streaminfo(const uint8_t* flac, uint64_t n_bytes, b200flac_stream_info* info)
{
    if (!flac || !info) return dfail(1, "flac/info is NULL");
    memset(info, 0, sizeof(*info));
    if (n_bytes < 4 || memcmp(flac, "fLaC", 4) != 0) return dfail(1, "not a FLAC file");
    u64 p = 4;
    bool have = false;
    for (;;) {
        if (p + 4 > n_bytes) return dfail(2, "EOF while reading metadata");
        const u32 last = flac[p] >> 7, type = flac[p] & 0x7F;
        const u32 len = ((u32)flac[p + 1] << 16) | ((u32)flac[p + 2] << 8) | flac[p + 3];
        p += 4;
        if (p + len > n_bytes) return dfail(2, "EOF while reading metadata");
        if (type == 0) {
            if (len < 34) return dfail(2, "EOF while reading metadata");
            const u8* s = flac + p;
            info->min_block_size = ((u32)s[0] << 8) | s[1];
            info->max_block_size = ((u32)s[2] << 8) | s[3];
            info->min_frame_size = ((u32)s[4] << 16) | ((u32)s[5] << 8) | s[6];
            info->max_frame_size = ((u32)s[7] << 16) | ((u32)s[8] << 8) | s[9];
            info->sample_rate = ((u32)s[10] << 12) | ((u32)s[11] << 4) | (s[12] >> 4);
            info->channels = ((s[12] >> 1) & 7) + 1;
            info->bits_per_sample = (((u32)(s[12] & 1) << 4) | (s[13] >> 4)) + 1;
            info->total_pcm_frames = ((u64)(s[13] & 15) << 32) | ((u64)s[14] << 24) | ((u64)s[15] << 16) | ((u64)s[16] << 8) | s[17];
            memcpy(info->md5, s + 18, 16);
            // default channel mask from the channel count, flac.c:626-658
            static const u32 by_count[9] = {0, 0x4, 0x3, 0x7, 0x33, 0x37, 0x3F, 0x70F, 0x63F};
            info->channel_mask = by_count[info->channels];
            have = true;
        } else if (type == 4 && have) {
            // a WAVEFORMATEXTENSIBLE_CHANNEL_MASK entry overrides it when its bit count equals the
            // channel count (flacdec_read_vorbis_comment, flac.c:509-566; entries compared in upper case)
            const u8* c = flac + p;
            const u64 end = len;
            u64 q = 0;
            auto rd32 = [&](u32* v) { if (q + 4 > end) return false; *v = (u32)c[q] | ((u32)c[q + 1] << 8) | ((u32)c[q + 2] << 16) | ((u32)c[q + 3] << 24); q += 4; return true; };
            u32 n = 0, lines = 0;
            if (rd32(&n) && q + n <= end) {
                q += n;
                if (rd32(&lines)) {
                    static const char prefix[] = "WAVEFORMATEXTENSIBLE_CHANNEL_MASK=";
                    const size_t plen = sizeof(prefix) - 1;
                    for (; lines > 0; lines--) {
                        u32 l = 0;
                        if (!rd32(&l) || q + l > end) break;
                        if (l > plen) {
                            bool match = true;
                            for (size_t k = 0; k < plen && match; k++) {
                                u8 ch = c[q + k];
                                if (ch >= 'a' && ch <= 'z') ch = (u8)(ch - 32);
                                match = ch == (u8)prefix[k];
                            }
                            if (match) {
                                char hex[17] = {0};
                                const size_t hl = std::min<size_t>(l - plen, 16);
                                memcpy(hex, c + q + plen, hl);
                                const u32 mask = (u32)strtoul(hex, nullptr, 16);
                                if ((u32)__builtin_popcount(mask) == info->channels) info->channel_mask = mask;
                            }
                        }
                        q += l;
                    }
                }
            }
        }
        p += len;
        if (last) break;
    }
    if (!have) return dfail(1, "STREAMINFO not found");
    info->first_frame_offset = p;
    return 0;
}

// Device buffers are kept between calls (grow-only, one set per process, calls serialised): allocating
// and freeing 1-2 GB per decode costs more than the kernels.  b200flac_pool_clear() releases them.
struct DecWork {
    int device = -1;
    u8* d_data = nullptr;        size_t cap_data = 0;
    bf_dec_cand* d_cands = nullptr; size_t cap_cands = 0;
    bf_dec_emit* d_emits = nullptr; size_t cap_emits = 0;
    int* d_scratch = nullptr;    size_t cap_scratch = 0;
    u8* d_pcm = nullptr;         size_t cap_pcm = 0;
    u8* d_chain = nullptr;       size_t cap_chain = 0;
    u32* d_count = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    void release()
    {
        if (device >= 0) cudaSetDevice(device);
        cudaFree(d_data); cudaFree(d_cands); cudaFree(d_emits); cudaFree(d_count); cudaFree(d_scratch); cudaFree(d_pcm); cudaFree(d_chain);
        for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
        d_data = nullptr; d_cands = nullptr; d_emits = nullptr; d_count = nullptr; d_scratch = nullptr; d_pcm = nullptr; d_chain = nullptr; cap_chain = 0;
        cap_data = cap_cands = cap_emits = cap_scratch = cap_pcm = 0;
        device = -1;
    }
};
static DecWork g_work;
static pthread_mutex_t g_work_mu = PTHREAD_MUTEX_INITIALIZER;

extern "C" void b200flac_internal_decoder_clear(void)
{
    pthread_mutex_lock(&g_work_mu);
    g_work.release();
    pthread_mutex_unlock(&g_work_mu);
}

template <typename T>
static cudaError_t grow(T** p, size_t* cap, size_t bytes)
{
    if (*cap >= bytes && *p) return cudaSuccess;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    const cudaError_t e = cudaMalloc((void**)p, bytes);
    if (e == cudaSuccess) *cap = bytes;
    return e;
}

#include <chrono>
static double dec_now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// frames: the bytes from the first frame to the end of the file (host or device memory, see flags)
static int decode_core(const b200flac_stream_info* info, const uint8_t* frames, uint64_t n_bytes, int frames_on_device,
                       int device, uint8_t* pcm_out, int pcm_on_device, uint64_t pcm_capacity,
                       uint64_t** frame_offsets, uint32_t** frame_pcm_frames, uint64_t* n_frames, float* kernel_ms)
{
    int rc = 0;
    if (b200flac_device_count() <= 0) return dfail(3, "no CUDA device available: the B200 FLAC engine has no CPU fallback");
    if (info->bits_per_sample % 8 || info->bits_per_sample < 8 || info->bits_per_sample > 24 || info->channels < 1 ||
        info->channels > B200FLAC_MAX_CHANNELS || info->max_block_size == 0)
        return dfail(3, "unsupported stream parameters");
    const u32 C = info->channels, B = info->bits_per_sample / 8;
    const u64 total = info->total_pcm_frames;
    const u64 pcm_bytes = total * C * B;
    if (pcm_out && pcm_capacity < pcm_bytes) return dfail(3, "PCM buffer too small");
    if (n_frames) *n_frames = 0;
    if (total == 0) return 0; // the reference's loops run while remaining samples > 0 (flac.c:196, :1402)

    DecStream S;
    S.sample_rate = info->sample_rate; S.channels = C; S.bits_per_sample = info->bits_per_sample;
    S.max_block_size = info->max_block_size;
    struct Guard { Guard() { pthread_mutex_lock(&g_work_mu); } ~Guard() { pthread_mutex_unlock(&g_work_mu); } } guard;
    DecWork& w = g_work;
    std::vector<bf_dec_cand> cands;
    std::vector<bf_dec_emit> emits;
    std::vector<u64> offs;
    std::vector<u32> lens;
    const u32 row_stride = (info->max_block_size + 3) & ~3u;
    // candidates: the real frames (at least min_block_size PCM frames each, except the last) plus look-alikes
    const u32 minb = info->min_block_size ? info->min_block_size : 16;
    u32 cap = (u32)std::min<u64>(total / minb + n_bytes / (1u << 16) + 4096, 0x7FFFFFFFu);
    u32 count = 0;
    double t_d2h = 0.0;
    bool chained = false;
    u32 n_emit = 0;
    u64* d_offs = nullptr;
    u32* d_lens = nullptr;
    const u8* d_frames = nullptr;
    if (w.device != device) w.release();
    DCK(cudaSetDevice(device));
    w.device = device;
    for (auto& e : w.ev) if (!e) DCK(cudaEventCreate(&e));
    if (frames_on_device) d_frames = frames;
    else {
        DCK(grow(&w.d_data, &w.cap_data, n_bytes + 64));
        DCK(cudaMemcpy(w.d_data, frames, n_bytes, cudaMemcpyHostToDevice));
        DCK(cudaMemset(w.d_data + n_bytes, 0, 64));
        d_frames = w.d_data;
    }
    if (!w.d_count) DCK(cudaMalloc((void**)&w.d_count, sizeof(u32)));
    for (int attempt = 0; attempt < 2; attempt++) {
        DCK(grow(&w.d_cands, &w.cap_cands, (size_t)cap * sizeof(bf_dec_cand)));
        DCK(cudaMemset(w.d_count, 0, sizeof(u32)));
        DCK(cudaEventRecord(w.ev[0]));
        k_dec_scan<<<148 * 8, 256>>>(d_frames, n_bytes, S, w.d_cands, cap, w.d_count);
        DCK(cudaGetLastError());
        DCK(cudaEventRecord(w.ev[1]));
        DCK(cudaMemcpy(&count, w.d_count, sizeof(u32), cudaMemcpyDeviceToHost));
        if (count <= cap) break;
        cap = count; // more look-alikes than allowed for: size for what was found
    }
    if (count == 0) { rc = dfail(1, "invalid sync code"); goto done; }
    {
        // one int32 row per (candidate, channel) for the whole stream: bounded by what the device has, said in words
        // (a file stuffed with header look-alikes has a candidate count of its author's choosing)
        const size_t need = (size_t)count * C * row_stride * sizeof(int);
        size_t free_b = 0, total_b = 0;
        if (need > w.cap_scratch && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && need > free_b + w.cap_scratch) {
            char m[200];
            snprintf(m, sizeof(m), "stream needs %.1f GB of decode scratch (%u frame candidates x %u channels x %u samples); "
                     "the device has %.1f GB free", need / 1e9, count, C, row_stride, (free_b + w.cap_scratch) / 1e9);
            rc = dfail(3, m);
            goto done;
        }
        DCK(grow(&w.d_scratch, &w.cap_scratch, need));
    }
    DCK(cudaEventRecord(w.ev[2]));
    k_dec_frames<<<(count + 31) / 32, 32>>>(d_frames, n_bytes, S, w.d_cands, count, w.d_scratch, row_stride);
    DCK(cudaGetLastError());
    DCK(cudaEventRecord(w.ev[3]));
    // ---- the chain on the device (the usual case) ----
    if (!getenv("B200FLAC_DEC_HOST_CHAIN") && count < (1u << 30)) {
        u32 bits = 4;
        while ((1u << bits) < 2 * count) bits++;
        const size_t T = (size_t)1 << bits, n8 = ((size_t)count + 1) & ~(size_t)1;
        // u64 arrays first: psum0, psum1, offs; then u32: lens, prev, up0, up1, pcnt0, pcnt1, table; then the result
        const size_t bytes = 3 * n8 * 8 + 6 * n8 * 4 + T * 4 + 64;
        DCK(grow(&w.d_chain, &w.cap_chain, bytes));
        u64* psum0 = (u64*)w.d_chain; u64* psum1 = psum0 + n8; d_offs = psum1 + n8;
        d_lens = (u32*)(d_offs + n8);
        u32* prev = d_lens + n8; u32* up0 = prev + n8; u32* up1 = up0 + n8; u32* pcnt0 = up1 + n8; u32* pcnt1 = pcnt0 + n8;
        u32* table = pcnt1 + n8;
        bf_dec_chain_result* d_res = (bf_dec_chain_result*)(table + T);
        DCK(grow(&w.d_emits, &w.cap_emits, (size_t)count * sizeof(bf_dec_emit)));
        DCK(cudaMemsetAsync(table, 0, T * 4 + 64));
        DCK(cudaMemsetAsync(prev, 0xFF, n8 * 4));
        const u32 g = (count + 255) / 256;
        k_chain_insert<<<g, 256>>>(w.d_cands, count, table, bits);
        k_chain_link<<<g, 256>>>(w.d_cands, count, table, bits, prev, d_res);
        k_chain_init<<<g, 256>>>(w.d_cands, count, prev, up0, psum0, pcnt0);
        u32 rounds = 1;
        while ((1u << rounds) < count) rounds++;
        for (u32 r = 0; r < rounds; r++) {
            k_chain_jump<<<g, 256>>>(count, up0, psum0, pcnt0, up1, psum1, pcnt1);
            std::swap(up0, up1); std::swap(psum0, psum1); std::swap(pcnt0, pcnt1);
        }
        k_chain_emit_list<<<g, 256>>>(w.d_cands, count, table, bits, up0, psum0, pcnt0, total, w.d_emits, d_offs, d_lens, d_res);
        DCK(cudaGetLastError());
        bf_dec_chain_result res;
        DCK(cudaMemcpy(&res, d_res, sizeof(res), cudaMemcpyDeviceToHost));
        if (res.bad == 0 && res.covered == total && res.n_emit > 0) { chained = true; n_emit = res.n_emit; }
    }
    if (!chained) {
    cands.resize(count);
    DCK(cudaMemcpy(cands.data(), w.d_cands, (size_t)count * sizeof(bf_dec_cand), cudaMemcpyDeviceToHost));
    t_d2h = dec_now();

    // ---- the reference's frame loop (flac.c:196-268, :1402-1476) over the decoded candidates: the same walk on
    // the host, for streams the device chain does not vouch for -- it finds and words the error ----
    {
        // position -> candidate through an open-addressing hash table (no sort: the walk itself visits the
        // frames in stream order; sorting 38,760 keys cost three times the rest of this block)
        u32 bits = 4;
        while ((1u << bits) < 2 * count) bits++;
        const u32 mask = (1u << bits) - 1;
        std::vector<u32> table((size_t)mask + 1, 0u);
        auto slot_of = [&](u64 pos) { return (u32)((pos * 0x9E3779B97F4A7C15ull) >> (64 - bits)); };
        for (u32 i = 0; i < count; i++) {
            u32 h = slot_of(cands[i].pos);
            while (table[h]) h = (h + 1) & mask;
            table[h] = i + 1;
        }
        emits.reserve(count); offs.reserve(count); lens.reserve(count);
        u64 pos = 0, done_frames = 0;
        while (done_frames < total) {
            u32 found = 0;
            for (u32 h = slot_of(pos); table[h]; h = (h + 1) & mask)
                if (cands[table[h] - 1].pos == pos) { found = table[h]; break; }
            u32 status;
            const bf_dec_cand* cd = nullptr;
            if (found) {
                cd = &cands[found - 1];
                status = cd->status;
            } else if (frames_on_device) {
                status = DS_INVALID_SYNC_CODE; // (the header bytes are not on the host to say more)
            } else {
                DecHeader h;
                status = pos >= n_bytes ? (u32)DS_EOF : dec_parse_header(frames + pos, n_bytes - pos, S, &h);
                if (status == DS_OK) status = DS_ERROR;
            }
            if (status == DS_OK && cd->block_size > total - done_frames) status = DS_MALFORMED;
            if (status != DS_OK) { rc = dfail(status == DS_EOF ? 2 : 1, ds_strerror(status)); goto done; }
            bf_dec_emit e;
            e.pcm_frame = done_frames; e.cand = found - 1; e.n = cd->block_size;
            emits.push_back(e);
            offs.push_back(pos);
            lens.push_back(cd->block_size);
            done_frames += cd->block_size;
            pos = cd->end;
        }
    }
    if (getenv("B200FLAC_DEC_TIMING")) fprintf(stderr, "decode host: walk %.3f ms (after the candidates' D2H)\n", (dec_now() - t_d2h) * 1e3);
    n_emit = (u32)emits.size();
    DCK(grow(&w.d_emits, &w.cap_emits, emits.size() * sizeof(bf_dec_emit)));
    DCK(cudaMemcpy(w.d_emits, emits.data(), emits.size() * sizeof(bf_dec_emit), cudaMemcpyHostToDevice));
    } else if (frame_offsets || frame_pcm_frames) {
        offs.resize(n_emit); lens.resize(n_emit);
        DCK(cudaMemcpy(offs.data(), d_offs, (size_t)n_emit * sizeof(u64), cudaMemcpyDeviceToHost));
        DCK(cudaMemcpy(lens.data(), d_lens, (size_t)n_emit * sizeof(u32), cudaMemcpyDeviceToHost));
    }
    if (pcm_out) {
        u8* d_pcm = pcm_out;
        if (!pcm_on_device) { DCK(grow(&w.d_pcm, &w.cap_pcm, pcm_bytes + 64)); d_pcm = w.d_pcm; }
        k_dec_emit<<<n_emit, 256>>>(w.d_cands, w.d_emits, w.d_scratch, row_stride, S, d_pcm);
        DCK(cudaGetLastError());
        DCK(cudaEventRecord(w.ev[4]));
        if (!pcm_on_device) DCK(cudaMemcpy(pcm_out, d_pcm, pcm_bytes, cudaMemcpyDeviceToHost));
        else DCK(cudaDeviceSynchronize());
        if (kernel_ms) {
            DCK(cudaEventElapsedTime(&kernel_ms[0], w.ev[0], w.ev[1]));
            DCK(cudaEventElapsedTime(&kernel_ms[1], w.ev[2], w.ev[3]));
            DCK(cudaEventElapsedTime(&kernel_ms[2], w.ev[3], w.ev[4])); // chain (device, or the host walk) + emit
        }
    }
    if (n_frames) *n_frames = n_emit;
    if (frame_offsets) {
        *frame_offsets = (uint64_t*)malloc((offs.size() ? offs.size() : 1) * sizeof(uint64_t));
        memcpy(*frame_offsets, offs.data(), offs.size() * sizeof(uint64_t));
    }
    if (frame_pcm_frames) {
        *frame_pcm_frames = (uint32_t*)malloc((lens.size() ? lens.size() : 1) * sizeof(uint32_t));
        memcpy(*frame_pcm_frames, lens.data(), lens.size() * sizeof(uint32_t));
    }
done:
    return rc;
}


// Reads a whole file; every failure (open, seek, tell, allocation, short read) is an error return, never an
// exception through the C ABI.
static int read_whole_file(const char* path, std::vector<uint8_t>& data)
{
    FILE* f = fopen(path, "rb");
    if (!f) {
        char msg[400];
        snprintf(msg, sizeof(msg), "cannot open \"%.300s\" for reading", path);
        return dfail(2, msg);
    }
    off_t end = -1;
    if (fseeko(f, 0, SEEK_END) == 0) end = ftello(f);
    if (end < 0 || fseeko(f, 0, SEEK_SET) != 0) { fclose(f); return dfail(2, "read error"); }
    try { data.resize((size_t)end + 1); } catch (const std::bad_alloc&) { fclose(f); return dfail(3, "out of memory reading the file"); }
    const size_t got = fread(data.data(), 1, (size_t)end, f);
    fclose(f);
    if (got != (size_t)end) return dfail(2, "read error");
    data.resize((size_t)end);
    return 0;
}

// STREAMINFO's total (36 bits, straight from the file) against what the file can hold: a frame is at least
// 8 bytes and at most 65535 PCM frames, so a larger total can only end the way the reference's frame loop
// ends on such a file -- at EOF (src/decoders/flac.c:196-268).  Checked before anything is sized from it.
static int check_total_against_size(const b200flac_stream_info* info, uint64_t frame_bytes)
{
    if (info->bits_per_sample % 8 || info->bits_per_sample < 8 || info->bits_per_sample > 24 || info->channels < 1 ||
        info->channels > B200FLAC_MAX_CHANNELS || info->max_block_size == 0)
        return dfail(3, "unsupported stream parameters");
    if (info->total_pcm_frames > (frame_bytes / 8 + 1) * 65535ull) return dfail(2, ds_strerror(DS_EOF));
    return 0;
}

extern "C" int b200flac_decode_memory(const uint8_t* flac, uint64_t n_bytes, int device, uint8_t* pcm,
                                      uint64_t pcm_capacity, b200flac_stream_info* info_out, int check_md5,
                                      uint64_t** frame_offsets, uint32_t** frame_pcm_frames, uint64_t* n_frames,
                                      float* kernel_ms)
{
    b200flac_stream_info info;
    int rc = b200flac_read_streaminfo(flac, n_bytes, &info);
    if (rc) return rc;
    if (info_out) *info_out = info;
    rc = check_total_against_size(&info, n_bytes - info.first_frame_offset);
    if (rc) return rc;
    if (!pcm) return 0; // sizing call: the caller may now size its buffer from info_out->total_pcm_frames
    rc = decode_core(&info, flac + info.first_frame_offset, n_bytes - info.first_frame_offset, 0, device, pcm, 0,
                     pcm_capacity, frame_offsets, frame_pcm_frames, n_frames, kernel_ms);
    if (rc) return rc;
    if (check_md5) {
        // FlacDecoder_verify_okay, flac.c:479-490: a blank MD5 in STREAMINFO always passes
        static const uint8_t blank[16] = {0};
        uint8_t digest[16];
        b200flac_internal_md5(pcm, (size_t)(info.total_pcm_frames * info.channels * (info.bits_per_sample / 8)), digest);
        if (memcmp(info.md5, blank, 16) != 0 && memcmp(info.md5, digest, 16) != 0)
            return dfail(1, "MD5 mismatch at end of stream");
    }
    return 0;
}

extern "C" int b200flac_decode_device(const b200flac_stream_info* info, const void* d_frames, uint64_t n_bytes,
                                      int device, void* d_pcm, uint64_t pcm_capacity, uint64_t* n_frames,
                                      float* kernel_ms)
{
    if (!info || !d_frames || !d_pcm) return dfail(3, "info/d_frames/d_pcm is NULL");
    if (((size_t)d_frames & 15) || ((size_t)d_pcm & 3)) return dfail(3, "d_frames must be 16-byte aligned, d_pcm 4-byte aligned");
    return decode_core(info, (const uint8_t*)d_frames, n_bytes, 1, device, (uint8_t*)d_pcm, 1, pcm_capacity, nullptr,
                       nullptr, n_frames, kernel_ms);
}

extern "C" int b200flac_verify_file(const char* flac_filename, int device)
{
    if (!flac_filename) return dfail(3, "filename is NULL");
    std::vector<uint8_t> data;
    int rc = read_whole_file(flac_filename, data);
    if (rc) return rc;
    const u64 n = data.size();
    b200flac_stream_info info;
    rc = b200flac_read_streaminfo(data.data(), n, &info);
    if (rc) return rc;
    rc = check_total_against_size(&info, n - info.first_frame_offset);
    if (rc) return rc;
    std::vector<uint8_t> pcm;
    try { pcm.resize((size_t)(info.total_pcm_frames * info.channels * (info.bits_per_sample / 8)) + 1); }
    catch (const std::bad_alloc&) { return dfail(3, "out of memory for the decoded PCM"); }
    return b200flac_decode_memory(data.data(), n, device, pcm.data(), pcm.size(), nullptr, 1, nullptr, nullptr, nullptr, nullptr);
}

// FLAC file -> RIFF WAVE file: WaveAudio.from_pcm(wave, FlacAudio(flac).to_pcm()) in one call
// (audiotools/wav.py:357-418 wave_header, :660-729 from_pcm): plain `fmt ` for up to 2 channels at up to
// 16 bits, WAVEFORMATEXTENSIBLE with the stream's channel mask otherwise, 8-bit samples unsigned, and the
// reference's pad byte, which follows the parity of the PCM frame count.
extern "C" int b200flac_decode_to_wave(const char* flac_filename, const char* wave_filename, int device)
{
    if (!flac_filename || !wave_filename) return dfail(3, "filename is NULL");
    std::vector<uint8_t> data;
    int rc = read_whole_file(flac_filename, data);
    if (rc) return rc;
    const u64 n = data.size();
    b200flac_stream_info info;
    rc = b200flac_read_streaminfo(data.data(), n, &info);
    if (rc) return rc;
    rc = check_total_against_size(&info, n - info.first_frame_offset);
    if (rc) return rc;
    const u32 B = info.bits_per_sample / 8, C = info.channels;
    const u64 data_size = info.total_pcm_frames * C * B;
    std::vector<uint8_t> fmt;
    auto le = [&](u64 v, int bytes) { for (int i = 0; i < bytes; i++) fmt.push_back((uint8_t)(v >> (8 * i))); };
    const bool plain = C <= 2 && info.bits_per_sample <= 16;
    le(plain ? 1 : 0xFFFE, 2); le(C, 2); le(info.sample_rate, 4); le((u64)info.sample_rate * C * B, 4); le(C * B, 2);
    le(info.bits_per_sample, 2);
    if (!plain) {
        static const uint8_t guid[16] = {0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x10, 0x00, 0x80, 0x00, 0x00, 0xaa, 0x00, 0x38, 0x9b, 0x71};
        static const u32 by_count[7] = {0, 0x4, 0x3, 0x7, 0x33, 0x37, 0x3F};
        u32 mask = info.channel_mask;
        if (mask == 0) mask = C <= 6 ? by_count[C] : 0;
        le(22, 2); le(info.bits_per_sample, 2); le(mask, 4);
        fmt.insert(fmt.end(), guid, guid + 16);
    }
    const u64 total_size = 4 + 8 + fmt.size() + 8 + data_size + (data_size % 2);
    if (total_size >= (1ull << 32)) return dfail(1, "total size too large for wave file");
    std::vector<uint8_t> pcm;
    try { pcm.resize((size_t)data_size + 1); } catch (const std::bad_alloc&) { return dfail(3, "out of memory for the decoded PCM"); }
    rc = b200flac_decode_memory(data.data(), n, device, pcm.data(), data_size, nullptr, 1, nullptr, nullptr, nullptr, nullptr);
    if (rc) return rc;
    if (B == 1) for (u64 i = 0; i < data_size; i++) pcm[i] ^= 0x80; // to_bytes(False, signed = bits_per_sample > 8)
    FILE* o = fopen(wave_filename, "wb");
    if (!o) {
        char msg[400];
        snprintf(msg, sizeof(msg), "cannot open \"%.300s\" for writing", wave_filename);
        return dfail(2, msg);
    }
    std::vector<uint8_t> head;
    auto put = [&](const void* p, size_t k) { head.insert(head.end(), (const uint8_t*)p, (const uint8_t*)p + k); };
    auto put32 = [&](u64 v) { for (int i = 0; i < 4; i++) head.push_back((uint8_t)(v >> (8 * i))); };
    put("RIFF", 4); put32(total_size); put("WAVE", 4); put("fmt ", 4); put32(fmt.size()); put(fmt.data(), fmt.size());
    put("data", 4); put32(data_size);
    bool ok = fwrite(head.data(), 1, head.size(), o) == head.size() && (data_size == 0 || fwrite(pcm.data(), 1, data_size, o) == data_size);
    if (ok && (info.total_pcm_frames % 2)) ok = fputc(0, o) != EOF; // wav.py:707-709
    if (fclose(o) != 0) ok = false;
    if (!ok) { remove(wave_filename); return dfail(2, "write error"); }
    return 0;
}
