// b200tta.cu -- the TTA encoder of the engine (include/b200tta.h), sm_100a.
//
// What the reference does one frame at a time (src/encoders/tta.c:144-262) is done here for every frame of a
// stream at once, in three kernels:
//
//   k_tta_residual   one THREAD per (frame, channel).  TTA's predictor is a sign-LMS adaptive filter whose
//                    weights move with the sign of the previous residual (hybrid_filter, tta.c:314-399), and
//                    its Rice parameters adapt after every code (tta.c:199-246): both are serial chains along
//                    a channel, but independent between channels and between frames, so the parallelism is
//                    frames x channels (6,890 chains for an hour of stereo).  The thread walks its channel:
//                    correlate_channels (:264-293) and fixed_prediction (:295-312) fused into the sample fetch,
//                    the filter, the adaptive Rice state; it stores one code descriptor per sample
//                    (bit length, Rice parameter, low bits) and the channel's bit total.
//   k_tta_sizes      frame sizes from the bit totals, exclusive scan -> every frame's byte offset.
//   k_tta_pack       one CTA per frame.  The frame's codes interleave the channels sample by sample
//                    (tta.c:199-200: i outer, c inner), so a code's bit position is a prefix sum over the
//                    interleaved descriptors: the CTA scans them in chunks of 1024 and every thread ORs its
//                    four codes into the output at their final position, least significant bit first (the
//                    little-endian BitstreamWriter, src/bitstream.c).  Then the frame's CRC-32
//                    (src/common/tta_crc.c): 256-byte segments, slicing-by-4 tables per thread, segment CRCs
//                    moved to their place with one carry-less multiply by a tabulated power of x and XOR-reduced.
//
// Same bytes as the reference: tests/test_tta_gpu.py compares with the tests' CPU checker, which is pinned to
// the compiled reference encoder.  No CPU fallback.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/b200tta.h"

typedef unsigned long long u64;
typedef unsigned int u32;

#define TTA_SEG 256u            // bytes per CRC segment
#define TTA_POW_N 16384u        // tabulated powers x^(8 * TTA_SEG * j): frames up to 4 MB; longer ones compute theirs
#define TTA_POLY 0xEDB88320u

static thread_local char t_err[512] = "";
extern "C" const char* b200tta_last_error(void) { return t_err; }
extern "C" void b200tta_free(void* p) { free(p); }
static int tfail(const char* msg) { snprintf(t_err, sizeof(t_err), "%s", msg); return 1; }
#define TCK(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            snprintf(t_err, sizeof(t_err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            rc = 1;                                                                                      \
            goto done;                                                                                   \
        }                                                                                                \
    } while (0)

extern "C" uint32_t b200tta_block_size(uint32_t sample_rate) { return (uint32_t)(((u64)sample_rate * 256) / 245); }

struct TtaFrame {
    u64 pcm_off;     // first PCM frame of the TTA frame inside the pcm buffer
    u64 ebase;       // first descriptor of the frame inside a channel's plane (frames are padded to 4 descriptors)
    u32 n;           // PCM frames in it
    u32 pad;
};

// code descriptor: low word = length in bits (ones + stop bit + k), high word = k << 27 | low bits
__device__ __forceinline__ u64 tta_desc(u32 msb, u32 k, u32 lsb) { return (u64)(msb + 1u + k) | ((u64)((k << 27) | lsb) << 32); }

// ---------------------------------------------------------------------------------------------------------
// one thread per (frame, channel)
// ---------------------------------------------------------------------------------------------------------
template <u32 B>
__global__ void __launch_bounds__(32)
k_tta_residual(const uint8_t* __restrict__ pcm, const TtaFrame* __restrict__ frames, u32 n_frames, u32 C, u32 bps,
               u64* __restrict__ desc, u64 plane, u64* __restrict__ bits_out)
{
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_frames * C) return;
    const u32 f = t / C, c = t % C;
    const TtaFrame fr = frames[f];
    const u32 n = fr.n;
    // correlate_channels, tta.c:264-293: channel c < C-1 codes ch[c+1] - ch[c]; the last one codes
    // ch[C-1] - (ch[C-1] - ch[C-2]) / 2 (truncating division); a single channel is coded as it is
    const u32 ca = C == 1 ? 0u : (c < C - 1 ? c : C - 2), cb = C == 1 ? 0u : (c < C - 1 ? c + 1 : C - 1);
    const bool last = C > 1 && c == C - 1;
    const uint8_t* src = pcm + fr.pcm_off * C * B;
    const u32 fshift = bps == 8 ? 4u : 5u;                 // fixed_prediction, tta.c:300
    const int hshift = bps == 16 ? 9 : 10;                 // hybrid_filter, tta.c:320-321
    const u32 round = 1u << (hshift - 1);
    u32 qm[8], dl[8];
    int dx[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { qm[j] = 0; dl[j] = 0; dx[j] = 0; }
    u32 k0 = 10, sum0 = 1u << 14, k1 = 10, sum1 = 1u << 14;    // tta.c:193-196
    int prevx = 0, rprev = 0;
    u64 bits = 0;
    u64* out = desc + (u64)c * plane + fr.ebase;
    u64 d4[4];
    // The channel is walked in blocks of 8 samples; the next block's PCM is fetched before the current one is
    // filtered, so no load sits on the serial chain.
    constexpr int TB = 8;
    // raw loads only: nothing that depends on a load's result is issued before the block that uses it (a sign
    // extension right behind the load made the warp wait for memory on the spot)
    u32 ra[TB][B == 3 ? 3 : 1], rb[TB][B == 3 ? 3 : 1];
    auto raw = [&](const uint8_t* p, u32 (&w)[B == 3 ? 3 : 1]) {
        if (B == 2) w[0] = *(const unsigned short*)p;
        else if (B == 1) w[0] = *p;
        else { w[0] = p[0]; w[1 % (B == 3 ? 3 : 1)] = p[1]; w[2 % (B == 3 ? 3 : 1)] = p[2]; }
    };
    auto cook = [&](const u32 (&w)[B == 3 ? 3 : 1]) -> int {
        if (B == 2) return (int)(short)w[0];
        if (B == 1) return (int)(signed char)w[0];
        return ((int)(w[0] << 8 | w[1 % (B == 3 ? 3 : 1)] << 16 | w[2 % (B == 3 ? 3 : 1)] << 24)) >> 8;
    };
    auto fetch = [&](u32 i0) {
#pragma unroll
        for (int j = 0; j < TB; j++) {
            const u32 i = min(i0 + (u32)j, n - 1);
            const uint8_t* p = src + (size_t)i * C * B;
            raw(p + ca * B, ra[j]);
            if (C > 1) raw(p + cb * B, rb[j]);
        }
    };
    fetch(0);
    for (u32 i0 = 0; i0 < n; i0 += TB) {
        int pr[TB];
        // correlate_channels + fixed_prediction (tta.c:306-310) of the block: independent of the filter.  (The
        // reference forms (x << shift) - x in 64 bits; samples of at most 25 bits keep it inside 32.)
#pragma unroll
        for (int j = 0; j < TB; j++) {
            const int a = cook(ra[j]), b = C > 1 ? cook(rb[j]) : 0;
            const int x = C == 1 ? a : (last ? b - ((b - a) / 2) : b - a);
            pr[j] = (i0 + j) ? x - (((int)((u32)prevx << fshift) - prevx) >> fshift) : x;
            prevx = x;
        }
        if (i0 + TB < n) fetch(i0 + TB);
#pragma unroll
        for (int j = 0; j < TB; j++) {
            const u32 i = i0 + j;
            if (i >= n) break;
            const int pred = pr[j];
            // hybrid_filter, tta.c:329-397.  The sum is 32-bit wrap-around arithmetic (every operand is int32_t), so
            // it may be regrouped freely: with s = sign of the previous residual, sum of dl * (qm + s dx) =
            // (round + sum of dl * qm) + s * (sum of dl * dx) -- both sums are formed before s is known, which
            // leaves sign -> multiply-add -> shift -> subtract on the serial chain instead of eight multiply-adds.
            u32 sa = round, sb = 0;
#pragma unroll
            for (int t = 0; t < 8; t++) { sa += dl[t] * qm[t]; sb += dl[t] * (u32)dx[t]; }
            const int sgn = i ? (rprev > 0) - (rprev < 0) : 0;
            const int r = i ? pred - ((int)(sa + (u32)sgn * sb) >> hshift) : pred;
#pragma unroll
            for (int t = 0; t < 8; t++) qm[t] += (u32)(sgn * dx[t]);
            rprev = r;
            dx[0] = dx[1]; dx[1] = dx[2]; dx[2] = dx[3]; dx[3] = dx[4];
            dx[4] = ((int)dl[4] >= 0) ? 1 : -1;
            dx[5] = ((int)dl[5] >= 0) ? 2 : -2;
            dx[6] = ((int)dl[6] >= 0) ? 2 : -2;
            dx[7] = ((int)dl[7] >= 0) ? 4 : -4;
            {
                const u32 pp = (u32)pred, t7 = pp - dl[7], t6 = t7 - dl[6], t5 = t6 - dl[5];
                dl[0] = dl[1]; dl[1] = dl[2]; dl[2] = dl[3]; dl[3] = dl[4];
                dl[4] = t5; dl[5] = t6; dl[6] = t7; dl[7] = pp;
            }
            // adaptive Rice code, tta.c:201-246
            const u32 u = r > 0 ? ((u32)r * 2u) - 1u : (u32)(-r) * 2u;
            u64 d;
            if (u < (1u << k0)) {
                d = tta_desc(0, k0, u);
            } else {
                const u32 shifted = u - (1u << k0);
                const u32 msb = 1u + (shifted >> k1);
                const u32 lsb = shifted - ((msb - 1u) << k1);
                d = tta_desc(msb, k1, lsb);
                sum1 += shifted - (sum1 >> 4);
                if (k1 > 0 && (int)sum1 < (1 << (k1 + 4))) k1 -= 1;
                else if ((int)sum1 > (1 << (k1 + 5))) k1 += 1;
            }
            sum0 += u - (sum0 >> 4);
            if (k0 > 0 && (int)sum0 < (1 << (k0 + 4))) k0 -= 1;
            else if ((int)sum0 > (1 << (k0 + 5))) k0 += 1;
            if ((u32)(d >> 59) > 27u) __trap();      // (a Rice parameter beyond the descriptor's field: input outside any PCM width)
            bits += (u32)d;
            // descriptors leave in 32-byte pieces (a frame's plane region starts on a multiple of 4)
            d4[j & 3] = d;
            if ((j & 3) == 3) {
                *(ulonglong2*)(out + i - 3) = make_ulonglong2(d4[0], d4[1]);
                *(ulonglong2*)(out + i - 1) = make_ulonglong2(d4[2], d4[3]);
            }
        }
    }
    for (u32 i = n & ~3u; i < n; i++) out[i] = d4[i & 3];
    bits_out[t] = bits;
}

// frame sizes (bits of all channels, byte aligned, + CRC-32) and their exclusive scan; one CTA
__global__ void __launch_bounds__(1024)
k_tta_sizes(const u64* __restrict__ bits, u32 n_frames, u32 C, u32* __restrict__ frame_bytes, u64* __restrict__ frame_off,
            u64* __restrict__ total)
{
    __shared__ u64 red[33];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 base = 0; base < n_frames; base += blockDim.x) {
        const u32 f = base + threadIdx.x;
        u64 v = 0;
        if (f < n_frames) {
            u64 b = 0;
            for (u32 c = 0; c < C; c++) b += bits[(size_t)f * C + c];
            v = ((b + 7) >> 3) + 4;
            frame_bytes[f] = v > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)v;
        }
        u64 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= (u32)o) inc += t;
        }
        if (lane == 31) red[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            u64 w = red[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u64 t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
                if (lane >= (u32)o) winc += t;
            }
            red[lane] = winc - w;
            if (lane == 31) red[32] = winc;
        }
        __syncthreads();
        if (f < n_frames) frame_off[f] = carry + red[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += red[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void k_tta_zero(uint4* __restrict__ out, const u64* __restrict__ total, u64 capacity)
{
    u64 nbytes = *total + 16;
    if (nbytes > capacity) nbytes = capacity;
    const u64 n16 = (nbytes + 15) >> 4;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (u64)gridDim.x * blockDim.x)
        out[i] = make_uint4(0, 0, 0, 0);
}

// a * b modulo the CRC-32 polynomial, reflected bit order (x^0 is bit 31)
__host__ __device__ inline u32 tta_multmodp(u32 a, u32 b)
{
    u32 m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ TTA_POLY : b >> 1;
    }
    return p;
}

// x^(8 * nbytes) modulo the polynomial
__host__ __device__ inline u32 tta_xpow8(u64 nbytes)
{
    u64 n = nbytes * 8;
    u32 p = 1u << 31, base = 1u << 30;
    while (n) {
        if (n & 1) p = tta_multmodp(base, p);
        base = tta_multmodp(base, base);
        n >>= 1;
    }
    return p;
}

// LSB-first writer into zero-initialised 32-bit words (little-endian byte order = stream order); every word is
// OR-merged, so runs of neighbouring threads -- and of neighbouring frames -- may share words
struct TtaSink {
    u32* words;
    u64 widx;
    u64 acc;
    u32 fill;
    __device__ __forceinline__ void init(u32* base, u64 bitpos) { words = base; widx = bitpos >> 5; fill = (u32)(bitpos & 31); acc = 0; }
    __device__ __forceinline__ void put(u32 v, u32 nbits)     // nbits <= 32, v < 2^nbits
    {
        acc |= (u64)v << fill;
        fill += nbits;
        if (fill >= 32) {
            if ((u32)acc) atomicOr(words + widx, (u32)acc);
            widx++; acc >>= 32; fill -= 32;
        }
    }
    __device__ __forceinline__ void finish() { if (fill && (u32)acc) atomicOr(words + widx, (u32)acc); }
};

__global__ void __launch_bounds__(256)
k_tta_pack(const u64* __restrict__ desc, u64 plane, const TtaFrame* __restrict__ frames, u32 C,
           const u32* __restrict__ frame_bytes, const u64* __restrict__ frame_off, uint8_t* __restrict__ out,
           const u64* __restrict__ total, u64 capacity, const u32* __restrict__ crc_tab, const u32* __restrict__ crc_pow)
{
    if (*total + 16 > capacity) return;            // does not fit: the host reports it
    __shared__ u32 tab[4][256];
    __shared__ u32 wtot[8];
    __shared__ u32 red[8];
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (u32 i = tid; i < 1024; i += 256) (&tab[0][0])[i] = crc_tab[i];
    const u32 f = blockIdx.x;
    const TtaFrame fr = frames[f];
    const u64 E = (u64)fr.n * C;
    const u64 off = frame_off[f];
    u32* words = (u32*)(out + (off & ~3ull));
    const u32 bit0 = (u32)(off & 3) * 8;
    const u64* dbase = desc + fr.ebase;
    u64 base = 0;
    for (u64 e0 = 0; e0 < E; e0 += 1024) {
        const u64 e = e0 + 4ull * tid;
        u64 d[4];
        u32 mybits = 0;
        {
            u64 i = e / C;
            u32 c = (u32)(e % C);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                d[j] = (e + j < E) ? dbase[(u64)c * plane + i] : 0ull;
                mybits += (u32)d[j];
                if (++c == C) { c = 0; i++; }
            }
        }
        // exclusive scan of the threads' bit counts (a chunk's total stays far below 2^32 unless a code is
        // pathologically long; 64-bit running base)
        u32 inc = mybits;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= (u32)o) inc += t;
        }
        __syncthreads();                      // (the previous chunk's wtot has been read)
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        u32 woff = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const u32 t = wtot[w]; woff += (u32)w < warp ? t : 0u; tot += t; }
        if (mybits) {
            TtaSink bs;
            bs.init(words, (u64)bit0 + base + woff + inc - mybits);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const u32 len = (u32)d[j];
                if (!len) continue;
                const u32 hi = (u32)(d[j] >> 32), k = hi >> 27, lsb = hi & 0x7FFFFFFu;
                u32 ones = len - 1u - k;
                while (ones >= 32) { bs.put(0xFFFFFFFFu, 32); ones -= 32; }     // (rare: a run of 32+ one-bits)
                if (ones) bs.put((1u << ones) - 1u, ones);
                bs.put(lsb << 1, k + 1);                                        // the stop bit, then k low bits
            }
            bs.finish();
        }
        base += tot;
    }
    // ---- CRC-32 of the frame's bytes (tta.c:252-255), read back from L2 where the ORs landed ----
    const u64 N = (base + 7) >> 3;
    if (N + 4 != frame_bytes[f]) __trap();        // the residual kernel's totals and the descriptors disagree
    __threadfence();
    __syncthreads();
    const u64 S = (N + TTA_SEG - 1) / TTA_SEG;                   // segment 0 is the short one
    const u64 first = N - (S - 1) * TTA_SEG;
    const u32 boff = bit0 >> 3;                                  // frame byte j is byte boff + j of `words`
    u32 acc = 0;
    for (u64 s = tid; s < S; s += 256) {
        const u64 a = s == 0 ? 0 : first + (s - 1) * TTA_SEG;
        const u64 b = s == 0 ? first : a + TTA_SEG;
        u32 crc = 0xFFFFFFFFu;
        u64 j = a;
        for (; j < b && ((boff + j) & 3); j++) {
            const u32 byte = (__ldcg(words + ((boff + j) >> 2)) >> (8 * ((boff + j) & 3))) & 0xFF;
            crc = tab[0][(crc ^ byte) & 0xFF] ^ (crc >> 8);
        }
        for (; j + 4 <= b; j += 4) {
            crc ^= __ldcg(words + ((boff + j) >> 2));
            crc = tab[3][crc & 0xFF] ^ tab[2][(crc >> 8) & 0xFF] ^ tab[1][(crc >> 16) & 0xFF] ^ tab[0][crc >> 24];
        }
        for (; j < b; j++) {
            const u32 byte = (__ldcg(words + ((boff + j) >> 2)) >> (8 * ((boff + j) & 3))) & 0xFF;
            crc = tab[0][(crc ^ byte) & 0xFF] ^ (crc >> 8);
        }
        crc ^= 0xFFFFFFFFu;
        const u64 after = S - 1 - s;                              // whole segments behind this one
        const u32 pw = after < TTA_POW_N ? crc_pow[after] : tta_xpow8(after * TTA_SEG);
        acc ^= after ? tta_multmodp(pw, crc) : crc;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) {
        u32 crc = 0;
        for (int w = 0; w < 8; w++) crc ^= red[w];
        TtaSink bs;
        bs.init(words, (u64)bit0 + N * 8);
        bs.put(crc, 32);
        bs.finish();
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static int check_params(const b200tta_params* p)
{
    if (!p) return tfail("params is NULL");
    if (p->channels < 1 || p->channels > 8) return tfail("unsupported channel count");
    if (p->bits_per_sample != 8 && p->bits_per_sample != 16 && p->bits_per_sample != 24) return tfail("bits_per_sample must be 8, 16 or 24");
    if (p->sample_rate == 0 || b200tta_block_size(p->sample_rate) == 0) return tfail("unsupported sample rate");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return tfail("no CUDA device available: the B200 TTA engine has no CPU fallback"); }
    return 0;
}

extern "C" uint64_t b200tta_output_bound(const b200tta_params* p, uint64_t n_pcm_frames, uint32_t n_frames)
{
    // two bits more than the sample width per sample is generous for anything but adversarial input (the size
    // is known exactly before packing, and checked against the capacity)
    return (n_pcm_frames * p->channels * (p->bits_per_sample + 2) + 7) / 8 + (u64)n_frames * 8 + 64;
}

static void crc_tables(std::vector<u32>& t)
{
    t.resize(1024 + TTA_POW_N);
    for (u32 i = 0; i < 256; i++) {
        u32 c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ TTA_POLY : c >> 1;
        t[i] = c;
    }
    for (u32 i = 0; i < 256; i++)
        for (u32 s = 1; s < 4; s++) t[256 * s + i] = t[t[256 * (s - 1) + i] & 0xFF] ^ (t[256 * (s - 1) + i] >> 8);
    const u32 step = tta_xpow8(TTA_SEG);
    u32 p = 1u << 31;
    for (u32 j = 0; j < TTA_POW_N; j++) { t[1024 + j] = p; p = tta_multmodp(step, p); }
}

static int build_frames(const b200tta_params* p, uint64_t n_pcm_frames, const uint32_t* lengths, uint32_t n_lengths,
                        std::vector<TtaFrame>& fr, u64* plane)
{
    const u32 block = b200tta_block_size(p->sample_rate);
    u64 pos = 0, eb = 0;
    if (lengths) {
        for (u32 i = 0; i < n_lengths; i++) {
            if (lengths[i] == 0) return tfail("a frame length is zero");
            TtaFrame f; f.pcm_off = pos; f.ebase = eb; f.n = lengths[i]; f.pad = 0;
            fr.push_back(f);
            pos += lengths[i];
            eb += ((u64)lengths[i] + 3) & ~3ull;
        }
        if (pos != n_pcm_frames) return tfail("frame lengths do not add up to the PCM frame count");
    } else {
        while (pos < n_pcm_frames) {
            const u32 n = (u32)((n_pcm_frames - pos) < block ? (n_pcm_frames - pos) : block);
            TtaFrame f; f.pcm_off = pos; f.ebase = eb; f.n = n; f.pad = 0;
            fr.push_back(f);
            pos += n;
            eb += ((u64)n + 3) & ~3ull;
        }
    }
    *plane = eb;
    return 0;
}

// d_out == NULL: the output is allocated here once its size is known and copied to *host_out (malloc'd)
static int encode_core(const b200tta_params* p, const uint8_t* d_pcm, uint64_t n_pcm_frames, const uint32_t* lengths,
                       uint32_t n_lengths, int device, uint8_t* d_out, uint64_t out_capacity, uint8_t** host_out,
                       uint64_t* out_bytes, std::vector<uint32_t>* sizes, float* kernel_ms)
{
    int rc = 0;
    std::vector<TtaFrame> fr;
    u64 plane = 0;
    if (build_frames(p, n_pcm_frames, lengths, n_lengths, fr, &plane)) return 1;
    const u32 nf = (u32)fr.size(), C = p->channels, B = p->bits_per_sample / 8;
    if (out_bytes) *out_bytes = 0;
    if (sizes) sizes->clear();
    if (nf == 0) { if (host_out) *host_out = (uint8_t*)malloc(1); return 0; }
    TtaFrame* d_fr = nullptr;
    u64 *d_desc = nullptr, *d_bits = nullptr, *d_off = nullptr, *d_total = nullptr;
    u32 *d_fbytes = nullptr, *d_tab = nullptr;
    uint8_t* d_own = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<u32> tab;
    u64 total = 0;
    TCK(cudaSetDevice(device));
    for (auto& e : ev) TCK(cudaEventCreate(&e));
    TCK(cudaMalloc((void**)&d_fr, nf * sizeof(TtaFrame)));
    TCK(cudaMalloc((void**)&d_desc, (size_t)plane * C * sizeof(u64) + 64));
    TCK(cudaMalloc((void**)&d_bits, (size_t)nf * C * sizeof(u64)));
    TCK(cudaMalloc((void**)&d_off, (size_t)nf * sizeof(u64)));
    TCK(cudaMalloc((void**)&d_fbytes, (size_t)nf * sizeof(u32)));
    TCK(cudaMalloc((void**)&d_total, 64));
    crc_tables(tab);
    TCK(cudaMalloc((void**)&d_tab, tab.size() * sizeof(u32)));
    TCK(cudaMemcpy(d_tab, tab.data(), tab.size() * sizeof(u32), cudaMemcpyHostToDevice));
    TCK(cudaMemcpy(d_fr, fr.data(), nf * sizeof(TtaFrame), cudaMemcpyHostToDevice));
    TCK(cudaEventRecord(ev[0]));
    if (B == 2) k_tta_residual<2><<<(nf * C + 31) / 32, 32>>>(d_pcm, d_fr, nf, C, p->bits_per_sample, d_desc, plane, d_bits);
    else if (B == 3) k_tta_residual<3><<<(nf * C + 31) / 32, 32>>>(d_pcm, d_fr, nf, C, p->bits_per_sample, d_desc, plane, d_bits);
    else k_tta_residual<1><<<(nf * C + 31) / 32, 32>>>(d_pcm, d_fr, nf, C, p->bits_per_sample, d_desc, plane, d_bits);
    TCK(cudaGetLastError());
    TCK(cudaEventRecord(ev[1]));
    k_tta_sizes<<<1, 1024>>>(d_bits, nf, C, d_fbytes, d_off, d_total);
    TCK(cudaGetLastError());
    TCK(cudaEventRecord(ev[2]));
    TCK(cudaMemcpy(&total, d_total, sizeof(u64), cudaMemcpyDeviceToHost));
    if (!d_out) {
        out_capacity = ((total + 16 + 15) & ~15ull) + 16;
        TCK(cudaMalloc((void**)&d_own, (size_t)out_capacity));
        d_out = d_own;
    } else if (total + 16 > out_capacity) {
        rc = tfail("encoded frames exceed the output buffer");
        goto done;
    }
    TCK(cudaEventRecord(ev[2]));
    k_tta_zero<<<148 * 4, 256>>>((uint4*)d_out, d_total, out_capacity & ~15ull);
    k_tta_pack<<<nf, 256>>>(d_desc, plane, d_fr, C, d_fbytes, d_off, d_out, d_total, out_capacity & ~15ull, d_tab, d_tab + 1024);
    TCK(cudaGetLastError());
    TCK(cudaEventRecord(ev[3]));
    TCK(cudaDeviceSynchronize());
    if (kernel_ms) {
        TCK(cudaEventElapsedTime(&kernel_ms[0], ev[0], ev[1]));
        TCK(cudaEventElapsedTime(&kernel_ms[1], ev[1], ev[2]));
        TCK(cudaEventElapsedTime(&kernel_ms[2], ev[2], ev[3]));
    }
    if (sizes) {
        sizes->resize(nf);
        TCK(cudaMemcpy(sizes->data(), d_fbytes, nf * sizeof(u32), cudaMemcpyDeviceToHost));
        for (u32 v : *sizes) if (v == 0xFFFFFFFFu) { rc = tfail("a frame exceeds 4 GiB"); goto done; }
    }
    if (host_out) {
        *host_out = (uint8_t*)malloc((size_t)total + 1);
        if (!*host_out) { rc = tfail("out of memory"); goto done; }
        TCK(cudaMemcpy(*host_out, d_out, (size_t)total, cudaMemcpyDeviceToHost));
    }
    if (out_bytes) *out_bytes = total;
done:
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    cudaFree(d_fr); cudaFree(d_desc); cudaFree(d_bits); cudaFree(d_off); cudaFree(d_fbytes); cudaFree(d_total);
    cudaFree(d_tab); cudaFree(d_own);
    return rc;
}

extern "C" int b200tta_encode_device(const b200tta_params* params, const void* d_pcm, uint64_t n_pcm_frames, int device,
                                     void* d_out, uint64_t out_capacity, uint64_t* out_bytes, uint32_t* frame_sizes,
                                     uint32_t* n_frames, float* kernel_ms)
{
    if (check_params(params)) return 1;
    if (!d_pcm || !d_out || ((uintptr_t)d_out & 15)) return tfail("d_pcm/d_out is NULL or d_out is not 16-byte aligned");
    std::vector<uint32_t> sizes;
    if (encode_core(params, (const uint8_t*)d_pcm, n_pcm_frames, nullptr, 0, device, (uint8_t*)d_out, out_capacity, nullptr,
                    out_bytes, &sizes, kernel_ms)) return 1;
    if (frame_sizes) memcpy(frame_sizes, sizes.data(), sizes.size() * sizeof(uint32_t));
    if (n_frames) *n_frames = (uint32_t)sizes.size();
    return 0;
}

extern "C" int b200tta_encode_frames(const b200tta_params* params, const uint8_t* pcm, uint64_t n_pcm_frames,
                                     const uint32_t* frame_lengths, uint32_t n_lengths, int device,
                                     uint8_t** out, uint64_t* out_bytes, uint32_t** frame_sizes, uint32_t* n_frames,
                                     float* kernel_ms)
{
    if (check_params(params)) return 1;
    if (!out || (!pcm && n_pcm_frames)) return tfail("pcm/out is NULL");
    int rc = 0;
    uint8_t* d_pcm = nullptr;
    std::vector<uint32_t> sizes;
    const size_t nbytes = (size_t)n_pcm_frames * params->channels * (params->bits_per_sample / 8);
    *out = nullptr;
    TCK(cudaSetDevice(device));
    TCK(cudaMalloc((void**)&d_pcm, nbytes + 64));
    if (nbytes) TCK(cudaMemcpy(d_pcm, pcm, nbytes, cudaMemcpyHostToDevice));
    rc = encode_core(params, d_pcm, n_pcm_frames, frame_lengths, n_lengths, device, nullptr, 0, out, out_bytes, &sizes, kernel_ms);
    if (!rc) {
        if (frame_sizes) {
            *frame_sizes = (uint32_t*)malloc((sizes.size() ? sizes.size() : 1) * sizeof(uint32_t));
            memcpy(*frame_sizes, sizes.data(), sizes.size() * sizeof(uint32_t));
        }
        if (n_frames) *n_frames = (uint32_t)sizes.size();
    }
done:
    cudaFree(d_pcm);
    return rc;
}

static void put_le(uint8_t* p, uint32_t v, int bytes) { for (int i = 0; i < bytes; i++) p[i] = (uint8_t)(v >> (8 * i)); }

static uint32_t host_crc32(const uint8_t* p, size_t n)
{
    uint32_t crc = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) {
        crc ^= p[i];
        for (int k = 0; k < 8; k++) crc = (crc & 1) ? (crc >> 1) ^ TTA_POLY : crc >> 1;
    }
    return crc ^ 0xFFFFFFFFu;
}

extern "C" int b200tta_encode_file(const char* filename, const b200tta_params* params, const uint8_t* pcm,
                                   uint64_t n_pcm_frames, int device)
{
    if (!filename) return tfail("filename is NULL");
    if (check_params(params)) return 1;
    if (n_pcm_frames > 0xFFFFFFFFull) return tfail("too many PCM frames for a TTA header");
    uint8_t* frames = nullptr;
    uint32_t* sizes = nullptr;
    uint64_t nbytes = 0;
    uint32_t nf = 0;
    if (b200tta_encode_frames(params, pcm, n_pcm_frames, nullptr, 0, device, &frames, &nbytes, &sizes, &nf, nullptr)) return 1;
    // write_header (tta.c:562-580) and write_seektable (:582-595)
    std::vector<uint8_t> head(22 + 4 * (size_t)nf + 4);
    memcpy(head.data(), "TTA1", 4);
    put_le(&head[4], 1, 2); put_le(&head[6], params->channels, 2); put_le(&head[8], params->bits_per_sample, 2);
    put_le(&head[10], params->sample_rate, 4); put_le(&head[14], (uint32_t)n_pcm_frames, 4);
    put_le(&head[18], host_crc32(head.data(), 18), 4);
    for (uint32_t i = 0; i < nf; i++) put_le(&head[22 + 4 * (size_t)i], sizes[i], 4);
    put_le(&head[22 + 4 * (size_t)nf], host_crc32(&head[22], 4 * (size_t)nf), 4);
    int rc = 0;
    FILE* f = fopen(filename, "wb");
    if (!f) {
        snprintf(t_err, sizeof(t_err), "cannot open \"%.300s\" for writing", filename);
        rc = 1;
    } else {
        if (fwrite(head.data(), 1, head.size(), f) != head.size() || (nbytes && fwrite(frames, 1, (size_t)nbytes, f) != nbytes)) rc = tfail("write error");
        if (fclose(f) != 0 && !rc) rc = tfail("write error");
    }
    free(frames);
    free(sizes);
    return rc;
}
