// flac_common.cuh -- device helpers shared by the FLAC kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "flac_types.h"

typedef unsigned long long u64;
typedef unsigned int u32;

// shared-memory sample arrays are skewed by one word per 32 so that threads
// owning contiguous 32-sample chunks hit 32 different banks
#define PADI(i) ((i) + ((i) >> 5))

// ---------------------------------------------------------------------------
// PCM access.  pcm is interleaved little-endian signed, B bytes per sample
// (what pcmreader hands the reference's MD5 callback, flac.c:188).
// The de-interleave of src/pcmconv.c:254-263 and the average/difference
// signals of flacenc_average_difference (flac.c:1507-1529) are fused into
// this loader: no planar copy of the PCM is ever written to HBM.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int ld_pcm(const uint8_t* __restrict__ pcm, u64 sidx, u32 B)
{
    const uint8_t* p = pcm + sidx * B;
    if (B == 2) return (int)(*(const short*)p);
    if (B == 3) {
        // aligned 32-bit loads and a funnel shift instead of three byte loads; the second word is only
        // touched when the sample straddles it, so nothing past the last sample's word is read
        const u32* w = (const u32*)((uintptr_t)p & ~(uintptr_t)3);
        const u32 sh = ((u32)(uintptr_t)p & 3u) * 8u;
        const u32 v = __funnelshift_r(w[0], sh > 8u ? w[1] : 0u, sh);
        return ((int)(v << 8)) >> 8;
    }
    if (B == 1) return (int)(*(const signed char*)p);
    return *(const int*)p;
}

// 24-bit little-endian sample whose first byte is byte `sh` (0..5) of the word pair (lo, hi), sign extended: ONE
// PRMT -- the fourth selector nibble has bit 3 set, which replicates the sign of the selected (top) byte.
// (prmt.b32 in PTX; __byte_perm documents only three selector bits per nibble.)
__device__ __forceinline__ int s24_from_words(u32 lo, u32 hi, int sh)
{
    int r;
    const u32 sel = 0xA210u + 0x1111u * (u32)sh;      // nibbles sh, sh+1, sh+2, 8|(sh+2): sh <= 5
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(lo), "r"(hi), "r"(sel));
    return r;
}

// candidate `cand` of PCM frame `t` (absolute index in the batch's pcm buffer)
__device__ __forceinline__ int ld_candidate(const uint8_t* __restrict__ pcm, u64 t, u32 cand,
                                            const bf_dev_params& P)
{
    if (!P.stereo) return ld_pcm(pcm, t * P.channels + cand, P.bytes_ps);
    if (P.bytes_ps == 2) {
        // one aligned 32-bit load fetches the left/right pair
        const u32 lr = *(const u32*)(pcm + t * 4);
        const int L = (int)(short)(lr & 0xFFFF), R = (int)(short)(lr >> 16);
        return cand == 0 ? L : cand == 1 ? R : cand == 2 ? ((L + R) >> 1) : (L - R);
    }
    const int L = ld_pcm(pcm, t * 2, P.bytes_ps), R = ld_pcm(pcm, t * 2 + 1, P.bytes_ps);
    return cand == 0 ? L : cand == 1 ? R : cand == 2 ? ((L + R) >> 1) : (L - R);
}

// 24-bit stereo block -> candidate samples in shared memory (dst indexed through the 4-words-per-32 skew of the v3
// kernels): eight PCM frames are 48 bytes, three 128-bit loads; one PRMT per sample picks its three bytes and
// replicates the sign of the top one.  (The generic per-sample loop is one dependent load per sample: a tenth of
// the analysis kernel's time at 24 bits.)  src 16-byte aligned, n a multiple of 8.  Out of line on purpose: the
// 16-bit kernels' hot code must not grow by it (instruction cache, see profiles/r02_icache_analysis.txt).
__device__ __noinline__ void load_stereo24_skewed(const uint8_t* __restrict__ src, int* __restrict__ dst, u32 n, u32 cand,
                                                  u32 tid, u32 nt, int first, u32* orv_io, u32* diff_io)
{
    const uint4* s4 = (const uint4*)src;
    u32 orv = 0, diff = 0;
#pragma unroll 2
    for (u32 g = tid; g < (n >> 3); g += nt) {
        const uint4 a = __ldg(s4 + 3 * g), b = __ldg(s4 + 3 * g + 1), c = __ldg(s4 + 3 * g + 2);
        const u32 w[13] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, 0u};
        int v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int oL = 6 * j, oR = 6 * j + 3;
            const int L = s24_from_words(w[oL >> 2], w[(oL >> 2) + 1], oL & 3);
            const int R = s24_from_words(w[oR >> 2], w[(oR >> 2) + 1], oR & 3);
            v[j] = cand == 0 ? L : cand == 1 ? R : cand == 2 ? ((L + R) >> 1) : (L - R);
            orv |= (u32)v[j]; diff |= (u32)(v[j] ^ first);
        }
        const u32 i = 8 * g, sk = i + ((i >> 5) << 2);
        *(int4*)(dst + sk) = make_int4(v[0], v[1], v[2], v[3]);
        *(int4*)(dst + sk + 4) = make_int4(v[4], v[5], v[6], v[7]);
    }
    *orv_io |= orv; *diff_io |= diff;
}

// bits-per-sample a candidate is coded at: the difference channel needs one more (flac.c:569)
__device__ __forceinline__ u32 candidate_bps(u32 cand, const bf_dev_params& P)
{
    return P.bps + ((P.stereo && cand == 3) ? 1u : 0u);
}

// ---------------------------------------------------------------------------
// block-wide reductions (blockDim.x multiple of 32, <= 1024); `red` is shared
// scratch of at least 33 u64.  All return the result to every thread.
// ---------------------------------------------------------------------------
__device__ __forceinline__ u64 block_sum_u64(u64 v, u64* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? red[lane] : 0ull;
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
        if (lane == 0) red[32] = v;
    }
    __syncthreads();
    return red[32];
}

__device__ __forceinline__ u32 block_or_u32(u32 v, u64* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = __reduce_or_sync(0xFFFFFFFFu, v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? (u32)red[lane] : 0u;
        v = __reduce_or_sync(0xFFFFFFFFu, v);
        if (lane == 0) red[32] = v;
    }
    __syncthreads();
    return (u32)red[32];
}

__device__ __forceinline__ u32 block_max_u32(u32 v, u64* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = __reduce_max_sync(0xFFFFFFFFu, v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nw ? (u32)red[lane] : 0u;
        v = __reduce_max_sync(0xFFFFFFFFu, v);
        if (lane == 0) red[32] = v;
    }
    __syncthreads();
    return (u32)red[32];
}

// exclusive prefix sum of one u32 per thread; also returns the block total
__device__ __forceinline__ u32 block_exscan_u32(u32 v, u64* red, u32* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) red[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < nw ? (u32)red[lane] : 0u, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
            if (lane >= o) winc += t;
        }
        red[lane] = winc - w;          // exclusive offset of each warp
        if (lane == 31) red[32] = winc;
    }
    __syncthreads();
    const u32 r = (u32)red[warp] + inc - v;
    *total = (u32)red[32];
    return r;
}

// ---------------------------------------------------------------------------
// residual helpers
// ---------------------------------------------------------------------------
// zig-zag fold of flac.c:1424-1428
// int -> double without the (quarter-rate) I2F.F64: 2^52 + 2^31 + v is exact for any 32-bit v,
// and so is the subtraction that follows
__device__ __forceinline__ double int2double_exact(int v)
{
    return __dsub_rn(__hiloint2double(0x43300000, (int)((u32)v ^ 0x80000000u)), 4503601774854144.0);
}

// acc + a * b as a 32 x 32 -> 64-bit multiply-add: ONE instruction (IMAD.WIDE).  Inline PTX on purpose: given
// (long long)a * (long long)b with both operands live across an unrolled loop, the compiler keeps them sign-extended
// and emits a 64 x 64-bit multiply per tap (IMAD.WIDE.U32 + 2 IMAD + IADD3 -- measured: 2.9 instructions per tap
// in the 24-bit residual loops).
__device__ __forceinline__ long long mad_wide(int a, int b, long long acc)
{
    long long r;
    asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(acc));
    return r;
}

__device__ __forceinline__ u32 zigzag(int r)
{
    return ((u32)r << 1) ^ (u32)(r >> 31);
}

// ---------------------------------------------------------------------------
// MSB-first bit writer into a zero-initialised output buffer viewed as
// big-endian 32-bit words.  Threads own disjoint bit ranges; words shared by
// two ranges are merged with atomicOr, so no ordering between threads is
// needed.  Equivalent of BitstreamWriter.write/write_unary
// (src/bitstream.c:1905-1947, 2233-2251) with the bit position known up front.
// ---------------------------------------------------------------------------
struct BitSink {
    u32* words;     // 4-byte aligned base of the output buffer
    u64 widx;       // next word to emit
    u64 acc;        // pending bits, left-aligned
    u32 fill;       // number of pending bits (< 32)

    __device__ __forceinline__ void init(u32* base, u64 bitpos)
    {
        words = base; widx = bitpos >> 5; fill = (u32)(bitpos & 31); acc = 0;
    }
    __device__ __forceinline__ void emit(u32 w)
    {
        if (w) atomicOr(words + widx, __byte_perm(w, 0, 0x0123));
        widx++;
    }
    // append the low nbits of v (0 <= nbits <= 32; v < 2^nbits)
    __device__ __forceinline__ void put(u32 v, u32 nbits)
    {
        if (nbits == 0) return;
        acc |= (u64)v << (64 - fill - nbits);
        fill += nbits;
        if (fill >= 32) { emit((u32)(acc >> 32)); acc <<= 32; fill -= 32; }
    }
    // two's complement in nbits (write_signed, src/bitstream.c:2020-2034)
    __device__ __forceinline__ void put_signed(int v, u32 nbits)
    {
        put(nbits >= 32 ? (u32)v : ((u32)v & ((1u << nbits) - 1u)), nbits);
    }
    // append n zero bits (n may be huge: unary runs)
    __device__ __forceinline__ void zeros(u64 n)
    {
        const u64 tot = (u64)fill + n;
        if (tot >= 32) {
            const u32 top = (u32)(acc >> 32);
            if (top) atomicOr(words + widx, __byte_perm(top, 0, 0x0123));
            widx += tot >> 5;
            acc = 0;
            fill = (u32)(tot & 31);
        } else {
            fill = (u32)tot;
        }
    }
    __device__ __forceinline__ void flush()
    {
        if (fill) { const u32 top = (u32)(acc >> 32); if (top) atomicOr(words + widx, __byte_perm(top, 0, 0x0123)); }
    }
    __device__ __forceinline__ u64 bitpos() const { return (widx << 5) + fill; }
};

// ---------------------------------------------------------------------------
// CRCs of src/common/flac_crc.c (MSB-first, init 0)
// ---------------------------------------------------------------------------
__device__ __forceinline__ u32 crc8_byte(u32 crc, u32 byte)
{
    crc ^= byte;
#pragma unroll
    for (int k = 0; k < 8; k++) crc = (crc & 0x80) ? ((crc << 1) ^ 0x07) & 0xFF : (crc << 1) & 0xFF;
    return crc;
}

__device__ __forceinline__ u32 crc16_byte(u32 crc, u32 byte)
{
    crc ^= byte << 8;
#pragma unroll
    for (int k = 0; k < 8; k++) crc = (crc & 0x8000) ? ((crc << 1) ^ 0x8005) & 0xFFFF : (crc << 1) & 0xFFFF;
    return crc;
}

// a * b mod (x^16 + x^15 + x^2 + 1) over GF(2)
__device__ __forceinline__ u32 gf16_mul(u32 a, u32 b)
{
    u32 r = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) r ^= ((b >> i) & 1u) ? (a << i) : 0u;
#pragma unroll
    for (int i = 30; i >= 16; i--) r ^= ((r >> i) & 1u) ? (0x18005u << (i - 16)) : 0u;
    return r;
}

// x^(8*nbytes) mod P
__device__ __forceinline__ u32 gf16_xpow8(u64 nbytes)
{
    u32 result = 1, base = 0x100; // x^8
    while (nbytes) {
        if (nbytes & 1) result = gf16_mul(result, base);
        base = gf16_mul(base, base);
        nbytes >>= 1;
    }
    return result;
}
