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

// 24-bit little-endian sample whose first byte is byte `sh` (0..3) of the word pair (lo, hi), sign extended: ONE
// PRMT -- the fourth selector nibble has bit 3 set, which replicates the sign of the selected (top) byte.
// (prmt.b32 in PTX; __byte_perm documents only three selector bits per nibble.)
__device__ __forceinline__ int s24_from_words(u32 lo, u32 hi, int sh)
{
    int r;
    const u32 sel = (u32)(sh | ((sh + 1) << 4) | ((sh + 2) << 8) | ((8 | (sh + 2)) << 12));
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
