// k_synth.cuh -- deterministic integer-only synthetic PCM, generated
// directly in HBM as packed little-endian samples.  The test suite carries a plain-C twin of
// these formulas and checks that both produce the same bytes (tests/test_gpu_parity.py).
// Benchmark/test input only -- not part of the encode path.
#pragma once
#include "flac_common.cuh"

__device__ __forceinline__ u64 synth_mix64(u64 z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__device__ __forceinline__ int synth_psin(u32 phase)
{
    const int x = (int)(phase >> 16) - 32768;
    const int ax = x < 0 ? -x : x;
    return (x * (32768 - ax)) >> 14;
}

__device__ int synth_sample(u64 seed, u32 c, u32 bps, u64 t)
{
    const int amp[4] = {12000, 8000, 5000, 3000};
    const int share_tab[8] = {3, 3, 4, 1, 1, 1, 2, 2};
    long long shared_v = 0, own = 0, tot;
    const bool lfe = (c == 3);
#pragma unroll
    for (u32 v = 0; v < 4; v++) {
        u32 inc = 5000000u + (u32)(synth_mix64(seed * 16 + v) % 400000000u);
        const u32 ph0 = (u32)synth_mix64(seed * 16 + 8 + v);
        if (lfe) inc >>= 5;
        shared_v += (long long)amp[v] * synth_psin(inc * (u32)t + ph0);
    }
#pragma unroll
    for (u32 v = 0; v < 2; v++) {
        u32 inc = 4000000u + (u32)(synth_mix64(seed * 16 + 64 + c * 4 + v) % 300000000u);
        const u32 ph0 = (u32)synth_mix64(seed * 16 + 128 + c * 4 + v);
        if (lfe) inc >>= 5;
        own += (long long)amp[v] * 2 * synth_psin(inc * (u32)t + ph0);
    }
    const int sh = share_tab[c & 7];
    tot = (shared_v * sh + own * (4 - sh)) >> 16;
    const int env = 16384 + synth_psin(9000u * (u32)t + (u32)(seed * 2654435761ull));
    tot = (tot * env) >> 16;
    const u64 h = synth_mix64(seed ^ ((u64)c << 56) ^ t);
    long long v;
    if (bps >= 16) {
        const int noise = (int)(h & ((1u << (bps - 6)) - 1)) - (int)(1u << (bps - 7));
        v = tot * (1ll << (bps - 16)) + noise;
    } else {
        const int noise = (int)(h & 0x3FF) - 512;
        v = (tot + noise) >> (16 - bps);
    }
    const long long lo = -(1ll << (bps - 1)), hi = (1ll << (bps - 1)) - 1;
    v = v < lo ? lo : (v > hi ? hi : v);
    return (int)v;
}

__global__ void k_synth_pcm(uint8_t* __restrict__ pcm, u64 seed, u32 channels, u32 bps, u64 first_frame, u64 n_frames)
{
    const u32 B = bps / 8;
    const u64 total = n_frames * channels;
    for (u64 idx = (u64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (u64)gridDim.x * blockDim.x) {
        const u64 t = idx / channels;
        const u32 c = (u32)(idx % channels);
        const int v = synth_sample(seed, c, bps, first_frame + t);
        uint8_t* p = pcm + idx * B;
        for (u32 b = 0; b < B; b++) p[b] = (uint8_t)((u32)v >> (8 * b));
    }
}
