// k_analyze_v2.cuh -- the model-search kernel used for every block that fits shared memory
// and one pass of <= 512 threads (anything else falls back to k_analyze.cuh).
//
// Same decisions, bit for bit, as k_analyze (and the reference, flac.c:673-1505).  One CTA per
// unit (frame, candidate).  What differs is how the work is laid out for the SM:
//   * the block's samples live in shared memory (skewed one word per 32, so a thread's contiguous
//     run is bank-conflict free); PCM is fetched with coalesced loads and the candidate signal
//     (L, R, (L+R)>>1, L-R) is formed on the way in;
//   * every thread owns a contiguous run of S = 8*m samples and walks it in CHUNKS OF 8 with
//     rolled loops: each loop body (8 samples, fully unrolled, predictor history in registers)
//     is a few hundred instructions, so the hot code stays inside the instruction cache --
//     the fully unrolled predecessor of this kernel was bound by instruction fetch;
//   * residuals are written to shared memory once and re-read for the exact bit count;
//   * the partition-order decision is data-parallel (prefix sums + closed-form Rice parameter),
//     with warp reductions instead of 64-bit shared atomics.
#pragma once
#include "flac_common.cuh"
#include "k_analyze.cuh"

// Rice parameter + estimate of one partition, same results as partition_estimate(): closed form when the 32-bit
// wrap of the reference's shift (H3) cannot change the outcome, the loop itself otherwise.
// With kc the smallest k such that plength * 2^k >= S_ in exact arithmetic: for k < kc the wrapped value is at
// most the exact one, hence still below S_ and the loop goes on; at kc the exact value is below 2 * S_ (kc is
// minimal), so for S_ < 2^31 it has not wrapped and the loop stops there -- whatever max_rice is.  (The earlier
// test, "no shift up to max_rice wraps", sent every 24-bit partition of four samples or more through the loop:
// max_rice is 30 there.)
__device__ __forceinline__ u64 partition_estimate_fast(u32 plength, u64 S_, u32 max_rice, u32* k_out)
{
    if (plength == 0 || S_ >= (1ull << 31)) return partition_estimate(plength, S_, max_rice, k_out);
    // from here on everything fits 32 bits: S < 2^31, and the estimate is at most 4 + 2 S + 32 plength
    const u32 s32 = (u32)S_;
    u32 k = 0;
    if (s32 > plength) {
        // smallest k with (plength << k) >= S: with a, b the bit lengths of S and plength (a >= b here),
        // plength << (a - b - 1) < 2^(a-1) <= S and plength << (a - b + 1) >= 2^a > S, so it is a - b or a - b + 1
        u32 kc = (u32)(__clz((int)plength) - __clz((int)s32));
        if ((plength << kc) < s32) kc++;
        k = kc < max_rice ? kc : max_rice;
    }
    *k_out = k;
    // k > 0: 4 + (S >> (k - 1)) + (1 + k) plength - plength / 2;  k == 0: 4 + 2 S + plength - plength / 2
    // (S >> (k - 1) may exceed 32 bits only when k was clamped to max_rice with a huge S: widened below)
    const u32 tail = 4u + (1u + k) * plength - (plength >> 1);
    return (u64)tail + (k ? (u64)(s32 >> (k - 1)) : ((u64)s32 << 1));
}

// shared scratch of the Rice search (double buffered by search parity, see rice_search_v2)
struct RiceScratch {
    u64 lvl_total[2][16];
    u64 bits_total[2];
    u32 lvl_maxk[2][16];
};


#define V2_CH 8
// address of the chunk starting at sample i0 (i0 % 8 == 0): the 8 samples are contiguous words,
// because the one-word skew per 32 samples never falls inside a chunk
#define CHP(arr, i0) ((arr) + ((i0) + ((i0) >> 5)))

struct V2Ctx {
    u32 n;            // samples in the block
    u32 S;            // samples per thread (multiple of 8)
    u32 base;         // first sample of this thread's run
    int* samp;        // [PADI(n)] wasted-shifted samples
    int* resid;       // [PADI(n)] residual of the model under test (by sample index)
    u64* fine;        // finest partition sums / prefix sums
    uint8_t* karr;    // Rice parameters, heap order (node-1)
    RiceScratch* rs;
};

// Rice search over c.resid (valid for sample indices order..n-1; entries below `order` may hold
// anything, entries >= n are never read).  All threads of the CTA call it.
__device__ __noinline__ void rice_search_v2(const V2Ctx& c, u32 parity, u32 order, const bf_dev_params& P,
                                            RiceChoice* out)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const u32 n = c.n, S = c.S, base = c.base;
    const u32 po_eff = min(P.po_lim, (u32)(__ffs((int)n) - 1));
    u32 F = po_eff;
    while (F > 0 && (n >> F) < order) F--;
    const u32 nfine = 1u << F;
    const u32 plenF = n >> F;
    const u32 lo = max(base, order), hi = min(base + S, n);
    u64* fine = c.fine;
    u64* lvl_total = c.rs->lvl_total[parity];
    u32* lvl_maxk = c.rs->lvl_maxk[parity];

    // how the per-thread sums reach the finest partitions (uniform per CTA):
    //   grouped: a partition is g whole runs, g a power of two <= 32  -> warp shuffles + plain store
    //   owned  : a run is whole partitions and partitions are whole chunks -> plain stores
    //   else   : shared atomics into a zeroed array
    const u32 g = plenF / S;
    const bool grouped = (plenF >= S) && (plenF % S == 0) && ((g & (g - 1)) == 0) && g <= 32;
    const bool owned = (plenF < S) && (S % plenF == 0) && (plenF % V2_CH == 0);
    if (!grouped && !owned) for (u32 i = tid; i < nfine; i += nt) fine[i] = 0ull;
    if (tid < 16) { lvl_total[tid] = 0ull; lvl_maxk[tid] = 0u; }
    if (tid == 0) c.rs->bits_total[parity] = 0ull;
    if (!grouped && !owned) __syncthreads();

    // ---- pass 1: sum |r| per finest partition ----
    if (grouped) {
        u64 run = 0;
        for (u32 i0 = base; i0 < base + S; i0 += V2_CH) {
            const int* rp = CHP(c.resid, i0);
            if (i0 >= order && i0 + V2_CH <= n) {
#pragma unroll
                for (int j = 0; j < V2_CH; j++) run += (u64)(u32)abs(rp[j]);
            } else {
#pragma unroll
                for (int j = 0; j < V2_CH; j++) {
                    const u32 i = i0 + j;
                    if (i >= order && i < n) run += (u64)(u32)abs(rp[j]);
                }
            }
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_xor_sync(0xFFFFFFFFu, run, o);
            if ((u32)o < g) run += t;
        }
        const u32 p = (u32)tid / g;
        if (((u32)tid & (g - 1)) == 0 && p < nfine) fine[p] = run;
    } else if (owned) {
        u64 run = 0;
        u32 p = base / plenF;
        u32 left = plenF;
        for (u32 i0 = base; i0 < base + S; i0 += V2_CH) {
            const int* rp = CHP(c.resid, i0);
            if (i0 >= order && i0 + V2_CH <= n) {
#pragma unroll
                for (int j = 0; j < V2_CH; j++) run += (u64)(u32)abs(rp[j]);
            } else {
#pragma unroll
                for (int j = 0; j < V2_CH; j++) {
                    const u32 i = i0 + j;
                    if (i >= order && i < n) run += (u64)(u32)abs(rp[j]);
                }
            }
            left -= V2_CH;
            if (left == 0) { if (p < nfine) fine[p] = run; run = 0; p++; left = plenF; }
        }
    } else if (lo < hi) {
        u32 p = lo / plenF;
        u32 next = (p + 1) * plenF;
        u64 run = 0;
        for (u32 i = lo; i < hi; i++) {
            if (i == next) { atomicAdd(&fine[p], run); run = 0; p++; next += plenF; }
            run += (u64)(u32)abs(c.resid[PADI(i)]);
        }
        atomicAdd(&fine[p], run);
    }
    __syncthreads();

    // ---- inclusive prefix sum of the finest sums (warp 0) ----
    if (tid < 32) {
        const u32 per = (nfine + 31) >> 5;
        const u32 b = lane * per;
        u64 run = 0;
        for (u32 i = 0; i < per; i++) if (b + i < nfine) { run += fine[b + i]; fine[b + i] = run; }
        u64 inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += t;
        }
        const u64 excl = inc - run;
        for (u32 i = 0; i < per; i++) if (b + i < nfine) fine[b + i] += excl;
    }
    __syncthreads();

    // ---- one thread per (level, partition), heap numbering: thread t takes node t ----
    const u32 heapn = 2 * nfine - 1;
    for (u32 node0 = 0; node0 <= heapn; node0 += nt) {
        const u32 node = node0 + tid;
        const bool act = node >= 1 && node <= heapn;
        u32 l = 0, k = 0;
        u64 est = 0;
        if (act) {
            l = 31u - (u32)__clz((int)node);
            const u32 p = node - (1u << l);
            const u32 w = nfine >> l;
            const u64 hi_sum = fine[(p + 1) * w - 1];
            const u64 lo_sum = p ? fine[p * w - 1] : 0ull;
            const u32 plength = (n >> l) - (p == 0 ? order : 0u);
            est = partition_estimate_fast(plength, hi_sum - lo_sum, P.max_rice, &k);
            c.karr[node - 1] = (uint8_t)k;
        }
        if (node0 + (u32)(tid & ~31) >= 32) {
            const u32 km = __reduce_max_sync(0xFFFFFFFFu, k);
#pragma unroll
            for (int o = 16; o; o >>= 1) est += __shfl_xor_sync(0xFFFFFFFFu, est, o);
            const u32 l0 = __shfl_sync(0xFFFFFFFFu, l, 0);
            if (lane == 0 && node <= heapn) { atomicAdd(&lvl_total[l0], est); atomicMax(&lvl_maxk[l0], km); }
        } else {
            const u32 gs = act ? (1u << l) : 1u;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const u64 te = __shfl_xor_sync(0xFFFFFFFFu, est, o);
                const u32 tk = __shfl_xor_sync(0xFFFFFFFFu, k, o);
                if ((u32)o < gs) { est += te; k = max(k, tk); }
            }
            if (act && node == gs) { lvl_total[l] = est; lvl_maxk[l] = k; }
        }
    }
    __syncthreads();

    // ---- first strict minimum over the levels (every thread, redundantly) ----
    u64 best = lvl_total[0];
    u32 po = 0, k0 = 0;
    for (u32 l = 1; l <= F; l++) {
        const u64 tot = lvl_total[l];
        if (tot < best) { best = tot; po = l; }
    }
    for (u32 l = F + 1; l <= po_eff; l++) { // underflow levels (rare): partition 0 swallows every residual (H3)
        u32 kk = 0;
        const u32 pl = n >> l;
        u64 tot = partition_estimate(pl - order, fine[nfine - 1], P.max_rice, &kk);
        tot += (u64)((1u << l) - 1u) * (4ull + (u64)pl - (u64)(pl / 2));
        if (tot < best) { best = tot; po = l; k0 = kk; }
    }
    const u32 under = po > F ? 1u : 0u;
    const u32 koff = (1u << po) - 1u;
    const u32 plen = n >> po;
    const u32 maxk = under ? k0 : lvl_maxk[po];

    // ---- pass 2: exact size ----
    u64 bits = 0;
    if (lo < hi) {
        if (under || (plen % V2_CH) == 0) {
            // a chunk never straddles a partition boundary
            u32 i0 = lo & ~(V2_CH - 1u);
            u32 p = under ? 0u : i0 / plen;
            u32 next = under ? 0xFFFFFFFFu : (p + 1) * plen;
            u32 k = under ? k0 : (u32)c.karr[koff + p];
            for (; i0 < hi; i0 += V2_CH) {
                if (i0 == next) { p++; next += plen; k = c.karr[koff + p]; }
                const int* rp = CHP(c.resid, i0);
                if (i0 >= lo && i0 + V2_CH <= hi) {
                    if (k >= 3) {
                        u32 a = 0; // 8 * (2^32 >> 3) < 2^32
#pragma unroll
                        for (int j = 0; j < V2_CH; j++) a += zigzag(rp[j]) >> k;
                        bits += (u64)a + (u64)V2_CH * (1u + k);
                    } else {
                        u64 a = 0;
#pragma unroll
                        for (int j = 0; j < V2_CH; j++) a += (u64)(zigzag(rp[j]) >> k);
                        bits += a + (u64)V2_CH * (1u + k);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < V2_CH; j++) {
                        const u32 i = i0 + j;
                        if (i >= lo && i < hi) bits += (u64)(zigzag(rp[j]) >> k) + 1u + k;
                    }
                }
            }
        } else {
            u32 p = lo / plen;
            u32 next = (p + 1) * plen;
            u32 k = c.karr[koff + p];
            for (u32 i = lo; i < hi; i++) {
                if (i == next) { p++; next += plen; k = c.karr[koff + p]; }
                bits += (u64)(zigzag(c.resid[PADI(i)]) >> k) + 1u + k;
            }
        }
    }
    if (!__any_sync(0xFFFFFFFFu, (bits >> 26) != 0)) {
        const u32 wsum = __reduce_add_sync(0xFFFFFFFFu, (u32)bits);
        if (lane == 0) atomicAdd(&c.rs->bits_total[parity], (u64)wsum);
    } else {
#pragma unroll
        for (int o = 16; o; o >>= 1) bits += __shfl_xor_sync(0xFFFFFFFFu, bits, o);
        if (lane == 0) atomicAdd(&c.rs->bits_total[parity], bits);
    }
    __syncthreads();
    out->po = po; out->under = under; out->k0 = k0;
    out->method = maxk > 14 ? 1u : 0u;
    out->bits = c.rs->bits_total[parity] + 6ull + (u64)(1u << po) * (maxk > 14 ? 5ull : 4ull);
}

// coalesced PCM load of the unit into shared memory; returns OR of samples / OR of (sample ^ first)
__device__ __forceinline__ void load_unit_v2(const V2Ctx& c, const uint8_t* __restrict__ pcm, u64 pcm_off, u32 cand,
                                             const bf_dev_params& P, u32* or_out, u32* diff_out)
{
    const u32 tid = threadIdx.x, nt = blockDim.x, n = c.n;
    const int first = ld_candidate(pcm, pcm_off, cand, P);
    u32 orv = 0, diff = 0;
    if (P.stereo && P.bytes_ps == 2) {
        // candidate = (cl*L + cr*R) >> sh with (cl, cr, sh) = (1,0,0) (0,1,0) (1,1,1) (1,-1,0)
        const int cl = cand == 1 ? 0 : 1, cr = cand == 0 ? 0 : (cand == 3 ? -1 : 1), sh = cand == 2 ? 1 : 0;
        const u32* src = (const u32*)(pcm + pcm_off * 4);
        u32 i = tid;
        for (; i + 3 * nt < n; i += 4 * nt) {
            u32 w[4];
#pragma unroll
            for (int q = 0; q < 4; q++) w[q] = __ldg(src + i + q * nt);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int v = (cl * (int)(short)(w[q] & 0xFFFF) + cr * ((int)w[q] >> 16)) >> sh;
                c.samp[PADI(i + q * nt)] = v;
                orv |= (u32)v; diff |= (u32)(v ^ first);
            }
        }
        for (; i < n; i += nt) {
            const u32 w = __ldg(src + i);
            const int v = (cl * (int)(short)(w & 0xFFFF) + cr * ((int)w >> 16)) >> sh;
            c.samp[PADI(i)] = v;
            orv |= (u32)v; diff |= (u32)(v ^ first);
        }
    } else {
        for (u32 i = tid; i < n; i += nt) {
            const int v = ld_candidate(pcm, pcm_off + i, cand, P);
            c.samp[PADI(i)] = v;
            orv |= (u32)v; diff |= (u32)(v ^ first);
        }
    }
    *or_out = orv; *diff_out = diff;
}

// LPC residual of the thread's run into c.resid, chunks of 8, history window in registers.
// OG: coefficients zero-padded to OG taps (exact).  WIDE: 64-bit accumulate (flac.c:1000-1005),
// otherwise 32-bit (only chosen when the sum provably fits).
template <int OG, bool WIDE>
__device__ __forceinline__ void lpc_residual_v2(const V2Ctx& c, const short* q_sm, int shift)
{
    const u32 n = c.n, base = c.base;
    if (base >= n) return;
    int q[OG];
#pragma unroll
    for (int t = 0; t < OG; t++) q[t] = q_sm[t];
    int w[OG + V2_CH]; // w[OG + j] = sample i0 + j, w[OG - 1 - t] = sample i0 - 1 - t
#pragma unroll
    for (int t = 0; t < OG; t++) {
        const int idx = (int)base - 1 - t;
        w[OG - 1 - t] = idx >= 0 ? c.samp[PADI(idx)] : 0;
    }
    const u32 end = min(base + c.S, n);
    for (u32 i0 = base; i0 < end; i0 += V2_CH) {
        const int* sp = CHP(c.samp, i0);
        int* rp = CHP(c.resid, i0);
        const bool full = i0 + V2_CH <= n;
        if (full) {
#pragma unroll
            for (int j = 0; j < V2_CH; j++) w[OG + j] = sp[j];
        } else {
#pragma unroll
            for (int j = 0; j < V2_CH; j++) w[OG + j] = (i0 + j < n) ? sp[j] : 0;
        }
        int res[V2_CH];
#pragma unroll
        for (int j = 0; j < V2_CH; j++) {
            int pred;
            if (WIDE) {
                long long acc = 0;
#pragma unroll
                for (int t = 0; t < OG; t++) acc = mad_wide(q[t], w[OG + j - 1 - t], acc);
                pred = (int)(acc >> shift);
            } else {
                int acc = 0;
#pragma unroll
                for (int t = 0; t < OG; t++) acc += q[t] * w[OG + j - 1 - t];
                pred = acc >> shift;
            }
            res[j] = (int)((u32)w[OG + j] - (u32)pred);
        }
        if (full) {
#pragma unroll
            for (int j = 0; j < V2_CH; j++) rp[j] = res[j];
        } else {
#pragma unroll
            for (int j = 0; j < V2_CH; j++) if (i0 + j < n) rp[j] = res[j];
        }
#pragma unroll
        for (int t = 0; t < OG; t++) w[t] = w[t + V2_CH];
    }
}

// FIXED: error sums of orders 0..4 over samples 4..n-1 of the thread's run (flac.c:877-893)
__device__ __forceinline__ void fixed_sums_v2(const V2Ctx& c, u32 sub_bps, u64 (&e)[5])
{
    const u32 n = c.n, base = c.base;
    e[0] = e[1] = e[2] = e[3] = e[4] = 0;
    if (base >= n) return;
    u32 a1 = 0, a2 = 0, a3 = 0, a4 = 0;
    if (base >= 1) a1 = (u32)c.samp[PADI(base - 1)];
    if (base >= 2) a2 = (u32)c.samp[PADI(base - 2)];
    if (base >= 3) a3 = (u32)c.samp[PADI(base - 3)];
    if (base >= 4) a4 = (u32)c.samp[PADI(base - 4)];
    u32 p1 = a1 - a2, p2 = p1 - (a2 - a3), p3 = p2 - ((a2 - a3) - (a3 - a4));
    u32 prev = a1;
    const u32 end = min(base + c.S, n);
    const bool narrow = sub_bps <= 18; // |d4| < 2^21, S <= 64 terms fit 32 bits
    u32 f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0;
    for (u32 i0 = base; i0 < end; i0 += V2_CH) {
        const int* sp = CHP(c.samp, i0);
        if (narrow && i0 >= 4 && i0 + V2_CH <= n) {
#pragma unroll
            for (int j = 0; j < V2_CH; j++) {
                const u32 x = (u32)sp[j];
                const u32 d1 = x - prev, d2 = d1 - p1, d3 = d2 - p2, d4 = d3 - p3;
                f0 += (u32)abs((int)x); f1 += (u32)abs((int)d1); f2 += (u32)abs((int)d2);
                f3 += (u32)abs((int)d3); f4 += (u32)abs((int)d4);
                prev = x; p1 = d1; p2 = d2; p3 = d3;
            }
        } else {
#pragma unroll
            for (int j = 0; j < V2_CH; j++) {
                const u32 i = i0 + j;
                if (i < n) {
                    const u32 x = (u32)sp[j];
                    const u32 d1 = x - prev, d2 = d1 - p1, d3 = d2 - p2, d4 = d3 - p3;
                    if (i >= 4) {
                        e[0] += (u64)(long long)abs((int)x); e[1] += (u64)(long long)abs((int)d1);
                        e[2] += (u64)(long long)abs((int)d2); e[3] += (u64)(long long)abs((int)d3);
                        e[4] += (u64)(long long)abs((int)d4);
                    }
                    prev = x; p1 = d1; p2 = d2; p3 = d3;
                }
            }
        }
    }
    e[0] += f0; e[1] += f1; e[2] += f2; e[3] += f3; e[4] += f4;
}

// FIXED residual of ORDER (1..4) for the thread's run into c.resid
template <int ORDER>
__device__ __forceinline__ void fixed_residual_order_v2(const V2Ctx& c)
{
    const u32 n = c.n, base = c.base;
    const u32 end = min(base + c.S, n);
    u32 a1 = 0, a2 = 0, a3 = 0, a4 = 0;
    if (base >= 1) a1 = (u32)c.samp[PADI(base - 1)];
    if (base >= 2) a2 = (u32)c.samp[PADI(base - 2)];
    if (base >= 3) a3 = (u32)c.samp[PADI(base - 3)];
    if (base >= 4) a4 = (u32)c.samp[PADI(base - 4)];
    u32 p1 = a1 - a2, p2 = p1 - (a2 - a3), p3 = p2 - ((a2 - a3) - (a3 - a4));
    u32 prev = a1;
    for (u32 i0 = base; i0 < end; i0 += V2_CH) {
        const int* sp = CHP(c.samp, i0);
        int* rp = CHP(c.resid, i0);
        if (i0 + V2_CH <= n) {
#pragma unroll
            for (int j = 0; j < V2_CH; j++) {
                const u32 x = (u32)sp[j];
                const u32 d1 = x - prev;
                u32 v = d1;
                if (ORDER >= 2) { const u32 d2 = d1 - p1; v = d2;
                    if (ORDER >= 3) { const u32 d3 = d2 - p2; v = d3;
                        if (ORDER >= 4) { const u32 d4 = d3 - p3; v = d4; }
                        p3 = d3; }
                    p2 = d2; }
                p1 = d1; prev = x;
                rp[j] = (int)v;
            }
        } else {
#pragma unroll
            for (int j = 0; j < V2_CH; j++) {
                if (i0 + j < n) {
                    const u32 x = (u32)sp[j];
                    const u32 d1 = x - prev, d2 = d1 - p1, d3 = d2 - p2, d4 = d3 - p3;
                    rp[j] = (int)(ORDER == 1 ? d1 : ORDER == 2 ? d2 : ORDER == 3 ? d3 : d4);
                    prev = x; p1 = d1; p2 = d2; p3 = d3;
                }
            }
        }
    }
}

// FIXED residual of `order` for the thread's run into c.resid
__device__ __forceinline__ void fixed_residual_v2(const V2Ctx& c, u32 order)
{
    const u32 n = c.n, base = c.base;
    if (base >= n) return;
    const u32 end = min(base + c.S, n);
    switch (order) {
    case 0:
        for (u32 i = base; i < end; i++) c.resid[PADI(i)] = c.samp[PADI(i)];
        break;
    case 1: fixed_residual_order_v2<1>(c); break;
    case 2: fixed_residual_order_v2<2>(c); break;
    case 3: fixed_residual_order_v2<3>(c); break;
    default: fixed_residual_order_v2<4>(c); break;
    }
}

template <int NTMAX, int MINB>
__global__ void __launch_bounds__(NTMAX, MINB)
k_analyze_v2(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd, bf_dev_params P, u32 S,
             const bf_lpc_head* __restrict__ heads, const short* __restrict__ coefs,
             b200flac_plan* __restrict__ plans, uint8_t* __restrict__ rice_out,
             const u32* __restrict__ frame_list)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ u64 red[40];
    __shared__ RiceScratch rs;
    __shared__ u64 fsum[5];
    __shared__ u32 sc[8];
    __shared__ short s_q[BF_MAX_ORDER];
    __shared__ bf_lpc_head s_head;

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    // frame_list: the frames this launch covers (k_analyze_v3 took the others), or NULL for all
    const u32 cand = blockIdx.x % P.K;
    const u32 frame = frame_list ? frame_list[blockIdx.x / P.K] : blockIdx.x / P.K;
    const u32 unit = frame * P.K + cand;
    const bf_frame_desc d = fd[frame];
    const u32 n = d.nsamp;
    const u32 bps = candidate_bps(cand, P);

    V2Ctx c;
    c.n = n; c.S = S; c.base = (u32)tid * S; c.rs = &rs;
    unsigned char* sp = dyn_smem;
    const size_t padn = (size_t)PADI(P.block_size) + 1;
    c.samp = (int*)sp; sp += padn * 4;
    c.resid = (int*)sp; sp += padn * 4;
    sp = (unsigned char*)(((uintptr_t)sp + 7) & ~(uintptr_t)7);
    c.fine = (u64*)sp; sp += (size_t)P.heap_entries * 8;
    c.karr = sp; sp += P.heap_entries;
    uint8_t* kfix = sp; sp += P.rice_stride;
    uint8_t* klpc = sp;
    AnalyzeCtx ac; // for save_rice()
    ac.karr = c.karr;

    b200flac_plan plan;
    plan.type = BF_VERBATIM; plan.order = 0; plan.wasted = 0; plan.precision = 0; plan.shift = 0;
    plan.coding_method = 0; plan.partition_order = 0; plan.flags = 0; plan.bits = 0;
#pragma unroll
    for (int i = 0; i < BF_MAX_ORDER; i++) plan.coeffs[i] = 0;
    uint8_t* my_rice = rice_out + (size_t)unit * P.rice_stride;

    if (tid == 0 && P.try_lpc) s_head = heads[unit];
    if (tid < 5) fsum[tid] = 0ull;

    // ---- load, constant check, wasted bits ----
    u32 orv, diff;
    load_unit_v2(c, pcm, d.pcm_off, cand, P, &orv, &diff);
    orv = __reduce_or_sync(0xFFFFFFFFu, orv);
    diff = __reduce_or_sync(0xFFFFFFFFu, diff);
    if (lane == 0) { red[tid >> 5] = orv; red[16 + (tid >> 5)] = diff; }
    __syncthreads();
    orv = 0; diff = 0;
    for (int w = 0; w < (nt >> 5); w++) { orv |= (u32)red[w]; diff |= (u32)red[16 + w]; }
    if (P.try_constant && diff == 0) {
        if (tid == 0) { plan.type = BF_CONSTANT; plan.bits = 8 + bps; plans[unit] = plan; }
        return;
    }
    const u32 wasted = orv ? (u32)(__ffs((int)orv) - 1) : 0u;
    if (wasted) {
        for (u32 i = tid; i < n; i += nt) c.samp[PADI(i)] >>= wasted;
        __syncthreads();
    }
    const u32 sub_bps = bps - wasted;
    const u32 hdr_bits = 8 + wasted;
    u32 search = 0;

    // ---- FIXED ----
    u64 fixed_bits = 0;
    u32 fixed_order = 0;
    RiceChoice rfix; rfix.po = 0; rfix.under = 0; rfix.k0 = 0; rfix.method = 0; rfix.bits = 0;
    if (P.try_fixed) {
        if (n > 4) {
            u64 e[5];
            fixed_sums_v2(c, sub_bps, e);
            if (!__any_sync(0xFFFFFFFFu, ((e[0] | e[1] | e[2] | e[3] | e[4]) >> 26) != 0)) {
                u32 f[5];
#pragma unroll
                for (int k = 0; k < 5; k++) f[k] = __reduce_add_sync(0xFFFFFFFFu, (u32)e[k]);
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < 5; k++) atomicAdd(&fsum[k], (u64)f[k]);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    u64 v = e[k];
#pragma unroll
                    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
                    if (lane == 0) atomicAdd(&fsum[k], v);
                }
            }
            __syncthreads();
            u64 best = fsum[0];
            if (fsum[1] < best) { best = fsum[1]; fixed_order = 1; }
            if (fsum[2] < best) { best = fsum[2]; fixed_order = 2; }
            if (fsum[3] < best) { best = fsum[3]; fixed_order = 3; }
            if (fsum[4] < best) { best = fsum[4]; fixed_order = 4; }
        }
        fixed_residual_v2(c, fixed_order);
        // each thread reads back only its own residuals in pass 1/2: no barrier needed here
        rice_search_v2(c, (search++) & 1u, fixed_order, P, &rfix);
        save_rice(ac, rfix, kfix);
        fixed_bits = hdr_bits + (u64)fixed_order * sub_bps + rfix.bits;
    }

    // ---- LPC ----
    u64 lpc_bits = 0;
    u32 lpc_order = 0, lpc_precision = 0;
    int lpc_shift = 0;
    RiceChoice rlpc; rlpc.po = 0; rlpc.under = 0; rlpc.k0 = 0; rlpc.method = 0; rlpc.bits = 0;
    const short* mycoef = coefs + (size_t)unit * P.model_stride;
    if (P.try_lpc) {
        const u32 best_order = s_head.best_order, dummy = s_head.dummy, precision = s_head.precision;
        const u32 o_first = (P.exhaustive && !dummy) ? 1u : best_order;
        const u32 o_last = (P.exhaustive && !dummy) ? P.max_lpc_order : best_order;
        u64 best_bits = 0xFFFFFFFFull;
        bool have = false;
        for (u32 o = o_first; o <= o_last; o++) {
            __syncthreads();
            if (tid < 32) {
                const int q = tid < (int)o ? (int)mycoef[(o * (o - 1)) / 2 + tid] : 0;
                s_q[tid] = (short)q;
                const u32 sumq = __reduce_add_sync(0xFFFFFFFFu, (u32)abs(q));
                if (tid == 0) sc[4] = (((u64)sumq << (sub_bps - 1)) < (1ull << 31)) ? 1u : 0u;
            }
            __syncthreads();
            const int shift = s_head.shift[o - 1];
            const bool narrow = sc[4] != 0;
            if (o <= 8) { if (narrow) lpc_residual_v2<8, false>(c, s_q, shift); else lpc_residual_v2<8, true>(c, s_q, shift); }
            else if (o <= 12) { if (narrow) lpc_residual_v2<12, false>(c, s_q, shift); else lpc_residual_v2<12, true>(c, s_q, shift); }
            else { if (narrow) lpc_residual_v2<32, false>(c, s_q, shift); else lpc_residual_v2<32, true>(c, s_q, shift); }
            RiceChoice rc;
            rice_search_v2(c, (search++) & 1u, o, P, &rc);
            const u64 bits = hdr_bits + (u64)o * sub_bps + 4 + 5 + (u64)o * precision + rc.bits;
            if (!have || (u32)bits < best_bits) {
                have = true;
                best_bits = (u32)bits;
                lpc_bits = bits; lpc_order = o; lpc_shift = shift; lpc_precision = precision;
                rlpc = rc;
                save_rice(ac, rc, klpc);
            }
        }
        __syncthreads();
    }

    // ---- choice, flac.c:727-809 ----
    const u32 fb = (u32)fixed_bits, lb = (u32)lpc_bits;
    const u32 vb = P.try_verbatim ? sub_bps * n : 0x7FFFFFFFu;
    u32 choice;
    if (P.try_fixed && P.try_lpc && P.try_verbatim)
        choice = (fb < min(lb, vb)) ? BF_FIXED : (lb < vb) ? BF_LPC : BF_VERBATIM;
    else if (!P.try_fixed && !P.try_lpc) choice = BF_VERBATIM;
    else if (P.try_fixed && !P.try_lpc && !P.try_verbatim) choice = BF_FIXED;
    else if (!P.try_fixed && P.try_lpc && !P.try_verbatim) choice = BF_LPC;
    else if (P.try_fixed && P.try_lpc && !P.try_verbatim) choice = (fb < lb) ? BF_FIXED : BF_LPC;
    else if (P.try_fixed && !P.try_lpc && P.try_verbatim) choice = (fb < vb) ? BF_FIXED : BF_VERBATIM;
    else choice = (lb < vb) ? BF_LPC : BF_VERBATIM;

    plan.wasted = (uint8_t)wasted;
    if (choice == BF_FIXED) {
        plan.type = BF_FIXED; plan.order = (uint8_t)fixed_order;
        plan.coding_method = (uint8_t)rfix.method; plan.partition_order = (uint8_t)rfix.po;
        plan.flags = (uint8_t)rfix.under; plan.bits = fb;
        for (u32 p = tid; p < (1u << rfix.po); p += nt) my_rice[p] = kfix[p];
    } else if (choice == BF_LPC) {
        plan.type = BF_LPC; plan.order = (uint8_t)lpc_order;
        plan.precision = (uint8_t)lpc_precision; plan.shift = (int8_t)lpc_shift;
        plan.coding_method = (uint8_t)rlpc.method; plan.partition_order = (uint8_t)rlpc.po;
        plan.flags = (uint8_t)rlpc.under; plan.bits = lb;
        if (tid == 0) {
            const short* bc = mycoef + (lpc_order * (lpc_order - 1)) / 2;
            u32 sumq = 0;
            for (u32 j = 0; j < lpc_order; j++) { plan.coeffs[j] = bc[j]; sumq += (u32)abs((int)bc[j]); }
            if (((u64)sumq << (sub_bps - 1)) < (1ull << 31)) plan.flags |= 2; // packer may accumulate in 32 bits
        }
        for (u32 p = tid; p < (1u << rlpc.po); p += nt) my_rice[p] = klpc[p];
    } else {
        plan.type = BF_VERBATIM;
        plan.bits = hdr_bits + sub_bps * n;
    }
    if (tid == 0) plans[unit] = plan;
}
