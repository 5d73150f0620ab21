// k_analyze_v3.cuh -- the model-search kernel for full-length blocks of the common shapes
// (k_analyze_v2 takes every other block: tails, exhaustive search, disabled subframe types,
// partition lengths that are not a whole number of thread runs).
//
// Same decisions, bit for bit, as k_analyze / k_analyze_v2 (and the reference, flac.c:673-1505);
// one CTA per unit (frame, candidate), every thread owns a contiguous run of S samples.  What
// v3 changes is how little it executes around the arithmetic the reference defines:
//   * one pass over the samples yields the FIXED error sums of all five orders AND, per thread
//     run, the sums of |residual| the Rice search needs for whichever order wins -- the FIXED
//     residual is never written anywhere;
//   * the LPC residual pass accumulates its own run sums while it stores the residual;
//   * the two Rice searches (FIXED, LPC) run side by side, each inside ONE warp, on prefix sums
//     of the run sums: no block-wide barrier, no redundant per-thread level loops;
//   * one fused pass then counts the exact bits of both models (FIXED residual recomputed on the
//     fly, LPC residual re-read); the per-partition (1 + k) terms come from the search warp;
//   * samples are skewed four words per 32 in shared memory, so every 8-sample chunk is two
//     conflict-free 128-bit accesses;
//   * five block-wide barriers per unit instead of fourteen.
#pragma once
#include "flac_common.cuh"
#include "k_analyze.cuh"
#include "k_analyze_v2.cuh"

#define V3_CH 8
// The chunk loops stay rolled even where the trip count is a compile-time constant (SC = 32 gives four
// iterations): unrolled, the kernel's hot code was ~40 KB against a 32 KB instruction cache.
#ifndef V3_ROLL
#define V3_ROLL 1
#endif
#if V3_ROLL
#define V3_LOOP _Pragma("unroll 1")
#else
#define V3_LOOP
#endif
#ifndef V3_F64_UNROLL
#define V3_F64_UNROLL 2
#endif
#if V3_F64_UNROLL == 2
#define V3_F64_LOOP _Pragma("unroll 2")
#else
#define V3_F64_LOOP _Pragma("unroll 1")
#endif
#define V3_SK(i) ((i) + (((i) >> 5) << 2))      // four words of skew per 32 samples
// Finest partition order handled: 7 (<= 128 partitions, one or more whole thread runs each), or 8 in the SUB = 2
// instantiations, where a thread's run of 32 samples is TWO finest partitions of 16 and every run sum is kept per
// half run.  (Order 8 with 256 threads x 16 samples was slower than k_analyze_v2: 10.6 vs 7.2 ms per 5 minutes of
// 96 kHz/24-bit -e -R8; the extra shared memory of order 8 costs the common shapes a resident CTA, hence the template.)
#define V3_MAX_F 7
#define V3_LVL_H0 9                              // lvl[..][9], [10]: the two halves of a finest level evaluated by two warps
#define V3_LVL_H1 10

// totals of one partition order of one model (written by whichever warp evaluated the level)
struct V3Level {
    u64 tot;         // the reference's size estimate of the residual block at this order
    u32 cnt;         // sum over partitions of (1 + k) * residuals in it
    u32 maxk;
};

template <int MF>
struct V3SharedT {
    static constexpr int HEAP = 2 << MF;
    V3Level lvl[2][11];
    u64 totF[5];          // FIXED: block totals of the error sums (flac.c:877-893), wide blocks
    u32 totF16[5][2];     // ... narrow blocks: sums of the low 16 bits / the rest of the warp sums
    u32 bits16[2][2];     // exact sum of (u >> k), per model, split the same way
    u64 corr[5];          // FIXED: sum of |r_k[i]| for k <= i < 4 (in the partition sums, not in the order choice)
    u32 red_or[16], red_diff[16];
    u32 lpc_narrow;       // sum of |coefficient| of the staged order (turned into "the sum fits 32 bits" by its readers)
    u32 lpc_narrow2[2];   // exhaustive search: the same for the two staged orders
    u32 bitsL[2][2];      // exhaustive search: exact sum of (u >> k) of the orders in flight (by order parity)
    short q[BF_MAX_ORDER];
    short q2[2][BF_MAX_ORDER];                  // exhaustive search: coefficients of the order being run and the next
    uint8_t kheapL[2][HEAP];                    // exhaustive search: Rice parameters of the two orders in flight
    alignas(4) bf_lpc_head head;      // (copied as words)
    uint8_t kheap[2][HEAP];
    uint8_t kbest[HEAP / 2];   // exhaustive search: Rice parameters of the best LPC order so far
};

// host and device agree on the dynamic shared memory through this
// (sub: run sums kept per thread run, 1 or 2; own: the exhaustive 128 x 32 instantiations, whose run sums stay in
// registers -- they only need 1 KB for the warps' level totals)
__host__ __device__ inline size_t v3_smem_bytes(u32 n, u32 NT, u32 sub = 1, bool own = false)
{
    const size_t padn = (size_t)V3_SK(n) + 8;
    return 2 * padn * 4 + (own ? (size_t)1024 : (size_t)6 * NT * sub * 8) + 32;
}

__device__ __forceinline__ u64 v3_warp_sum_u64(u64 v)
{
    if (!__any_sync(0xFFFFFFFFu, (v >> 26) != 0)) return (u64)__reduce_add_sync(0xFFFFFFFFu, (u32)v);
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// One warp evaluates a share of the partition orders of one model (flac.c:1437-1505) from the run
// sums: runs[t] = sum of |r| over thread run t (run 0 without its first `order` samples, up to
// first_extra, which is added here).  One lane per partition ("node"), 32 nodes per step.
//   part 0: the finest order F -- a node is g consecutive run sums;
//   part 1: every order below F -- nodes are differences of the prefix sums of the finest sums,
//           which this warp builds in `pre` (shared, nfine entries) first.
// The two parts are independent, so two warps can run them side by side.  Each part also comes in two halves
// for four warps (the exhaustive search runs one model at a time): parts 2 and 3 take alternate steps of the
// finest order and leave their totals in lvl[V3_LVL_H0] / [V3_LVL_H1] (v3_pick_level adds them up),
// part 4 takes orders 0..4 (nodes 1..31) and order 7 if there is one below the finest, part 5 orders 5 and 6;
// 4 and 5 each build their own prefix sums.  S: samples per run sum.
__device__ __noinline__ void v3_levels(const u64* __restrict__ runs, u64 first_extra, u32 S, u32 n, u32 order,
                                       u32 F, u32 max_rice, uint8_t* kheap, V3Level* lvl, u64* pre, u32 part)
{
    const u32 lane = threadIdx.x & 31;
    const u32 nfine = 1u << F;
    const u32 g = (n >> F) / S;                          // thread runs per finest partition
    if (part == 0 || part == 2 || part == 3) {
        u64 est_acc = 0; u32 cnt_acc = 0, k_acc = 0;
        const u32 p0 = part == 3 ? 32u : 0u, pstep = part == 0 ? 32u : 64u;
        for (u32 p = p0 + lane; p < nfine; p += pstep) {
            const u64* r = runs + (size_t)p * g;
            u64 sum = p == 0 ? first_extra : 0ull;
#pragma unroll 1
            for (u32 j = 0; j < g; j++) sum += r[j];
            const u32 plength = (n >> F) - (p == 0 ? order : 0u);
            u32 k;
            est_acc += partition_estimate_fast(plength, sum, max_rice, &k);
            kheap[nfine - 1u + p] = (uint8_t)k;
            cnt_acc += (1u + k) * plength;
            k_acc = max(k_acc, k);
        }
        const u64 tot = v3_warp_sum_u64(est_acc);
        const u32 cnt = __reduce_add_sync(0xFFFFFFFFu, cnt_acc);
        const u32 mk = __reduce_max_sync(0xFFFFFFFFu, k_acc);
        V3Level* out = part == 0 ? lvl + F : lvl + V3_LVL_H0 + (part - 2);
        if (lane == 0) { out->tot = tot; out->cnt = cnt; out->maxk = mk; }
        return;
    }
    if (F == 0) return;
    // inclusive prefix sums of the finest sums: lane owns c = ceil(nfine / 32) consecutive ones
    {
        const u32 c = (nfine + 31) >> 5, b = lane * c;
        u64 run = 0;
        for (u32 i = 0; i < c; i++) {
            const u32 p = b + i;
            if (p < nfine) {
                const u64* r = runs + (size_t)p * g;
                u64 sum = p == 0 ? first_extra : 0ull;
#pragma unroll 1
                for (u32 j = 0; j < g; j++) sum += r[j];
                run += sum;
                pre[p] = run;
            }
        }
        u64 inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= (u32)o) inc += t;
        }
        const u64 excl = inc - run;
        for (u32 i = 0; i < c; i++) if (b + i < nfine) pre[b + i] += excl;
    }
    __syncwarp();
    for (u32 node0 = part == 5 ? 32u : 0u; node0 < (part == 5 ? min(nfine, 128u) : nfine); node0 += 32) {
        if (part == 4 && node0 == 32) { node0 = 96; continue; }        // (orders 5 and 6 are part 5's)
        const u32 node = node0 + lane;
        const bool act = node >= 1 && node < nfine;
        u32 l = 0, k = 0, cnt = 0;
        u64 est = 0;
        if (act) {
            l = 31u - (u32)__clz((int)node);
            const u32 p = node - (1u << l);
            const u32 w = nfine >> l;                       // finest partitions per partition of this order
            const u64 hi = pre[(p + 1) * w - 1];
            const u64 lo = p ? pre[p * w - 1] : 0ull;
            const u32 plength = (n >> l) - (p == 0 ? order : 0u);
            est = partition_estimate_fast(plength, hi - lo, max_rice, &k);
            kheap[node - 1] = (uint8_t)k;
            cnt = (1u + k) * plength;
        }
        if (node0 >= 32) {
            // all 32 nodes of this step belong to one order; an order of 64 partitions takes two steps
            const u32 l0 = 31u - (u32)__clz((int)node0);
            const u64 tot = v3_warp_sum_u64(est);
            const u32 cs = __reduce_add_sync(0xFFFFFFFFu, cnt);
            const u32 km = __reduce_max_sync(0xFFFFFFFFu, k);
            if (lane == 0) {
                if ((node0 & (node0 - 1)) == 0) { lvl[l0].tot = tot; lvl[l0].cnt = cs; lvl[l0].maxk = km; }
                else { lvl[l0].tot += tot; lvl[l0].cnt += cs; lvl[l0].maxk = max(lvl[l0].maxk, km); }
            }
        } else {
            // nodes 1..31: orders 0..4, order l occupies lanes 2^l .. 2^(l+1)-1
            const u32 gs = act ? (1u << l) : 1u;
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) {
                const u64 te = __shfl_xor_sync(0xFFFFFFFFu, est, o);
                const u32 tc = __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
                const u32 tk = __shfl_xor_sync(0xFFFFFFFFu, k, o);
                if ((u32)o < gs) { est += te; cnt += tc; k = max(k, tk); }
            }
            if (act && node == gs) { lvl[l].tot = est; lvl[l].cnt = cnt; lvl[l].maxk = k; }
        }
    }
}

// every warp: first strict minimum of the estimates over the partition orders (flac.c:1365-1400);
// lane l looks at order l, the minimum of (estimate, order) pairs is taken with two warp reductions
// split: the finest order's totals are the sums of the two halves left by parts 2 and 3 of v3_levels
__device__ __forceinline__ void v3_pick_level(const V3Level* lvl, u32 F, u32* po_out, u32* method_out, u64* side_bits,
                                              bool split = false)
{
    const u32 lane = threadIdx.x & 31;
    const V3Level* h0 = lvl + V3_LVL_H0;
    const V3Level* h1 = lvl + V3_LVL_H1;
    const u64 tot = lane > F ? ~0ull : (split && lane == F) ? h0->tot + h1->tot : lvl[lane].tot;
    const u32 hi = (u32)(tot >> 32), lo = (u32)tot;
    const u32 mhi = __reduce_min_sync(0xFFFFFFFFu, hi);
    const u32 mlo = __reduce_min_sync(0xFFFFFFFFu, hi == mhi ? lo : 0xFFFFFFFFu);
    const u32 po = __reduce_min_sync(0xFFFFFFFFu, (hi == mhi && lo == mlo) ? lane : 32u);
    const bool fin = split && po == F;
    const u32 maxk = fin ? max(h0->maxk, h1->maxk) : lvl[po].maxk;
    *po_out = po;
    *method_out = maxk > 14 ? 1u : 0u;
    *side_bits = 6ull + (u64)(1u << po) * (maxk > 14 ? 5ull : 4ull) + (u64)(fin ? h0->cnt + h1->cnt : lvl[po].cnt);
}

// ---- exhaustive search, 128 threads x 32 samples (n = 4096): the Rice search of one LPC order straight from the
// threads' run sums, which never leave the registers.  A warp's 32 runs are 1/4 of the block, so every partition of
// orders 2 and up lies inside one warp: each warp evaluates the nodes under its own runs -- order 8 (two per thread,
// SUB = 2), order 7 (one per thread), orders 6..2 (16 + 8 + 4 + 2 + 1 = 31 nodes, one per lane, from a warp prefix
// sum) -- with no barrier before it, and leaves per-level totals for v3_pick_own, which adds the four warps up and
// evaluates the three nodes of orders 1 and 0 after the order's single barrier.
#define V3_OWN_LEVELS 7                          // per-warp totals of orders 2..8
template <int SUB>
__device__ __forceinline__ void v3_search_own(u64 s0, u64 s1, u32 order, u32 F, u32 n, u32 max_rice,
                                              uint8_t* __restrict__ kheap, V3Level* __restrict__ part, u64* __restrict__ wsum)
{
    const u32 tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 warm = tid == 0 ? order : 0u;
    const u64 r = SUB == 2 ? s0 + s1 : s0;
    if (SUB == 2 && F >= 8) {
        u64 est = 0; u32 cnt = 0, mk = 0;
#pragma unroll 1
        for (u32 h = 0; h < 2; h++) {
            const u32 pl = (n >> 8) - (h ? 0u : warm);
            u32 k;
            est += partition_estimate_fast(pl, h ? s1 : s0, max_rice, &k);
            kheap[255u + 2u * tid + h] = (uint8_t)k;
            cnt += (1u + k) * pl; mk = max(mk, k);
        }
        const u64 tot = v3_warp_sum_u64(est);
        const u32 cs = __reduce_add_sync(0xFFFFFFFFu, cnt), km = __reduce_max_sync(0xFFFFFFFFu, mk);
        if (lane == 0) { part[6].tot = tot; part[6].cnt = cs; part[6].maxk = km; }
    }
    if (F >= 7) {
        const u32 pl = (n >> 7) - warm;
        u32 k;
        const u64 est = partition_estimate_fast(pl, r, max_rice, &k);
        kheap[127u + tid] = (uint8_t)k;
        const u64 tot = v3_warp_sum_u64(est);
        const u32 cs = __reduce_add_sync(0xFFFFFFFFu, (1u + k) * pl), km = __reduce_max_sync(0xFFFFFFFFu, k);
        if (lane == 0) { part[5].tot = tot; part[5].cnt = cs; part[5].maxk = km; }
    }
    // inclusive prefix sums of the warp's run sums
    u64 inc = r;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u64 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (u32)o) inc += t;
    }
    if (lane == 31) *wsum = inc;
    // lane j >= 1: node j of the warp's tree -- local level l = floor(log2 j) (order l + 2), w = 32 >> l runs wide
    {
        const u32 l = 31u - (u32)__clz((int)(lane | 1u));
        const u32 idx = lane - (1u << l), w = 32u >> l;
        const u64 hi = __shfl_sync(0xFFFFFFFFu, inc, (idx + 1u) * w - 1u);
        const u64 lo = __shfl_sync(0xFFFFFFFFu, inc, (idx * w - 1u) & 31u);
        const bool act = lane >= 1 && l + 2u <= F;
        u32 k = 0, cnt = 0;
        u64 est = 0;
        if (act) {
            const u32 p = (warp << l) + idx;
            const u32 pl = (n >> (l + 2u)) - (p == 0 ? order : 0u);
            est = partition_estimate_fast(pl, hi - (idx ? lo : 0ull), max_rice, &k);
            kheap[(4u << l) - 1u + p] = (uint8_t)k;
            cnt = (1u + k) * pl;
        }
        // level l occupies lanes 2^l .. 2^(l+1)-1: sums within aligned groups of 2^l lanes
        const u32 gs = 1u << l;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const u64 te = __shfl_xor_sync(0xFFFFFFFFu, est, o);
            const u32 tc = __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
            const u32 tk = __shfl_xor_sync(0xFFFFFFFFu, k, o);
            if ((u32)o < gs) { est += te; cnt += tc; k = max(k, tk); }
        }
        if (act && lane == gs) { part[l].tot = est; part[l].cnt = cnt; part[l].maxk = k; }
    }
}

// after the barrier, every warp: totals of orders 2..F from the four warps' parts (lane = order), the three nodes of
// orders 1 and 0 from the warps' sums (lanes 29..31, moved to lanes 0 and 1), then the first strict minimum of the
// estimates (flac.c:1365-1400).  All warps write the same three Rice parameters.
__device__ __forceinline__ void v3_pick_own(const V3Level* __restrict__ parts, const u64* __restrict__ wsum, u32 order,
                                            u32 F, u32 n, u32 max_rice, uint8_t* __restrict__ kheap,
                                            u32* po_out, u32* method_out, u64* side_bits)
{
    const u32 lane = threadIdx.x & 31;
    u64 tot = ~0ull;
    u32 cnt = 0, mk = 0;
    if (lane >= 2 && lane <= F) {
        tot = 0;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const V3Level v = parts[w * V3_OWN_LEVELS + (lane - 2)];
            tot += v.tot; cnt += v.cnt; mk = max(mk, v.maxk);
        }
    }
    u64 te = 0;
    u32 tc = 0, tk = 0;
    if (lane >= 29) {
        const u64 a = wsum[0] + wsum[1], b = wsum[2] + wsum[3];
        const u32 L = lane == 29 ? 0u : 1u;
        const u32 pl = (n >> L) - (lane == 31 ? 0u : order);
        te = partition_estimate_fast(pl, lane == 29 ? a + b : lane == 30 ? a : b, max_rice, &tk);
        kheap[lane - 29u] = (uint8_t)tk;
        tc = (1u + tk) * pl;
    }
    {
        const u64 e0 = __shfl_sync(0xFFFFFFFFu, te, 29), e1 = __shfl_sync(0xFFFFFFFFu, te, 30), e2 = __shfl_sync(0xFFFFFFFFu, te, 31);
        const u32 c0 = __shfl_sync(0xFFFFFFFFu, tc, 29), c1 = __shfl_sync(0xFFFFFFFFu, tc, 30), c2 = __shfl_sync(0xFFFFFFFFu, tc, 31);
        const u32 k0 = __shfl_sync(0xFFFFFFFFu, tk, 29), k1 = __shfl_sync(0xFFFFFFFFu, tk, 30), k2 = __shfl_sync(0xFFFFFFFFu, tk, 31);
        if (lane == 0) { tot = e0; cnt = c0; mk = k0; }
        if (lane == 1 && F >= 1) { tot = e1 + e2; cnt = c1 + c2; mk = max(k1, k2); }
    }
    __syncwarp();
    const u32 hi = (u32)(tot >> 32), lo = (u32)tot;
    const u32 mhi = __reduce_min_sync(0xFFFFFFFFu, hi);
    const u32 mlo = __reduce_min_sync(0xFFFFFFFFu, hi == mhi ? lo : 0xFFFFFFFFu);
    const u32 po = __reduce_min_sync(0xFFFFFFFFu, (hi == mhi && lo == mlo) ? lane : 32u);
    const u32 maxk = __shfl_sync(0xFFFFFFFFu, mk, po);
    const u32 c = __shfl_sync(0xFFFFFFFFu, cnt, po);
    *po_out = po;
    *method_out = maxk > 14 ? 1u : 0u;
    *side_bits = 6ull + (u64)(1u << po) * (maxk > 14 ? 5ull : 4ull) + (u64)c;
}

// predictor history of a FIXED pass: differences of the four samples before `base` (zeros for run 0)
__device__ __forceinline__ void v3_fixed_history(const int* __restrict__ samp, u32 base, u32& prev, u32& p1, u32& p2, u32& p3)
{
    int4 h = make_int4(0, 0, 0, 0);
    if (base) h = *(const int4*)(samp + V3_SK(base - 4));
    const u32 a4 = (u32)h.x, a3 = (u32)h.y, a2 = (u32)h.z, a1 = (u32)h.w;
    p1 = a1 - a2; p2 = p1 - (a2 - a3); p3 = p2 - ((a2 - a3) - (a3 - a4));
    prev = a1;
}

// FIXED error sums of orders 0..4 over the thread's run: e[k] = sum of |r_k[i]| (flac.c:877-893).
// Run 0 is summed like the others (zero history) and corrected by v3_fixed_head afterwards.
// SUB = 2: the sums of the first half run go to half0[k * stride] when the half ends, e[] gets the second half's.
template <typename SumT, int SUB>
__device__ __forceinline__ void v3_fixed_sums(const int* __restrict__ samp, u32 base, u32 S, SumT (&e)[5],
                                              u64* __restrict__ half0, u32 stride)
{
    u32 prev, p1, p2, p3;
    v3_fixed_history(samp, base, prev, p1, p2, p3);
    SumT f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0;
    V3_LOOP
    for (u32 i0 = base; i0 < base + S; i0 += V3_CH) {
        if (SUB == 2 && i0 == base + (S >> 1)) {
            half0[0] = (u64)f0; half0[stride] = (u64)f1; half0[2 * stride] = (u64)f2; half0[3 * stride] = (u64)f3;
            half0[4 * stride] = (u64)f4;
            f0 = 0; f1 = 0; f2 = 0; f3 = 0; f4 = 0;
        }
        const int4 va = *(const int4*)(samp + V3_SK(i0));
        const int4 vb = *(const int4*)(samp + V3_SK(i0) + 4);
        const int xs[V3_CH] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
        if (sizeof(SumT) == 4) {
            // 32-bit sums (samples of at most 23 bits): |a - b| + c is ONE instruction (VABSDIFF, __sad), so
            // an order's error never needs its difference formed first -- 3 subtractions and 5 VABSDIFFs per
            // sample instead of 4 + 5 IABS + 5 adds.  No difference can overflow 32 bits here, so the true
            // |a - b| equals the reference's abs() of its wrapped int difference.
            u32 g0 = (u32)f0, g1 = (u32)f1, g2 = (u32)f2, g3 = (u32)f3, g4 = (u32)f4;
#pragma unroll
            for (int j = 0; j < V3_CH; j++) {
                const int x = xs[j];
                g0 = __sad(x, 0, g0);
                g1 = __sad(x, (int)prev, g1);
                const u32 d1 = (u32)x - prev;
                g2 = __sad((int)d1, (int)p1, g2);
                const u32 d2 = d1 - p1;
                g3 = __sad((int)d2, (int)p2, g3);
                const u32 d3 = d2 - p2;
                g4 = __sad((int)d3, (int)p3, g4);
                prev = (u32)x; p1 = d1; p2 = d2; p3 = d3;
            }
            f0 = g0; f1 = g1; f2 = g2; f3 = g3; f4 = g4;
        } else {
#pragma unroll
            for (int j = 0; j < V3_CH; j++) {
                const u32 x = (u32)xs[j];
                const u32 d1 = x - prev, d2 = d1 - p1, d3 = d2 - p2, d4 = d3 - p3;
                f0 += (u32)abs((int)x); f1 += (u32)abs((int)d1); f2 += (u32)abs((int)d2);
                f3 += (u32)abs((int)d3); f4 += (u32)abs((int)d4);
                prev = x; p1 = d1; p2 = d2; p3 = d3;
            }
        }
    }
    e[0] = f0; e[1] = f1; e[2] = f2; e[3] = f3; e[4] = f4;
}

// LPC residual of the thread's run (flac.c:999-1008) into resid, chunks of 8, history window in
// registers; returns the run's sum of |r|.  OG: taps (coefficients zero-padded, exact).
// WIDE: 64-bit accumulate, otherwise 32-bit (only chosen when the sum provably fits).
// SUB = 2: the sum of the first half run goes to *half0 when the half ends, the second half's is returned.
#define V3_ACC_I32 0
#define V3_ACC_I64 1
// UNR: chunks per loop iteration (2 in the exhaustive 128 x 32 kernel of partition orders <= 7 -- level 8's --, whose
// instruction cache has room; at order 8, 24-bit territory, the narrow pass is cold code and only costs space: the history window is
// back in place after 16 samples at 8 taps and has moved by 4 at 12 -- none or half of the register moves per chunk).
template <int OG, int WIDE, int SUB, int UNR = 1>
__device__ __forceinline__ u64 v3_lpc_residual(const int* __restrict__ samp, int* __restrict__ resid, u32 base, u32 S,
                                               const short* q_sm, int shift, u64* __restrict__ half0)
{
    int q[OG];
#pragma unroll
    for (int t = 0; t < OG; t++) q[t] = q_sm[t];
    int w[OG + V3_CH]; // w[OG + j] = sample i0 + j, w[OG - 1 - t] = sample i0 - 1 - t
    // the OG samples before the run: whole 128-bit words of the previous run (OG is a multiple of 4)
#pragma unroll
    for (int t = 0; t < OG; t += 4) {
        int4 h = make_int4(0, 0, 0, 0);
        if (base >= (u32)(OG - t)) h = *(const int4*)(samp + V3_SK(base - (OG - t)));
        w[t] = h.x; w[t + 1] = h.y; w[t + 2] = h.z; w[t + 3] = h.w;
    }
    u64 run = 0;
    V3_LOOP
    for (u32 ib = base; ib < base + S; ib += UNR * V3_CH) {
#pragma unroll
      for (int un = 0; un < UNR; un++) {
        const u32 i0 = ib + un * V3_CH;
        if (SUB == 2 && i0 == base + (S >> 1)) { *half0 = run; run = 0; }
        const int4 va = *(const int4*)(samp + V3_SK(i0));
        const int4 vb = *(const int4*)(samp + V3_SK(i0) + 4);
        w[OG + 0] = va.x; w[OG + 1] = va.y; w[OG + 2] = va.z; w[OG + 3] = va.w;
        w[OG + 4] = vb.x; w[OG + 5] = vb.y; w[OG + 6] = vb.z; w[OG + 7] = vb.w;
        int res[V3_CH];
#pragma unroll
        for (int j = 0; j < V3_CH; j++) {
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
        if (WIDE == V3_ACC_I32) {
            // every |residual| is below 2^29 here (see is_narrow): |a - 0| + c is one VABSDIFF
            u32 cs = 0;
#pragma unroll
            for (int j = 0; j < V3_CH; j++) cs = __sad(res[j], 0, cs);
            run += cs;
        } else {
#pragma unroll
            for (int j = 0; j < V3_CH; j++) run += (u64)(u32)abs(res[j]);
        }
        *(int4*)(resid + V3_SK(i0)) = make_int4(res[0], res[1], res[2], res[3]);
        *(int4*)(resid + V3_SK(i0) + 4) = make_int4(res[4], res[5], res[6], res[7]);
#pragma unroll
        for (int t = 0; t < OG; t++) w[t] = w[t + V3_CH];
      }
    }
    return run;
}

// The same pass for wide samples (more than 32 bits of sum) in FP64: coefficients of at most 16 bits times samples of
// at most 26 are exact products, and up to 32 of them add up below 2^53, so the DFMA chain IS the reference's 64-bit
// integer sum -- one instruction per tap on the otherwise idle FP64 pipe, where the integer form costs two (ptxas
// splits mad.wide into IMAD.WIDE + 64-bit IADD3: the accumulating form runs at half rate) on the pipe that limits
// this kernel.  floor(acc / 2^shift) mod 2^32 -- the reference's arithmetic shift and truncating cast -- is one more
// DFMA: acc * 2^-shift + 1.5 * 2^52, rounded DOWN, has the integer in its low mantissa word.  No I2F/F2I (quarter rate).
template <int OG, int SUB>
__device__ __forceinline__ u64 v3_lpc_residual_f64(const int* __restrict__ samp, int* __restrict__ resid, u32 base, u32 S,
                                                   const short* q_sm, int shift, u64* __restrict__ half0)
{
    double q[OG];
#pragma unroll
    for (int t = 0; t < OG; t++) q[t] = int2double_exact((int)q_sm[t]);
    double w[OG + V3_CH];
#pragma unroll
    for (int t = 0; t < OG; t += 4) {
        int4 h = make_int4(0, 0, 0, 0);
        if (base >= (u32)(OG - t)) h = *(const int4*)(samp + V3_SK(base - (OG - t)));
        w[t] = int2double_exact(h.x); w[t + 1] = int2double_exact(h.y);
        w[t + 2] = int2double_exact(h.z); w[t + 3] = int2double_exact(h.w);
    }
    const double scale = __hiloint2double((1023 - shift) << 20, 0);      // 2^-shift
    u64 run = 0;
    // (two chunks per iteration: the window of 8 + 8 doubles is back in place after 16 samples, that of 12 + 8
    // has moved by 4 -- half or none of the 2 x OG register moves a single-chunk loop pays per chunk)
    V3_F64_LOOP
    for (u32 i0 = base; i0 < base + S; i0 += V3_CH) {
        if (SUB == 2 && i0 == base + (S >> 1)) { *half0 = run; run = 0; }
        const int4 va = *(const int4*)(samp + V3_SK(i0));
        const int4 vb = *(const int4*)(samp + V3_SK(i0) + 4);
        const int xs[V3_CH] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int j = 0; j < V3_CH; j++) w[OG + j] = int2double_exact(xs[j]);
        int res[V3_CH];
#pragma unroll
        for (int j = 0; j < V3_CH; j++) {
            double acc = 0.0;
#pragma unroll
            for (int t = 0; t < OG; t++) acc = __fma_rn(q[t], w[OG + j - 1 - t], acc);
            const int pred = __double2loint(__fma_rd(acc, scale, 6755399441055744.0));
            res[j] = (int)((u32)xs[j] - (u32)pred);
        }
        {
            // every |residual| is below 2^28 here (see is_f64): |a - 0| + c is one VABSDIFF
            u32 cs = 0;
#pragma unroll
            for (int j = 0; j < V3_CH; j++) cs = __sad(res[j], 0, cs);
            run += cs;
        }
        *(int4*)(resid + V3_SK(i0)) = make_int4(res[0], res[1], res[2], res[3]);
        *(int4*)(resid + V3_SK(i0) + 4) = make_int4(res[4], res[5], res[6], res[7]);
#pragma unroll
        for (int t = 0; t < OG; t++) w[t] = w[t + V3_CH];
    }
    return run;
}

// zigzag(r) >> k for k >= 1 without forming zigzag(r): with t = r >= 0 ? r : ~r, zigzag(r) = 2 t + (r < 0),
// and the low bit falls off: (2 t + s) >> k == t >> (k - 1)
__device__ __forceinline__ u32 v3_fold_shift(int r, u32 km1)
{
    return ((u32)(r ^ (r >> 31))) >> km1;
}

// FIXED residual of any order in closed form: r[i] = s[i] + c1 s[i-1] + c2 s[i-2] + c3 s[i-3] + c4 s[i-4] with
// the binomial coefficients of the order -- identical modulo 2^32 to the iterated differences of
// flac.c:918-930.  Four multiply-adds on the FMA pipe, and ONE copy of the loop for the five orders.
struct V3FixedCoef { int c1, c2, c3, c4; };
__device__ __forceinline__ V3FixedCoef v3_fixed_coef(u32 order)
{
    V3FixedCoef c;
    c.c1 = -(int)order;
    c.c2 = order < 2 ? 0 : order == 2 ? 1 : order == 3 ? 3 : 6;
    c.c3 = order < 3 ? 0 : order == 3 ? -1 : -4;
    c.c4 = order == 4 ? 1 : 0;
    return c;
}

// Run 0 only: the reference sums the errors from sample 4 on for every order, so samples 0..3 (as v3_fixed_sums
// counted them, with zero history) come out of thread 0's sums again -- g -- while the true residuals |r_k[i]|,
// k <= i < 4, which the partition sums of order k do contain, are collected in c.  Lane k of a warp does order k
// (closed form of the differences, zeros before the block): a fifth of the instructions thread 0 alone spent on it,
// and of the hot code.
__device__ __forceinline__ void v3_fixed_head_lane(const int* __restrict__ samp, u32 k, u32& g, u32& c)
{
    const int4 v = *(const int4*)samp;
    const V3FixedCoef f = v3_fixed_coef(min(k, 4u));
    const int r0 = v.x;
    const int r1 = v.y + f.c1 * v.x;
    const int r2 = v.z + f.c1 * v.y + f.c2 * v.x;
    const int r3 = v.w + f.c1 * v.z + f.c2 * v.y + f.c3 * v.x;
    const u32 a0 = (u32)abs(r0), a1 = (u32)abs(r1), a2 = (u32)abs(r2), a3 = (u32)abs(r3);
    g = a0 + a1 + a2 + a3;
    c = (k == 0 ? a0 : 0u) + (k <= 1 ? a1 : 0u) + (k <= 2 ? a2 : 0u) + (k <= 3 ? a3 : 0u);
}

// sum of (zigzag(r) >> k) over the thread's run for the FIXED residual of `order`, recomputed from the
// samples (the sum fits 32 bits: see the Rice parameter rule, flac.c:1478).  skip: this is run 0, whose
// first `order` positions are warm-up samples and do not count.
// (k >= 1; a Rice parameter of 0 -- a partition of near silence -- takes v3_fixed_bits_k0: one hot copy of the loop)
__device__ __forceinline__ u32 v3_fixed_bits_loop(const int* __restrict__ samp, u32 base, u32 S, u32 km1, const V3FixedCoef c)
{
    int4 h = make_int4(0, 0, 0, 0);                       // samples base-4 .. base-1
    if (base) h = *(const int4*)(samp + V3_SK(base - 4));
    u32 acc = 0;
    V3_LOOP
    for (u32 i0 = base; i0 < base + S; i0 += V3_CH) {
        const int4 va = *(const int4*)(samp + V3_SK(i0));
        const int4 vb = *(const int4*)(samp + V3_SK(i0) + 4);
        const int w[V3_CH + 4] = {h.x, h.y, h.z, h.w, va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
        for (int j = 0; j < V3_CH; j++) {
            const int r = w[j + 4] + c.c1 * w[j + 3] + c.c2 * w[j + 2] + c.c3 * w[j + 1] + c.c4 * w[j];
            acc += v3_fold_shift(r, km1);
        }
        h = vb;
    }
    return acc;
}

// the same sum for k == 0, one sample per iteration: rare, so small code matters more than speed
__device__ __noinline__ u32 v3_fixed_bits_k0(const int* __restrict__ samp, u32 base, u32 S, const V3FixedCoef c)
{
    u32 acc = 0;
#pragma unroll 1
    for (u32 i = base; i < base + S; i++) {
        const int s1 = i >= 1 ? samp[V3_SK(i - 1)] : 0, s2 = i >= 2 ? samp[V3_SK(i - 2)] : 0;
        const int s3 = i >= 3 ? samp[V3_SK(i - 3)] : 0, s4 = i >= 4 ? samp[V3_SK(i - 4)] : 0;
        acc += zigzag(samp[V3_SK(i)] + c.c1 * s1 + c.c2 * s2 + c.c3 * s3 + c.c4 * s4);
    }
    return acc;
}

__device__ __forceinline__ u32 v3_fixed_bits_any(const int* __restrict__ samp, u32 base, u32 S, u32 k, u32 order, u32 skip)
{
    const V3FixedCoef c = v3_fixed_coef(order);
    u32 acc = k ? v3_fixed_bits_loop(samp, base, S, k - 1, c) : v3_fixed_bits_k0(samp, base, S, c);
    if (skip) {
        // run 0: take the warm-up positions back out (they were evaluated with zeros before the block)
#pragma unroll 1
        for (u32 j = 0; j < order; j++) {
            const int s0 = samp[V3_SK(j)];
            const int s1 = j >= 1 ? samp[V3_SK(j - 1)] : 0, s2 = j >= 2 ? samp[V3_SK(j - 2)] : 0;
            const int s3 = j >= 3 ? samp[V3_SK(j - 3)] : 0;
            const int r = s0 + c.c1 * s1 + c.c2 * s2 + c.c3 * s3;     // (s[j-4] is before the block for j < 4)
            acc -= zigzag(r) >> k;
        }
    }
    return acc;
}

// sum of (zigzag(r) >> k) over the thread's run of a stored residual; the first `skip` entries
// (warm-up positions of run 0) are left out
__device__ __forceinline__ u32 v3_stored_bits(const int* __restrict__ resid, u32 base, u32 S, u32 k, u32 skip)
{
    u32 acc = 0;
    if (k) {
        const u32 km1 = k - 1;
        V3_LOOP
        for (u32 i0 = base; i0 < base + S; i0 += V3_CH) {
            const int4 va = *(const int4*)(resid + V3_SK(i0));
            const int4 vb = *(const int4*)(resid + V3_SK(i0) + 4);
            acc += v3_fold_shift(va.x, km1) + v3_fold_shift(va.y, km1) + v3_fold_shift(va.z, km1) + v3_fold_shift(va.w, km1);
            acc += v3_fold_shift(vb.x, km1) + v3_fold_shift(vb.y, km1) + v3_fold_shift(vb.z, km1) + v3_fold_shift(vb.w, km1);
        }
    } else {
#pragma unroll 1
        for (u32 i = base; i < base + S; i++) acc += zigzag(resid[V3_SK(i)]);      // (rare: small code, not speed)
    }
#pragma unroll 1
    for (u32 i = 0; i < skip; i++) acc -= zigzag(resid[V3_SK(base + i)]) >> k;
    return acc;
}

// S: samples per thread (multiple of 8, <= 32); blockDim.x * S == block_size; F: finest partition
// order searched, (block_size >> F) a multiple of S.  Units whose block is not block_size long are
// left to k_analyze_v2 (launched over the same grid, which skips the others).
// EXH: exhaustive order search (flac.c:1070-1120): FIXED as usual, then every LPC order 1..max in turn
// (residual, Rice search, exact bits), keeping the first strict minimum of the exact sizes.
// SUB: run sums per thread run (2: the finest partition is half a run, see V3_MAX_F).
// LONG: LPC orders above 12 may occur (the 32-tap residual loops, ~14 KB of code, are compiled in).  The common
// shapes have an instantiation without them: the hot code of a kernel that never needs them stays in one piece.
template <bool EXH, int SC, int SUB, bool LONG>
__device__ __forceinline__ void v3_unit(unsigned char* dyn_smem, V3SharedT<(SUB == 2 ? 8 : V3_MAX_F)>& sh, u32 unit,
                                        const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd,
                                        const bf_dev_params& P, u32 S_rt, u32 F,
                                        const bf_lpc_head* __restrict__ heads, const short* __restrict__ coefs,
                                        b200flac_plan* __restrict__ plans, uint8_t* __restrict__ rice_out)
{
    const u32 S = SC ? (u32)SC : S_rt;       // samples per thread: a compile-time 32 for the common shapes
    const u32 tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
    const u32 frame = P.K == 4 ? unit >> 2 : unit / P.K, cand = P.K == 4 ? unit & 3u : unit % P.K;
    const bf_frame_desc d = fd[frame];
    const u32 n = d.nsamp;
    if (n != P.block_size) return;
    const u32 bps = candidate_bps(cand, P);
    const u32 base = tid * S;

    const size_t padn = (size_t)V3_SK(n) + 8;
    int* samp = (int*)dyn_smem;
    int* resid = samp + padn;
    const u32 R = nt * SUB;                     // run sums per row: entry SUB * tid + h is half h of thread tid's run
    constexpr bool OWN = EXH && SC == 32;       // run sums in registers, searched by their own warps (v3_search_own)
    u64* runsF = (u64*)(resid + padn);          // [5][R]   (not OWN)
    u64* runsL = OWN ? runsF : runsF + 5 * (size_t)R;         // [R]; OWN: 1 KB of level totals
    V3Level* parts = (V3Level*)runsL;                        // OWN: [2][4][V3_OWN_LEVELS], by order parity
    u64* wsum = (u64*)(parts + 2 * 4 * V3_OWN_LEVELS);       // OWN: [2][4]

    // ---- LPC model of the unit (last warp; overlaps the PCM load of the others) ----
    const short* mycoef = coefs + (size_t)unit * P.model_stride;
    if (warp == nw - 1) {
        // (36 bytes of byte-sized fields: copied as nine words, one per lane -- the struct assignment was 87 instructions)
        static_assert(sizeof(bf_lpc_head) == 36, "bf_lpc_head is copied as nine 32-bit words");
        if (lane < 9) ((u32*)&sh.head)[lane] = ((const u32*)(heads + unit))[lane];
        __syncwarp();
        if (!EXH) {
            const u32 o = sh.head.best_order;
            const int q = lane < o ? (int)mycoef[(o * (o - 1)) / 2 + lane] : 0;
            sh.q[lane] = (short)q;
            const u32 sumq = __reduce_add_sync(0xFFFFFFFFu, (u32)abs(q));
            if (lane == 0) sh.lpc_narrow = sumq;     // turned into the flag once wasted bits are known
        }
    }
    if (tid < 5) { sh.totF[tid] = 0ull; sh.totF16[tid][0] = 0u; sh.totF16[tid][1] = 0u; }
    if (tid < 4) sh.bits16[tid >> 1][tid & 1] = 0u;

    // ---- load, constant check, wasted bits (flac.c:691-724) ----
    u32 orv = 0, diff = 0;
    {
        const int first = ld_candidate(pcm, d.pcm_off, cand, P);
        if (P.stereo && P.bytes_ps == 2) {
            const int coef = cand == 0 ? 0x0001 : cand == 1 ? 0x0100 : cand == 2 ? 0x0101 : 0xFF01;
            const int shv = cand == 2 ? 1 : 0;
            const uint8_t* src = pcm + d.pcm_off * 4;
            if ((((uintptr_t)src) & 15) == 0) {
                const uint4* s4 = (const uint4*)src;
                // all the loads of a batch are issued before the first is used: one memory latency
                // per batch of 4 x 16 bytes, not per load (batches of 8, unrolled, were 100 instructions more
                // of hot code for the same time)
                const u32 stride = nt * 4;
                u32 i = tid * 4;
                V3_LOOP
                for (; i + 3 * stride < n; i += 4 * stride) {
                    uint4 w8[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) w8[q] = __ldg(s4 + ((i + q * stride) >> 2));
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        int4 v;
                        v.x = __dp2a_lo((int)w8[q].x, coef, 0) >> shv; v.y = __dp2a_lo((int)w8[q].y, coef, 0) >> shv;
                        v.z = __dp2a_lo((int)w8[q].z, coef, 0) >> shv; v.w = __dp2a_lo((int)w8[q].w, coef, 0) >> shv;
                        *(int4*)(samp + V3_SK(i + q * stride)) = v;
                        orv |= (u32)(v.x | v.y | v.z | v.w);
                        diff |= (u32)((v.x ^ first) | (v.y ^ first) | (v.z ^ first) | (v.w ^ first));
                    }
                }
                for (; i < n; i += stride) {
                    const uint4 w4 = __ldg(s4 + (i >> 2));
                    int4 v;
                    v.x = __dp2a_lo((int)w4.x, coef, 0) >> shv; v.y = __dp2a_lo((int)w4.y, coef, 0) >> shv;
                    v.z = __dp2a_lo((int)w4.z, coef, 0) >> shv; v.w = __dp2a_lo((int)w4.w, coef, 0) >> shv;
                    *(int4*)(samp + V3_SK(i)) = v;
                    orv |= (u32)(v.x | v.y | v.z | v.w);
                    diff |= (u32)((v.x ^ first) | (v.y ^ first) | (v.z ^ first) | (v.w ^ first));
                }
            } else {
                const u32* s1 = (const u32*)src;
                for (u32 i = tid; i < n; i += nt) {
                    const int v = __dp2a_lo((int)__ldg(s1 + i), coef, 0) >> shv;
                    samp[V3_SK(i)] = v;
                    orv |= (u32)v; diff |= (u32)(v ^ first);
                }
            }
        } else if (P.stereo && P.bytes_ps == 3 && (n & 7u) == 0 && (((uintptr_t)(pcm + d.pcm_off * 6)) & 15) == 0) {
            load_stereo24_skewed(pcm + d.pcm_off * 6, samp, n, cand, tid, nt, first, &orv, &diff);
        } else {
            for (u32 i = tid; i < n; i += nt) {
                const int v = ld_candidate(pcm, d.pcm_off + i, cand, P);
                samp[V3_SK(i)] = v;
                orv |= (u32)v; diff |= (u32)(v ^ first);
            }
        }
    }
    orv = __reduce_or_sync(0xFFFFFFFFu, orv);
    diff = __reduce_or_sync(0xFFFFFFFFu, diff);
    if (lane == 0) { sh.red_or[warp] = orv; sh.red_diff[warp] = diff; }
    __syncthreads();                                                             // (1)
    orv = __reduce_or_sync(0xFFFFFFFFu, lane < nw ? sh.red_or[lane] : 0u);
    diff = __reduce_or_sync(0xFFFFFFFFu, lane < nw ? sh.red_diff[lane] : 0u);

    if (diff == 0) {
        // CONSTANT, always written with wasted = 0 (H8)
        if (tid == 0) {
            b200flac_plan plan;
            memset(&plan, 0, sizeof(plan));
            plan.type = BF_CONSTANT; plan.bits = 8 + bps;
            plans[unit] = plan;
        }
        return;
    }
    const u32 wasted = orv ? (u32)(__ffs((int)orv) - 1) : 0u;
    if (wasted) {
        for (u32 i = tid; i < n; i += nt) samp[V3_SK(i)] >>= wasted;                 // arithmetic (H10)
        __syncthreads();
    }
    const u32 sub_bps = bps - wasted;
    const u32 hdr_bits = 8 + wasted;

    // ---- pass A: FIXED sums of all orders + LPC residual, run sums to shared memory ----
    u64 ownE[5] = {0, 0, 0, 0, 0}, ownH[5] = {0, 0, 0, 0, 0};   // OWN: FIXED sums of the run (second half, SUB = 2) / first half
    {
        u64* mine = runsF + SUB * tid;             // this thread's (first) entry of order 0's row
        if constexpr (OWN) {
            if (sub_bps <= 23) {
                u32 e[5];
                v3_fixed_sums<u32, SUB>(samp, base, S, e, ownH, 1);
#pragma unroll
                for (int k = 0; k < 5; k++) ownE[k] = (u64)e[k];
            } else {
                v3_fixed_sums<u64, SUB>(samp, base, S, ownE, ownH, 1);
            }
            if (warp == 0) {
                u32 g, c;
                v3_fixed_head_lane(samp, lane, g, c);
                if (lane < 5) sh.corr[lane] = (u64)c;
#pragma unroll
                for (int k = 0; k < 5; k++) {
                    const u32 gk = __shfl_sync(0xFFFFFFFFu, g, k);
                    if (tid == 0) { if (SUB == 2) ownH[k] -= (u64)gk; else ownE[k] -= (u64)gk; }
                }
            }
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const u64 mysum = ownE[k] + (SUB == 2 ? ownH[k] : 0ull);
                if (sub_bps <= 23) {
                    const u32 ws = __reduce_add_sync(0xFFFFFFFFu, (u32)mysum);
                    if (lane == 0) { atomicAdd(&sh.totF16[k][0], ws & 0xFFFFu); atomicAdd(&sh.totF16[k][1], ws >> 16); }
                } else {
                    u64 ws = mysum;
#pragma unroll
                    for (int o = 16; o; o >>= 1) ws += __shfl_xor_sync(0xFFFFFFFFu, ws, o);
                    if (lane == 0) atomicAdd(&sh.totF[k], ws);
                }
            }
        } else {
        if (sub_bps <= 23) {
            u32 e[5];
            v3_fixed_sums<u32, SUB>(samp, base, S, e, mine, R);
#pragma unroll
            for (int k = 0; k < 5; k++) mine[k * R + (SUB - 1)] = (u64)e[k];
        } else {
            u64 e[5];
            v3_fixed_sums<u64, SUB>(samp, base, S, e, mine, R);
#pragma unroll
            for (int k = 0; k < 5; k++) mine[k * R + (SUB - 1)] = e[k];
        }
        if (warp == 0) {
            __syncwarp();                   // thread 0's sums are in shared memory
            u32 g, c;
            v3_fixed_head_lane(samp, lane, g, c);
            if (lane < 5) { runsF[lane * R] -= (u64)g; sh.corr[lane] = (u64)c; }
            __syncwarp();                   // ... and corrected before it reads them back
        }
        // block totals; each thread reads its own sums back, one rolled copy of the reduction for the five orders
        if (sub_bps <= 23) {
            // 32 runs x 2^26 fit 32 bits; the two halves are summed over <= 16 warps with native
            // 32-bit shared atomics (a 64-bit one is a compare-and-swap loop)
#pragma unroll 1
            for (int k = 0; k < 5; k++) {
                const u32 ws = __reduce_add_sync(0xFFFFFFFFu, (u32)mine[k * R] + (SUB == 2 ? (u32)mine[k * R + 1] : 0u));
                if (lane == 0) { atomicAdd(&sh.totF16[k][0], ws & 0xFFFFu); atomicAdd(&sh.totF16[k][1], ws >> 16); }
            }
        } else {
#pragma unroll 1
            for (int k = 0; k < 5; k++) {
                u64 ws = mine[k * R] + (SUB == 2 ? mine[k * R + 1] : 0ull);
#pragma unroll
                for (int o = 16; o; o >>= 1) ws += __shfl_xor_sync(0xFFFFFFFFu, ws, o);
                if (lane == 0) atomicAdd(&sh.totF[k], ws);
            }
        }
        }
    }
    u32 lpc_order = sh.head.best_order;
    const u32 precision = sh.head.precision;
    // "narrow": the prediction sum fits 32 bits (sum of |coefficient| * 2^(sub_bps - 1) < 2^31) AND every |residual|
    // stays below 2^29 (|prediction| <= 2^31 >> shift with shift >= 3, |sample| < 2^27), so that eight of them add up
    // in 32 bits (v3_lpc_residual sums a chunk with VABSDIFF).  Anything else takes the wide paths.
    auto is_narrow = [&](u32 sum_abs_q, int shift) -> bool {
        return ((u64)sum_abs_q << (sub_bps - 1)) < (1ull << 31) && shift >= 3 && sub_bps <= 28;
    };
    bool lpc_narrow = false;
    int lpc_shift = 0;
    // residual of order o with the coefficients staged in qs; s0 (and s1, SUB = 2) get the thread's run sum(s)
    // the FP64 pass (v3_lpc_residual_f64) takes wide sums of up to 12 taps: sub_bps <= 26 and |coefficient| < 2^15 keep
    // every partial sum below 2^45; |prediction| <= sum|q| * 2^(sub_bps-1) >> shift < 2^27 keeps every |residual|
    // below 2^28, so that it, too, may add eight of them up in 32 bits
    auto is_f64 = [&](u32 sum_abs_q, int shift) -> bool {
        return sub_bps <= 26 && shift >= 0 && (((u64)sum_abs_q << (sub_bps - 1)) >> shift) < (1ull << 27);
    };
    auto lpc_pass = [&](u32 o, int shift, bool narrow, u32 sum_abs_q, const short* qs, u64& s0, u64& s1) {
        u64 run;
        u64 first = 0;
        u64* h0 = &first;
        // (no 8-tap variant: a second hot copy of the residual loop costs more in instruction fetch than
        // the four extra multiply-adds of a padded low order cost on the otherwise idle FMA pipe)
        // (the exhaustive search walks every order, so there the 8-tap copy pays for itself)
        // wide sums of up to 12 taps go through the FP64 pipe (exact: see v3_lpc_residual_f64 and is_f64)
        const bool f64 = !narrow && is_f64(sum_abs_q, shift);
        if (EXH && o <= 8) run = narrow ? v3_lpc_residual<8, V3_ACC_I32, SUB, (SC == 32 && SUB == 1 ? 2 : 1)>(samp, resid, base, S, qs, shift, h0)
                                 : f64 ? v3_lpc_residual_f64<8, SUB>(samp, resid, base, S, qs, shift, h0)
                                       : v3_lpc_residual<12, V3_ACC_I64, SUB>(samp, resid, base, S, qs, shift, h0);
        else if (o <= 12) run = narrow ? v3_lpc_residual<12, V3_ACC_I32, SUB, (EXH && SC == 32 && SUB == 1 ? 2 : 1)>(samp, resid, base, S, qs, shift, h0)
                                : f64 ? v3_lpc_residual_f64<12, SUB>(samp, resid, base, S, qs, shift, h0)
                                      : v3_lpc_residual<12, V3_ACC_I64, SUB>(samp, resid, base, S, qs, shift, h0);
        else if constexpr (LONG) run = narrow ? v3_lpc_residual<32, V3_ACC_I32, SUB>(samp, resid, base, S, qs, shift, h0)
                                              : v3_lpc_residual<32, V3_ACC_I64, SUB>(samp, resid, base, S, qs, shift, h0);
        else run = 0;       // (unreachable: the host launches this instantiation for max_lpc_order <= 12 only)
        if (SUB == 2) { s0 = first; s1 = run; } else { s0 = run; s1 = 0; }
        if (warp == 0) {
            // the warm-up positions (o <= 32 / SUB, all in thread 0's run) do not count: one lane each instead of a
            // loop in thread 0 that every other warp of the CTA ends up waiting for
            __syncwarp();
            const u32 a = lane < o ? (u32)abs(resid[V3_SK(lane)]) : 0u;
            const u64 warm = (u64)__reduce_add_sync(0xFFFFFFFFu, a & 0xFFFFu) + ((u64)__reduce_add_sync(0xFFFFFFFFu, a >> 16) << 16);
            if (tid == 0) s0 -= warm;
        }
    };
    auto store_runs = [&](u64 s0, u64 s1) { runsL[SUB * tid] = s0; if (SUB == 2) runsL[2 * tid + 1] = s1; };
    if (!EXH) {
        lpc_shift = sh.head.shift[lpc_order - 1];
        lpc_narrow = is_narrow(sh.lpc_narrow, lpc_shift);
        u64 s0, s1;
        lpc_pass(lpc_order, lpc_shift, lpc_narrow, sh.lpc_narrow, sh.q, s0, s1);
        store_runs(s0, s1);
    }
    __syncthreads();                                                             // (2)

    // ---- FIXED order: first strict minimum of the block totals (flac.c:877-893) ----
    u32 fixed_order = 0;
    {
        u64 best = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const u64 t = sh.totF[k] + (u64)sh.totF16[k][0] + ((u64)sh.totF16[k][1] << 16);
            if (k == 0 || t < best) { best = t; fixed_order = k; }
        }
    }
    // ---- the two Rice searches as four warp tasks (model x part, see v3_levels); the tasks rotate
    // over the warps with the unit so that no scheduler always gets the extra work ----
    if constexpr (OWN) {
        auto sel5 = [&](const u64 (&a)[5]) -> u64 {
            return fixed_order == 0 ? a[0] : fixed_order == 1 ? a[1] : fixed_order == 2 ? a[2] : fixed_order == 3 ? a[3] : a[4];
        };
        u64 s0 = SUB == 2 ? sel5(ownH) : sel5(ownE);
        const u64 s1 = SUB == 2 ? sel5(ownE) : 0ull;
        if (tid == 0) s0 += sh.corr[fixed_order];     // |r[i]| of order <= i < 4, which the partition sums do contain
        v3_search_own<SUB>(s0, s1, fixed_order, F, n, P.max_rice, sh.kheap[0], parts + warp * V3_OWN_LEVELS, wsum + warp);
    } else {
        const u32 role = (nw & (nw - 1)) == 0 ? ((warp - unit) & (nw - 1)) : (warp + nw - unit % nw) % nw;
        for (u32 task = role; task < (EXH ? 2u : 4u); task += nw) {
            // prefix sums go to the run-sum rows of two FIXED orders that lost; one call site for both models
            const bool lpc = task >= 2;
            v3_levels(lpc ? runsL : runsF + (size_t)fixed_order * R, lpc ? 0ull : sh.corr[fixed_order], S / SUB, n,
                      lpc ? lpc_order : fixed_order, F, P.max_rice, sh.kheap[lpc ? 1 : 0], sh.lvl[lpc ? 1 : 0],
                      runsF + (size_t)((fixed_order + (lpc ? 2u : 1u)) % 5) * R, task & 1u);
        }
    }
    __syncthreads();                                                             // (3)

    // ---- pass B: exact bits of both models ----
    // finest partition this thread's run lies in (its partition at order po is pF >> (F - po)); SUB = 2: the
    // run is the finest partitions pF and pF + 1, and has two Rice parameters when the finest order is chosen
    const u32 gruns = SUB == 2 ? 1u : (n >> F) / S;
    const u32 pF = SUB == 2 ? 2u * tid : (gruns & (gruns - 1)) == 0 ? tid >> (31 - __clz((int)gruns)) : tid / gruns;
    // exact bits of the thread's run for a model whose Rice parameters are kh[] at partition order po
    auto run_bits = [&](const uint8_t* kh, u32 po, bool fixed, u32 order) -> u32 {
        const u32 nseg = (SUB == 2 && po == F) ? 2u : 1u, len = S / nseg;
        u32 acc = 0;
#pragma unroll 1
        for (u32 h = 0; h < nseg; h++) {
            const u32 k = kh[(1u << po) - 1u + ((pF + h) >> (F - po))];
            const u32 skip = (tid == 0 && h == 0) ? 1u : 0u;
            acc += fixed ? v3_fixed_bits_any(samp, base + h * len, len, k, order, skip)
                         : v3_stored_bits(resid, base + h * len, len, k, 0u);
        }
        if (!fixed && warp == 0) {
            // LPC: the warm-up positions (all in partition 0, thread 0's run) were counted above; take them back
            // out, one lane each (thread 0 alone looped over them while the CTA waited)
            const u32 k0 = kh[(1u << po) - 1u];
            const u32 v = lane < order ? (zigzag(resid[V3_SK(lane)]) >> k0) : 0u;
            const u32 sub = __reduce_add_sync(0xFFFFFFFFu, v);
            if (tid == 0) acc -= sub;
        }
        return acc;
    };
    u32 poF, poL = 0, methodF, methodL = 0;
    u64 sideF, sideL = 0;
    if constexpr (OWN) v3_pick_own(parts, wsum, fixed_order, F, n, P.max_rice, sh.kheap[0], &poF, &methodF, &sideF);
    else v3_pick_level(sh.lvl[0], F, &poF, &methodF, &sideF);
    if (!EXH) v3_pick_level(sh.lvl[1], F, &poL, &methodL, &sideL);
    {
        // a warp's sum stays far below 2^32 (each run's is bounded by ~2 * partition length + 32 * 2^18)
        const u32 bF = __reduce_add_sync(0xFFFFFFFFu, run_bits(sh.kheap[0], poF, true, fixed_order));
        if (lane == 0) { atomicAdd(&sh.bits16[0][0], bF & 0xFFFFu); atomicAdd(&sh.bits16[0][1], bF >> 16); }
        if (!EXH) {
            const u32 bL = __reduce_add_sync(0xFFFFFFFFu, run_bits(sh.kheap[1], poL, false, lpc_order));
            if (lane == 0) { atomicAdd(&sh.bits16[1][0], bL & 0xFFFFu); atomicAdd(&sh.bits16[1][1], bL >> 16); }
        }
    }
    __syncthreads();                                                             // (4)

    u64 lpc_bits = 0;
    if (!EXH) {
        lpc_bits = hdr_bits + (u64)lpc_order * sub_bps + 4 + 5 + (u64)lpc_order * precision + sideL +
                   (u64)sh.bits16[1][0] + ((u64)sh.bits16[1][1] << 16);
    } else {
        // ---- every LPC order in turn; `unsigned best_bits = UINT_MAX`, strict <, ascending (flac.c:1079-1108) ----
        // Two block-wide barriers per order (after the residual pass, after the Rice search): the next order's
        // coefficients are staged while this one runs, and an order's exact size is looked at one iteration later,
        // after the barrier that follows the next residual pass -- Rice parameters and bit sums of two orders are
        // in flight, by order parity.  (It was four barriers per order; barriers were the top stall, 3.1 per issue.)
        u32 best32 = 0xFFFFFFFFu;
        bool have = false;
        const u32 role = (nw & (nw - 1)) == 0 ? ((warp - unit) & (nw - 1)) : (warp + nw - unit % nw) % nw;
        const u32 L = P.max_lpc_order;
        auto stage = [&](u32 o) {            // warp nw - 1: coefficients of order o -> q2[o & 1]
            const int q = lane < o ? (int)mycoef[(o * (o - 1)) / 2 + lane] : 0;
            sh.q2[o & 1][lane] = (short)q;
            const u32 sumq = __reduce_add_sync(0xFFFFFFFFu, (u32)abs(q));
            if (lane == 0) sh.lpc_narrow2[o & 1] = sumq;
        };
        // what the iteration before left to be decided
        u32 pend_po = 0, pend_method = 0;
        u64 pend_side = 0;
        int pend_shift = 0;
        bool pend_narrow = false;
        auto decide = [&](u32 o) {           // order o's exact size is complete: first strict minimum
            const u64 bits = hdr_bits + (u64)o * sub_bps + 4 + 5 + (u64)o * precision + pend_side +
                             (u64)sh.bitsL[o & 1][0] + ((u64)sh.bitsL[o & 1][1] << 16);
            if (!have || (u32)bits < best32) {
                have = true;
                best32 = (u32)bits;
                lpc_bits = bits; lpc_order = o; lpc_shift = pend_shift; lpc_narrow = pend_narrow;
                poL = pend_po; methodL = pend_method;
                const u32 koff = (1u << pend_po) - 1u;
                for (u32 p = tid; p < (1u << pend_po); p += nt) sh.kbest[p] = sh.kheapL[o & 1][koff + p];
            }
        };
        if constexpr (SC == 32) {
            // 128 threads x 32 samples: ONE barrier per order.  The run sums stay in registers and every warp searches the
            // partitions under its own runs before the barrier (v3_search_own); after it every warp adds the four
            // warps' level totals up, picks the order (v3_pick_own) and counts the exact bits of its runs.  The residual
            // buffer is only ever read and written by the thread that owns the run, so the next pass may start at once.
            // Rice parameters and bit sums of three orders are in flight (o mod 3): order o's are written before and
            // after barrier o, read by decide(o) after barrier o + 1, while order o + 1's search is already writing.
            auto kbuf = [&](u32 o) -> uint8_t* { const u32 m = o % 3u; return m == 0 ? sh.kheapL[0] : m == 1 ? sh.kheapL[1] : sh.kheap[1]; };
            auto bbuf = [&](u32 o) -> u32* { const u32 m = o % 3u; return m == 0 ? sh.bitsL[0] : m == 1 ? sh.bitsL[1] : sh.bits16[1]; };
            auto decide3 = [&](u32 o) {
                const u32* b = bbuf(o);
                const u64 bits = hdr_bits + (u64)o * sub_bps + 4 + 5 + (u64)o * precision + pend_side + (u64)b[0] + ((u64)b[1] << 16);
                if (!have || (u32)bits < best32) {
                    have = true;
                    best32 = (u32)bits;
                    lpc_bits = bits; lpc_order = o; lpc_shift = pend_shift; lpc_narrow = pend_narrow;
                    poL = pend_po; methodL = pend_method;
                    const u32 koff = (1u << pend_po) - 1u;
                    const uint8_t* kb = kbuf(o);
                    for (u32 p = tid; p < (1u << pend_po); p += nt) sh.kbest[p] = kb[koff + p];
                }
            };
            if (tid < 4) sh.bitsL[tid >> 1][tid & 1] = 0u;
            if (tid < 2) sh.bits16[1][tid] = 0u;
            if (warp == nw - 1) stage(1);
            __syncthreads();
#pragma unroll 1
            for (u32 o = 1; o <= L; o++) {
                const int shift = sh.head.shift[o - 1];
                const bool narrow = is_narrow(sh.lpc_narrow2[o & 1], shift);
                u64 s0, s1;
                lpc_pass(o, shift, narrow, sh.lpc_narrow2[o & 1], sh.q2[o & 1], s0, s1);
                if (warp == nw - 1 && o < L) stage(o + 1);      // (its buffer was last read by order o - 1's pass)
                uint8_t* kb = kbuf(o);
                v3_search_own<SUB>(s0, s1, o, F, n, P.max_rice, kb, parts + ((o & 1u) * 4u + warp) * V3_OWN_LEVELS,
                                   wsum + (o & 1u) * 4u + warp);
                __syncthreads();
                if (o > 1) decide3(o - 1);
                if (tid < 2) bbuf(o + 1)[tid] = 0u;             // last read by decide(o - 2), before this barrier
                u64 side;
                v3_pick_own(parts + (o & 1u) * 4u * V3_OWN_LEVELS, wsum + (o & 1u) * 4u, o, F, n, P.max_rice, kb,
                            &pend_po, &pend_method, &side);
                pend_side = side; pend_shift = shift; pend_narrow = narrow;
                const u32 bL = __reduce_add_sync(0xFFFFFFFFu, run_bits(kb, pend_po, false, o));
                u32* bb = bbuf(o);
                if (lane == 0) { atomicAdd(&bb[0], bL & 0xFFFFu); atomicAdd(&bb[1], bL >> 16); }
            }
            __syncthreads();
            decide3(L);
            __syncthreads();
        } else {
        if (tid < 4) sh.bitsL[tid >> 1][tid & 1] = 0u;
        if (warp == nw - 1) stage(1);
        __syncthreads();
#pragma unroll 1
        for (u32 o = 1; o <= L; o++) {
            const int shift = sh.head.shift[o - 1];
            const bool narrow = is_narrow(sh.lpc_narrow2[o & 1], shift);
            u64 s0, s1;
            lpc_pass(o, shift, narrow, sh.lpc_narrow2[o & 1], sh.q2[o & 1], s0, s1);
            store_runs(s0, s1);
            if (warp == nw - 1 && o < L) stage(o + 1);      // (its buffer was last read by order o - 1's pass)
            __syncthreads();
            if (o > 1) decide(o - 1);
            // the Rice search of the order as four warp tasks (both halves of the finest order, orders 0..4, orders 5..)
            for (u32 task = role; task < 4; task += nw)
                v3_levels(runsL, 0ull, S / SUB, n, o, F, P.max_rice, sh.kheapL[o & 1], sh.lvl[1],
                          runsF + (size_t)((fixed_order + 2 + (task & 1u)) % 5) * R, 2u + task);
            __syncthreads();
            if (tid < 2) sh.bitsL[(o + 1) & 1][tid] = 0u;   // order o - 1's sum has been read by everyone
            u64 side;
            v3_pick_level(sh.lvl[1], F, &pend_po, &pend_method, &side, true);
            pend_side = side; pend_shift = shift; pend_narrow = narrow;
            const u32 bL = __reduce_add_sync(0xFFFFFFFFu, run_bits(sh.kheapL[o & 1], pend_po, false, o));
            if (lane == 0) { atomicAdd(&sh.bitsL[o & 1][0], bL & 0xFFFFu); atomicAdd(&sh.bitsL[o & 1][1], bL >> 16); }
        }
        __syncthreads();
        decide(L);
        __syncthreads();
        }
    }

    // ---- choice, flac.c:727-809 (every subframe type enabled) ----
    const u64 fixed_bits = hdr_bits + (u64)fixed_order * sub_bps + sideF + (u64)sh.bits16[0][0] + ((u64)sh.bits16[0][1] << 16);
    const u32 fb = (u32)fixed_bits, lb = (u32)lpc_bits;
    const u32 vb = sub_bps * n;                       // header NOT counted (H2)
    const u32 choice = (fb < min(lb, vb)) ? BF_FIXED : (lb < vb) ? BF_LPC : BF_VERBATIM;
    uint8_t* my_rice = rice_out + (size_t)unit * P.rice_stride;
    if (choice == BF_FIXED) {
        const u32 koff = (1u << poF) - 1u;
        for (u32 p = tid; p < (1u << poF); p += nt) my_rice[p] = sh.kheap[0][koff + p];
    } else if (choice == BF_LPC) {
        const u32 koff = (1u << poL) - 1u;
        for (u32 p = tid; p < (1u << poL); p += nt) my_rice[p] = EXH ? sh.kbest[p] : sh.kheap[1][koff + p];
    }
    // the plan goes out as 19 words, one per thread: two of byte fields, the size, sixteen pairs of coefficients
    // (the struct assignment it replaces was ~100 instructions for thread 0)
    static_assert(sizeof(b200flac_plan) == 76 && B200FLAC_MAX_LPC_ORDER == 32, "b200flac_plan is written as 19 words");
    if (tid < 19) {
        u32 word = 0;
        const bool is_lpc = choice == BF_LPC, is_fixed = choice == BF_FIXED;
        if (tid == 0)      // type | order | wasted | precision
            word = choice | ((is_fixed ? fixed_order : is_lpc ? lpc_order : 0u) << 8) | (wasted << 16) | ((is_lpc ? precision : 0u) << 24);
        else if (tid == 1) // shift | coding method | partition order | flags (bit 1: the packer may accumulate in 32 bits)
            word = (is_lpc ? ((u32)lpc_shift & 0xFFu) : 0u) | ((is_fixed ? methodF : is_lpc ? methodL : 0u) << 8) |
                   ((is_fixed ? poF : is_lpc ? poL : 0u) << 16) | ((is_lpc && lpc_narrow ? 2u : 0u) << 24);
        else if (tid == 2) // exact size; VERBATIM: flac.c:832-854
            word = is_fixed ? fb : is_lpc ? lb : hdr_bits + sub_bps * n;
        else if (is_lpc) {
            const u32 j = 2 * (tid - 3);
            const short c0 = j < lpc_order ? (EXH ? mycoef[(lpc_order * (lpc_order - 1)) / 2 + j] : sh.q[j]) : (short)0;
            const short c1 = j + 1 < lpc_order ? (EXH ? mycoef[(lpc_order * (lpc_order - 1)) / 2 + j + 1] : sh.q[j + 1]) : (short)0;
            word = (u32)(unsigned short)c0 | ((u32)(unsigned short)c1 << 16);
        }
        ((u32*)(plans + unit))[tid] = word;
    }
}

// The grid is normally one CTA per unit; any smaller grid walks the units with the grid's stride
// (measured: a persistent single wave keeps the CTAs of an SM in the same phase of the unit, which
// overlaps their load and search phases worse than staggered CTAs do).
template <int MINB, bool EXH, int SC, int SUB = 1, bool LONG = true>
__global__ void __launch_bounds__(MINB >= 4 ? 128 : MINB >= 3 ? 256 : 512, MINB)
k_analyze_v3(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd, bf_dev_params P, u32 S, u32 F,
             u32 n_units, const bf_lpc_head* __restrict__ heads, const short* __restrict__ coefs,
             b200flac_plan* __restrict__ plans, uint8_t* __restrict__ rice_out)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ V3SharedT<(SUB == 2 ? 8 : V3_MAX_F)> sh;
    for (u32 unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
        v3_unit<EXH, SC, SUB, LONG>(dyn_smem, sh, unit, pcm, fd, P, S, F, heads, coefs, plans, rice_out);
        __syncthreads();        // shared memory is reused by the next unit
    }
}
