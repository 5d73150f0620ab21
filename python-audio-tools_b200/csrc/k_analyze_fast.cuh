// k_analyze_fast.cuh -- register-resident version of k_analyze for the common
// shapes: the whole block fits one pass (n <= blockDim * S) and shared memory.
//
// Same decisions, bit for bit, as k_analyze (k_analyze.cuh) and therefore as the
// reference (flac.c:673-1505); only the data movement differs:
//   * every thread keeps its S contiguous samples (after the wasted-bits shift)
//     in registers for the whole kernel; shared memory holds a copy (one pad word
//     per thread, so runs are bank-conflict free) only for the predictor history;
//   * FIXED differences, the LPC multiply-accumulate and both residual passes of
//     the Rice search are fully unrolled over those registers;
//   * the LPC accumulate runs in 32 bits whenever sum|q| * 2^(bps-1) < 2^31
//     proves the 64-bit sum of flac.c:1000-1005 cannot leave int32 (always true
//     for 16-bit input at order <= 15), in 64 bits otherwise;
//   * the partition-order decision is data-parallel: a prefix sum over the finest
//     partition sums gives every (level, partition) sum as a difference, one thread
//     per partition computes the Rice parameter in closed form (falling back to the
//     reference's loop where its 32-bit wrap could matter) and its size estimate.
#pragma once
#include "flac_common.cuh"
#include "k_analyze.cuh"

// shared-memory index of sample i when every thread owns S samples: one pad word per thread
template <int S>
__device__ __forceinline__ u32 pad_idx(u32 i) { return i + i / (u32)S; }

// ints of shared memory needed for a block of bs samples in that layout
static inline size_t fast_samp_ints(unsigned bs, int S) { return (size_t)bs + (bs + S - 1) / S + 2; }

// order groups: coefficients are zero-padded to the group size (exact: 0 * x adds nothing)
// (few groups on purpose: the multiply-accumulate is not the bottleneck, instruction fetch is)
__device__ __forceinline__ int order_group(u32 o) { return o <= 12 ? 12 : 32; }

// r[j] = s[j] - (int)((sum_t q[t] * x[j-1-t]) >> shift), x = own samples then history from smem
template <int S, int OG, bool WIDE>
__device__ __forceinline__ void lpc_residual_regs(const int (&s)[S], const int* samp, u32 base, u32 n,
                                                  const short* q_sm, int shift, int (&r)[S])
{
    int q[OG], h[OG];
#pragma unroll
    for (int t = 0; t < OG; t++) {
        q[t] = q_sm[t];
        const int idx = (int)base - 1 - t;
        h[t] = (idx >= 0 && idx < (int)n) ? samp[pad_idx<S>((u32)idx)] : 0; // idle threads must not read past the block
    }
#pragma unroll
    for (int j = 0; j < S; j++) {
        if (WIDE) {
            long long acc = 0;
#pragma unroll
            for (int t = 0; t < OG; t++) {
                const int x = (j - 1 - t >= 0) ? s[(j - 1 - t >= 0) ? (j - 1 - t) : 0] : h[(t - j >= 0 && t - j < OG) ? (t - j) : 0];
                acc += (long long)q[t] * (long long)x;
            }
            acc >>= shift;
            r[j] = (int)((u32)s[j] - (u32)(int)acc);
        } else {
            int acc = 0;
#pragma unroll
            for (int t = 0; t < OG; t++) {
                const int x = (j - 1 - t >= 0) ? s[(j - 1 - t >= 0) ? (j - 1 - t) : 0] : h[(t - j >= 0 && t - j < OG) ? (t - j) : 0];
                acc += q[t] * x;
            }
            acc >>= shift;
            r[j] = (int)((u32)s[j] - (u32)acc);
        }
    }
}

template <int S, bool WIDE>
__device__ __forceinline__ void lpc_residual_dispatch(const int (&s)[S], const int* samp, u32 base, u32 n, u32 order,
                                                      const short* q_sm, int shift, int (&r)[S])
{
    if (order_group(order) == 12) lpc_residual_regs<S, 12, WIDE>(s, samp, base, n, q_sm, shift, r);
    else lpc_residual_regs<S, 32, WIDE>(s, samp, base, n, q_sm, shift, r);
}

// FIXED residual of ORDER for the thread's run (closed forms of the iterated differences)
template <int S, int ORDER>
__device__ __forceinline__ void fixed_residual_order(const int (&s)[S], u32 h1, u32 h2, u32 h3, u32 h4, int (&r)[S])
{
#pragma unroll
    for (int j = 0; j < S; j++) {
        const u32 s0 = (u32)s[j];
        const u32 s1 = j >= 1 ? (u32)s[j >= 1 ? j - 1 : 0] : h1;
        const u32 s2 = j >= 2 ? (u32)s[j >= 2 ? j - 2 : 0] : (j == 1 ? h1 : h2);
        const u32 s3 = j >= 3 ? (u32)s[j >= 3 ? j - 3 : 0] : (j == 2 ? h1 : j == 1 ? h2 : h3);
        const u32 s4 = j >= 4 ? (u32)s[j >= 4 ? j - 4 : 0] : (j == 3 ? h1 : j == 2 ? h2 : j == 1 ? h3 : h4);
        u32 v;
        if (ORDER == 0) v = s0;
        else if (ORDER == 1) v = s0 - s1;
        else if (ORDER == 2) v = s0 - 2u * s1 + s2;
        else if (ORDER == 3) v = s0 - 3u * s1 + 3u * s2 - s3;
        else v = s0 - 4u * s1 + 6u * s2 - 4u * s3 + s4;
        r[j] = (int)v;
    }
}

template <int S>
__device__ __forceinline__ void fixed_residual_regs(const int (&s)[S], const int* samp, u32 base, u32 n, u32 order, int (&r)[S])
{
    u32 h1 = 0, h2 = 0, h3 = 0, h4 = 0;
    if (base >= 1 && base - 1 < n) h1 = (u32)samp[pad_idx<S>(base - 1)];
    if (base >= 2 && base - 2 < n) h2 = (u32)samp[pad_idx<S>(base - 2)];
    if (base >= 3 && base - 3 < n) h3 = (u32)samp[pad_idx<S>(base - 3)];
    if (base >= 4 && base - 4 < n) h4 = (u32)samp[pad_idx<S>(base - 4)];
    switch (order) {
    case 0: fixed_residual_order<S, 0>(s, h1, h2, h3, h4, r); break;
    case 1: fixed_residual_order<S, 1>(s, h1, h2, h3, h4, r); break;
    case 2: fixed_residual_order<S, 2>(s, h1, h2, h3, h4, r); break;
    case 3: fixed_residual_order<S, 3>(s, h1, h2, h3, h4, r); break;
    default: fixed_residual_order<S, 4>(s, h1, h2, h3, h4, r); break;
    }
}

// zero the residuals outside [order, n) so that the sums below need no per-sample predicate
template <int S>
__device__ __forceinline__ void mask_residuals(int (&r)[S], u32 base, u32 order, u32 n)
{
    if (base < order || base + S > n) {
#pragma unroll
        for (int j = 0; j < S; j++) {
            const u32 i = base + j;
            if (i < order || i >= n) r[j] = 0;
        }
    }
}

// Rice parameter + estimate of one partition, same results as partition_estimate():
// closed form when no shift of the reference's loop can wrap 32 bits, the loop itself otherwise.
__device__ __forceinline__ u64 partition_estimate_fast(u32 plength, u64 S_, u32 max_rice, u32* k_out)
{
    u32 k;
    if (S_ <= (u64)plength) {
        k = 0;
    } else if (plength == 0 || ((u64)plength << max_rice) >= (1ull << 32)) {
        return partition_estimate(plength, S_, max_rice, k_out);
    } else {
        // smallest k with (plength << k) >= S_; none of the shifts up to max_rice wraps
        int kc = (64 - __clzll((long long)S_)) - (32 - __clz((int)plength)) - 1;
        if (kc < 0) kc = 0;
        while (((u64)plength << kc) < S_) kc++;
        k = (u32)kc < max_rice ? (u32)kc : max_rice;
    }
    u64 est;
    if (k > 0) est = 4ull + (S_ >> (k - 1)) + (u64)(u32)((1u + k) * plength) - (u64)(plength / 2);
    else       est = 4ull + (S_ << 1) + (u64)plength - (u64)(plength / 2);
    *k_out = k;
    return est;
}

// shared scratch of the fast Rice search (double buffered by search parity, see rice_search_regs)
struct RiceScratch {
    u64 lvl_total[2][16];
    u64 bits_total[2];
    u32 lvl_maxk[2][16];
};

// Rice search over the thread-resident residual r[] (already masked).  Same algorithm and
// same results as rice_search() in k_analyze.cuh.  `parity` alternates between consecutive
// calls so that a fast thread zeroing the next search's totals never races a slow reader.
template <int S>
__device__ __forceinline__ void rice_search_regs(const AnalyzeCtx& c, RiceScratch* rs, u32 parity, const int (&r)[S],
                                                 u32 base, u32 order, const bf_dev_params& P, RiceChoice* out)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const u32 n = (u32)c.n;
    const u32 po_eff = min(P.po_lim, (u32)(__ffs((int)n) - 1));
    u32 F = po_eff;
    while (F > 0 && (n >> F) < order) F--;
    const u32 nfine = 1u << F;
    const u32 plenF = n >> F;
    const u32 lo = max(base, order), hi = min(base + (u32)S, n);
    u64* fine = c.psum;              // [nfine] finest partition sums, then inclusive prefix sums in place
    u64* lvl_total = rs->lvl_total[parity];
    u32* lvl_maxk = rs->lvl_maxk[parity];

    // How the per-thread sums reach the finest partitions, decided once per CTA:
    //   grouped : a partition is g = plenF/S whole runs, g a power of two <= 32 -> warp shuffles, plain store
    //   owned   : a run is S/plenF whole partitions                            -> plain stores
    //   else    : shared-memory atomics into a zeroed array (odd tail blocks)
    const u32 g = plenF / S;
    const bool grouped = (plenF >= (u32)S) && (plenF % S == 0) && ((g & (g - 1)) == 0) && g <= 32;
    const bool owned = (plenF < (u32)S) && ((u32)S % plenF == 0);
    if (!grouped && !owned) for (u32 i = tid; i < nfine; i += nt) fine[i] = 0ull;
    if (tid < 16) { lvl_total[tid] = 0ull; lvl_maxk[tid] = 0u; }
    if (tid == 0) rs->bits_total[parity] = 0ull;
    if (!grouped && !owned) __syncthreads();

    // ---- pass 1: sum |r| per finest partition (r[] is zero outside [order, n)) ----
    if (grouped) {
        u64 run = 0;
#pragma unroll
        for (int j = 0; j < S; j++) run += (u64)(u32)abs(r[j]);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_xor_sync(0xFFFFFFFFu, run, o);
            if ((u32)o < g) run += t;
        }
        const u32 p = (u32)tid / g;
        if (((u32)tid & (g - 1)) == 0 && p < nfine) fine[p] = run;
    } else if (owned) {
        u32 p = base / plenF;
        u32 next = plenF;
        u64 run = 0;
#pragma unroll
        for (int j = 0; j < S; j++) {
            if ((u32)j == next) {
                if (p < nfine) fine[p] = run;
                run = 0; p++; next += plenF;
            }
            run += (u64)(u32)abs(r[j]);
        }
        if (p < nfine) fine[p] = run;
    } else if (lo < hi) {
        u32 p = lo / plenF;
        u32 next = (p + 1) * plenF - base;
        u64 run = 0;
#pragma unroll
        for (int j = 0; j < S; j++) {
            if ((u32)j == next) {
                if (p < nfine) atomicAdd(&fine[p], run);
                run = 0; p++; next += plenF;
            }
            run += (u64)(u32)abs(r[j]);
        }
        if (p < nfine) atomicAdd(&fine[p], run);
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

    // ---- one thread per (level, partition): Rice parameter, estimate.  Nodes are numbered as in a
    // binary heap (node 1 = level 0; level l = nodes 2^l .. 2^(l+1)-1), thread t takes node t, so a
    // level's nodes are an aligned lane group: its total and max are warp reductions, not atomics ----
    const u32 heapn = 2 * nfine - 1;
    for (u32 node0 = 0; node0 <= heapn; node0 += nt) {
        const u32 node = node0 + tid;
        const bool act = node >= 1 && node <= heapn;
        u32 l = 0, k = 0;
        u64 est = 0;
        if (act) {
            l = 31u - (u32)__clz((int)node);
            const u32 p = node - (1u << l);
            const u32 w = nfine >> l;                   // finest partitions per partition of level l
            const u64 hi_sum = fine[(p + 1) * w - 1];
            const u64 lo_sum = p ? fine[p * w - 1] : 0ull;
            const u32 plength = (n >> l) - (p == 0 ? order : 0u);
            est = partition_estimate_fast(plength, hi_sum - lo_sum, P.max_rice, &k);
            c.karr[node - 1] = (uint8_t)k;
        }
        if (node0 + (u32)(tid & ~31) >= 32) {
            // the whole warp sits inside one level (or past the heap)
            u32 km = __reduce_max_sync(0xFFFFFFFFu, k);
#pragma unroll
            for (int o = 16; o; o >>= 1) est += __shfl_xor_sync(0xFFFFFFFFu, est, o);
            const u32 l0 = __shfl_sync(0xFFFFFFFFu, l, 0);
            if (lane == 0 && node <= heapn) { atomicAdd(&lvl_total[l0], est); atomicMax(&lvl_maxk[l0], km); }
        } else {
            // nodes 1..31 = levels 0..4: segmented butterfly inside groups of 2^l lanes
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
#pragma unroll
    for (u32 l = 1; l <= BF_MAX_PO; l++) {
        if (l <= po_eff) {
            u64 tot;
            u32 kk = 0;
            if (l <= F) {
                tot = lvl_total[l];
            } else {
                // underflow level: partition 0 swallows every residual, the others are empty (H3)
                const u32 pl = n >> l;
                tot = partition_estimate(pl - order, fine[nfine - 1], P.max_rice, &kk);
                tot += (u64)((1u << l) - 1u) * (4ull + (u64)pl - (u64)(pl / 2));
            }
            if (tot < best) { best = tot; po = l; k0 = kk; }
        }
    }
    const u32 under = po > F ? 1u : 0u;
    const u32 koff = (1u << po) - 1u;
    const u32 plen = n >> po;
    const u32 maxk = under ? k0 : lvl_maxk[po];

    // ---- pass 2: exact size ----
    u64 bits = 0;
    if (lo < hi) {
        const u32 p0 = under ? 0u : lo / plen;
        if (under || (hi - 1) / plen == p0) {
            const u32 k = under ? k0 : (u32)c.karr[koff + p0];
            if (k >= 6) {
                u32 acc = 0; // S * (2^32 >> 6) < 2^32
#pragma unroll
                for (int j = 0; j < S; j++) acc += zigzag(r[j]) >> k;
                bits = acc;
            } else {
#pragma unroll
                for (int j = 0; j < S; j++) bits += (u64)(zigzag(r[j]) >> k);
            }
            bits += (u64)(hi - lo) * (1u + k);
        } else {
            u32 p = p0;
            u32 next = (p + 1) * plen - base;
            u32 k = c.karr[koff + p];
#pragma unroll
            for (int j = 0; j < S; j++) {
                if ((u32)j == next) { p++; next += plen; k = c.karr[koff + min(p, (1u << po) - 1u)]; }
                const u32 i = base + j;
                if (i >= lo && i < hi) bits += (u64)(zigzag(r[j]) >> k) + 1u + k;
            }
        }
    }
    // warp sum (32-bit REDUX when it cannot overflow), one shared atomic per warp
    if (!__any_sync(0xFFFFFFFFu, (bits >> 26) != 0)) {
        const u32 wsum = __reduce_add_sync(0xFFFFFFFFu, (u32)bits);
        if (lane == 0) atomicAdd(&rs->bits_total[parity], (u64)wsum);
    } else {
#pragma unroll
        for (int o = 16; o; o >>= 1) bits += __shfl_xor_sync(0xFFFFFFFFu, bits, o);
        if (lane == 0) atomicAdd(&rs->bits_total[parity], bits);
    }
    __syncthreads();
    out->po = po; out->under = under; out->k0 = k0;
    out->method = maxk > 14 ? 1u : 0u;
    out->bits = rs->bits_total[parity] + 6ull + (u64)(1u << po) * (maxk > 14 ? 5ull : 4ull);
}

// loads the thread's run of candidate samples into registers (0 beyond n)
template <int S>
__device__ __forceinline__ void load_run(const uint8_t* __restrict__ pcm, u64 pcm_off, u32 base, u32 n, u32 cand,
                                         const bf_dev_params& P, int (&s)[S])
{
    if (P.stereo && P.bytes_ps == 2 && base + S <= n && (((pcm_off + base) * 4) & 15) == 0 && (S % 4) == 0) {
        // stereo 16-bit: 16-byte loads of left/right pairs, candidate chosen once per thread
        const uint4* src = (const uint4*)(pcm + (pcm_off + base) * 4);
        u32 raw[S];
#pragma unroll
        for (int v = 0; v < S / 4; v++) {
            const uint4 w = __ldg(src + v);
            raw[v * 4 + 0] = w.x; raw[v * 4 + 1] = w.y; raw[v * 4 + 2] = w.z; raw[v * 4 + 3] = w.w;
        }
        if (cand == 0) {
#pragma unroll
            for (int j = 0; j < S; j++) s[j] = (int)(short)(raw[j] & 0xFFFF);
        } else if (cand == 1) {
#pragma unroll
            for (int j = 0; j < S; j++) s[j] = (int)raw[j] >> 16;
        } else if (cand == 2) {
#pragma unroll
            for (int j = 0; j < S; j++) s[j] = ((int)(short)(raw[j] & 0xFFFF) + ((int)raw[j] >> 16)) >> 1;
        } else {
#pragma unroll
            for (int j = 0; j < S; j++) s[j] = (int)(short)(raw[j] & 0xFFFF) - ((int)raw[j] >> 16);
        }
    } else {
#pragma unroll
        for (int j = 0; j < S; j++) {
            const u32 i = base + j;
            s[j] = i < n ? ld_candidate(pcm, pcm_off + i, cand, P) : 0;
        }
    }
}

// copies the thread's run to shared memory (one pad word per thread)
template <int S>
__device__ __forceinline__ void store_run(int* samp, const int (&s)[S], u32 tid, u32 base, u32 n)
{
    int* dst = samp + (size_t)tid * (S + 1);
    if (base + S <= n) {
#pragma unroll
        for (int j = 0; j < S; j++) dst[j] = s[j];
    } else {
#pragma unroll
        for (int j = 0; j < S; j++) if (base + j < n) dst[j] = s[j];
    }
}

template <int S>
__global__ void __launch_bounds__(512)
k_analyze_fast(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd, bf_dev_params P,
               const bf_lpc_head* __restrict__ heads, const short* __restrict__ coefs,
               b200flac_plan* __restrict__ plans, uint8_t* __restrict__ rice_out)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ u64 red[40];
    __shared__ RiceScratch rs;
    __shared__ u64 fsum[5];
    __shared__ u32 sc[8];
    __shared__ short s_q[BF_MAX_ORDER];
    __shared__ bf_lpc_head s_head;

    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const u32 unit = blockIdx.x;
    const u32 frame = unit / P.K, cand = unit % P.K;
    const bf_frame_desc d = fd[frame];
    const u32 n = d.nsamp;
    const u32 bps = candidate_bps(cand, P);
    const u32 base = (u32)tid * S;

    AnalyzeCtx c;
    c.n = (int)n; c.red = red; c.lvl_total = nullptr; c.sc = sc; c.resid = nullptr;
    unsigned char* sp = dyn_smem;
    c.samp = (int*)sp; sp += ((size_t)P.block_size + (P.block_size + S - 1) / S + 2) * 4;
    sp = (unsigned char*)(((uintptr_t)sp + 7) & ~(uintptr_t)7);
    c.psum = (u64*)sp; sp += (size_t)P.heap_entries * 8;
    c.karr = sp; sp += P.heap_entries;
    uint8_t* kfix = sp; sp += P.rice_stride;
    uint8_t* klpc = sp;

    b200flac_plan plan;
    plan.type = BF_VERBATIM; plan.order = 0; plan.wasted = 0; plan.precision = 0; plan.shift = 0;
    plan.coding_method = 0; plan.partition_order = 0; plan.flags = 0; plan.bits = 0;
#pragma unroll
    for (int i = 0; i < BF_MAX_ORDER; i++) plan.coeffs[i] = 0;
    uint8_t* my_rice = rice_out + (size_t)unit * P.rice_stride;

    if (tid == 0 && P.try_lpc) s_head = heads[unit];
    if (tid < 5) fsum[tid] = 0ull;

    // ---- load, constant check, wasted bits ----
    int s[S];
    load_run<S>(pcm, d.pcm_off, base, n, cand, P, s);
    const int first = ld_candidate(pcm, d.pcm_off, cand, P);
    u32 orv = 0, diff = 0;
    if (base + S <= n) {
#pragma unroll
        for (int j = 0; j < S; j++) { orv |= (u32)s[j]; diff |= (u32)(s[j] ^ first); }
    } else {
#pragma unroll
        for (int j = 0; j < S; j++) { orv |= (u32)s[j]; if (base + j < n) diff |= (u32)(s[j] ^ first); }
    }
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
#pragma unroll
    for (int j = 0; j < S; j++) s[j] >>= wasted;
    store_run<S>(c.samp, s, (u32)tid, base, n);
    __syncthreads();
    const u32 sub_bps = bps - wasted;
    const u32 hdr_bits = 8 + wasted;
    int r[S];
    u32 search = 0;

    // ---- FIXED ----
    u64 fixed_bits = 0;
    u32 fixed_order = 0;
    RiceChoice rfix; rfix.po = 0; rfix.under = 0; rfix.k0 = 0; rfix.method = 0; rfix.bits = 0;
    if (P.try_fixed) {
        if (n > 4) {
            const u32 lo = max(base, 4u), hi = min(base + (u32)S, n);
            u32 f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0;       // 32-bit partial sums (fast threads)
            u64 e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;       // 64-bit partial sums (edge threads / wide samples)
            bool slow = false;
            if (lo < hi) {
                u32 a1 = 0, a2 = 0, a3 = 0, a4 = 0;
                if (base >= 1) a1 = (u32)c.samp[pad_idx<S>(base - 1)]; // lo < hi <= n here, so base < n
                if (base >= 2) a2 = (u32)c.samp[pad_idx<S>(base - 2)];
                if (base >= 3) a3 = (u32)c.samp[pad_idx<S>(base - 3)];
                if (base >= 4) a4 = (u32)c.samp[pad_idx<S>(base - 4)];
                u32 p1 = a1 - a2, p2 = p1 - (a2 - a3), p3 = p2 - ((a2 - a3) - (a3 - a4));
                u32 prev = a1;
                if ((base + S <= n) && sub_bps <= 18) {
                    // |d4| < 2^(sub_bps+3) <= 2^21: S <= 36 terms and then 32 lanes still fit 32 bits
#pragma unroll
                    for (int j = 0; j < S; j++) {
                        const u32 x = (u32)s[j];
                        const u32 d1 = x - prev, d2 = d1 - p1, d3 = d2 - p2, d4 = d3 - p3;
                        f0 += (u32)abs((int)x); f1 += (u32)abs((int)d1); f2 += (u32)abs((int)d2);
                        f3 += (u32)abs((int)d3); f4 += (u32)abs((int)d4);
                        prev = x; p1 = d1; p2 = d2; p3 = d3;
                    }
                    if (base < 4) {
                        // the block's first four samples are warm-up for every order (flac.c:877-889):
                        // take their terms (computed above with a zero history) out again
                        u32 q0 = 0, q1 = 0, q2 = 0, q3 = 0;
#pragma unroll
                        for (int j = 0; j < 4 && j < S; j++) {
                            if (base + j < 4) {
                                const u32 x = (u32)s[j];
                                const u32 d1 = x - q0, d2 = d1 - q1, d3 = d2 - q2, d4 = d3 - q3;
                                f0 -= (u32)abs((int)x); f1 -= (u32)abs((int)d1); f2 -= (u32)abs((int)d2);
                                f3 -= (u32)abs((int)d3); f4 -= (u32)abs((int)d4);
                                q0 = x; q1 = d1; q2 = d2; q3 = d3;
                            }
                        }
                    }
                } else {
                    slow = true;
#pragma unroll
                    for (int j = 0; j < S; j++) {
                        const u32 x = (u32)s[j];
                        const u32 d1 = x - prev, d2 = d1 - p1, d3 = d2 - p2, d4 = d3 - p3;
                        const u32 i = base + j;
                        if (i >= 4 && i < n) {
                            e0 += (u64)(long long)abs((int)x); e1 += (u64)(long long)abs((int)d1);
                            e2 += (u64)(long long)abs((int)d2); e3 += (u64)(long long)abs((int)d3);
                            e4 += (u64)(long long)abs((int)d4);
                        }
                        prev = x; p1 = d1; p2 = d2; p3 = d3;
                    }
                }
            }
            f0 = __reduce_add_sync(0xFFFFFFFFu, f0); f1 = __reduce_add_sync(0xFFFFFFFFu, f1);
            f2 = __reduce_add_sync(0xFFFFFFFFu, f2); f3 = __reduce_add_sync(0xFFFFFFFFu, f3);
            f4 = __reduce_add_sync(0xFFFFFFFFu, f4);
            if (lane == 0) {
                atomicAdd(&fsum[0], (u64)f0); atomicAdd(&fsum[1], (u64)f1); atomicAdd(&fsum[2], (u64)f2);
                atomicAdd(&fsum[3], (u64)f3); atomicAdd(&fsum[4], (u64)f4);
            }
            if (slow) {
                atomicAdd(&fsum[0], e0); atomicAdd(&fsum[1], e1); atomicAdd(&fsum[2], e2);
                atomicAdd(&fsum[3], e3); atomicAdd(&fsum[4], e4);
            }
            __syncthreads();
            // flac.c:877-893: first strict minimum over orders 0..4
            u64 best = fsum[0];
            if (fsum[1] < best) { best = fsum[1]; fixed_order = 1; }
            if (fsum[2] < best) { best = fsum[2]; fixed_order = 2; }
            if (fsum[3] < best) { best = fsum[3]; fixed_order = 3; }
            if (fsum[4] < best) { best = fsum[4]; fixed_order = 4; }
        }
        fixed_residual_regs<S>(s, c.samp, base, n, fixed_order, r);
        mask_residuals<S>(r, base, fixed_order, n);
        rice_search_regs<S>(c, &rs, (search++) & 1u, r, base, fixed_order, P, &rfix);
        save_rice(c, rfix, kfix);
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
                // sum|q| * 2^(sub_bps-1) < 2^31  =>  the 64-bit accumulator of flac.c:1000-1005 stays inside int32
                const u32 sumq = __reduce_add_sync(0xFFFFFFFFu, (u32)abs(q));
                if (tid == 0) sc[4] = (((u64)sumq << (sub_bps - 1)) < (1ull << 31)) ? 1u : 0u;
            }
            __syncthreads();
            const int shift = s_head.shift[o - 1];
            if (sc[4]) lpc_residual_dispatch<S, false>(s, c.samp, base, n, o, s_q, shift, r);
            else lpc_residual_dispatch<S, true>(s, c.samp, base, n, o, s_q, shift, r);
            mask_residuals<S>(r, base, o, n);
            RiceChoice rc;
            rice_search_regs<S>(c, &rs, (search++) & 1u, r, base, o, P, &rc);
            const u64 bits = hdr_bits + (u64)o * sub_bps + 4 + 5 + (u64)o * precision + rc.bits;
            if (!have || (u32)bits < best_bits) {
                have = true;
                best_bits = (u32)bits;
                lpc_bits = bits; lpc_order = o; lpc_shift = shift; lpc_precision = precision;
                rlpc = rc;
                save_rice(c, rc, klpc);
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
            // bit 1: the packer may accumulate this predictor in 32 bits (same test as above)
            if (((u64)sumq << (sub_bps - 1)) < (1ull << 31)) plan.flags |= 2;
        }
        for (u32 p = tid; p < (1u << rlpc.po); p += nt) my_rice[p] = klpc[p];
    } else {
        plan.type = BF_VERBATIM;
        plan.bits = hdr_bits + sub_bps * n;
    }
    if (tid == 0) plans[unit] = plan;
}
