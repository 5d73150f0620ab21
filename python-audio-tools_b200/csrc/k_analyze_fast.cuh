// k_analyze_fast.cuh -- register-resident version of k_analyze for the common
// shapes: the whole block fits one pass (n <= blockDim * S) and shared memory.
//
// Same decisions, bit for bit, as k_analyze (k_analyze.cuh) and therefore as the
// reference (flac.c:673-1505); only the data movement differs:
//   * every thread keeps its S contiguous samples (after the wasted-bits shift)
//     in registers for the whole kernel; shared memory holds a copy only so that
//     neighbours can fetch their predictor history;
//   * FIXED differences, the LPC multiply-accumulate and both residual passes of
//     the Rice search are fully unrolled over those registers: no index math,
//     no shared-memory traffic in the inner loops;
//   * the LPC accumulate runs in 32 bits whenever sum|q| * 2^(bps-1) < 2^31
//     proves the 64-bit sum of flac.c:1000-1005 cannot leave int32 (always true
//     for 16-bit input at order <= 15), in 64 bits otherwise;
//   * the partition-order decision (merge, Rice parameters, estimates, argmin)
//     is done by warp 0 alone between two block barriers.
#pragma once
#include "flac_common.cuh"
#include "k_analyze.cuh"

// order groups: coefficients are zero-padded to the group size (exact: 0 * x adds nothing)
__device__ __forceinline__ int order_group(u32 o) { return o <= 4 ? 4 : o <= 8 ? 8 : o <= 12 ? 12 : o <= 16 ? 16 : o <= 24 ? 24 : 32; }

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
        h[t] = (idx >= 0 && idx < (int)n) ? samp[PADI(idx)] : 0; // idle threads (base >= n) must not read past the block
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
    switch (order_group(order)) {
    case 4: lpc_residual_regs<S, 4, WIDE>(s, samp, base, n, q_sm, shift, r); break;
    case 8: lpc_residual_regs<S, 8, WIDE>(s, samp, base, n, q_sm, shift, r); break;
    case 12: lpc_residual_regs<S, 12, WIDE>(s, samp, base, n, q_sm, shift, r); break;
    case 16: lpc_residual_regs<S, 16, WIDE>(s, samp, base, n, q_sm, shift, r); break;
    case 24: lpc_residual_regs<S, 24, WIDE>(s, samp, base, n, q_sm, shift, r); break;
    default: lpc_residual_regs<S, 32, WIDE>(s, samp, base, n, q_sm, shift, r); break;
    }
}

// FIXED residual of `order` for the thread's run (closed forms of the iterated differences)
template <int S>
__device__ __forceinline__ void fixed_residual_regs(const int (&s)[S], const int* samp, u32 base, u32 n, u32 order, int (&r)[S])
{
    u32 h1 = 0, h2 = 0, h3 = 0, h4 = 0;
    if (base >= 1 && base - 1 < n) h1 = (u32)samp[PADI(base - 1)];
    if (base >= 2 && base - 2 < n) h2 = (u32)samp[PADI(base - 2)];
    if (base >= 3 && base - 3 < n) h3 = (u32)samp[PADI(base - 3)];
    if (base >= 4 && base - 4 < n) h4 = (u32)samp[PADI(base - 4)];
#pragma unroll
    for (int j = 0; j < S; j++) {
        const u32 s0 = (u32)s[j];
        const u32 s1 = j >= 1 ? (u32)s[j >= 1 ? j - 1 : 0] : h1;
        const u32 s2 = j >= 2 ? (u32)s[j >= 2 ? j - 2 : 0] : (j == 1 ? h1 : h2);
        const u32 s3 = j >= 3 ? (u32)s[j >= 3 ? j - 3 : 0] : (j == 2 ? h1 : j == 1 ? h2 : h3);
        const u32 s4 = j >= 4 ? (u32)s[j >= 4 ? j - 4 : 0] : (j == 3 ? h1 : j == 2 ? h2 : j == 1 ? h3 : h4);
        u32 v;
        switch (order) {
        case 0: v = s0; break;
        case 1: v = s0 - s1; break;
        case 2: v = s0 - 2u * s1 + s2; break;
        case 3: v = s0 - 3u * s1 + 3u * s2 - s3; break;
        default: v = s0 - 4u * s1 + 6u * s2 - 4u * s3 + s4; break;
        }
        r[j] = (int)v;
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

// Rice search over the thread-resident residual r[] (already masked).  Same algorithm and
// same results as rice_search() in k_analyze.cuh.
template <int S>
__device__ __forceinline__ void rice_search_regs(const AnalyzeCtx& c, const int (&r)[S], u32 base, u32 order,
                                                 const bf_dev_params& P, RiceChoice* out)
{
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
    const u32 n = (u32)c.n;
    const u32 po_eff = min(P.po_lim, (u32)(__ffs((int)n) - 1));
    u32 F = po_eff;
    while (F > 0 && (n >> F) < order) F--;
    const u32 plenF = n >> F;
    const u32 lo = max(base, order), hi = min(base + (u32)S, n);

    for (u32 i = tid; i < (1u << F); i += nt) c.psum[(1u << F) - 1u + i] = 0ull;
    __syncthreads();

    // ---- pass 1: sum |r| per finest partition ----
    if (lo < hi) {
        const u32 p0 = lo / plenF;
        if ((hi - 1) / plenF == p0) {
            // whole run inside one partition (the common case: partition length multiple of S)
            u64 run = 0;
#pragma unroll
            for (int j = 0; j < S; j++) run += (u64)(u32)abs(r[j]);
            atomicAdd(&c.psum[(1u << F) - 1u + p0], run);
        } else {
            u32 p = p0;
            u32 next = (p + 1) * plenF - base; // relative index of the next boundary
            u64 run = 0;
#pragma unroll
            for (int j = 0; j < S; j++) {
                if ((u32)j == next) {
                    if (p < (1u << F)) atomicAdd(&c.psum[(1u << F) - 1u + p], run);
                    run = 0; p++; next += plenF;
                }
                run += (u64)(u32)abs(r[j]);
            }
            if (p < (1u << F)) atomicAdd(&c.psum[(1u << F) - 1u + p], run);
        }
    }
    __syncthreads();

    // ---- decision by warp 0: merge levels, Rice parameters, estimates, first strict minimum ----
    if (tid < 32) {
        for (int l = (int)F - 1; l >= 0; l--) {
            for (u32 p = lane; p < (1u << l); p += 32)
                c.psum[(1u << l) - 1u + p] = c.psum[(2u << l) - 1u + 2 * p] + c.psum[(2u << l) - 1u + 2 * p + 1];
            __syncwarp();
        }
        u64 best = ~0ull;
        u32 best_l = 0, best_k0 = 0;
        for (u32 l = 0; l <= po_eff; l++) {
            u64 tot = 0;
            u32 k0 = 0;
            if (l <= F) {
                for (u32 p = lane; p < (1u << l); p += 32) {
                    const u32 plength = (n >> l) - (p == 0 ? order : 0u);
                    u32 k;
                    tot += partition_estimate(plength, c.psum[(1u << l) - 1u + p], P.max_rice, &k);
                    c.karr[(1u << l) - 1u + p] = (uint8_t)k;
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
            } else {
                const u32 pl = n >> l;
                tot = partition_estimate(pl - order, c.psum[0], P.max_rice, &k0);
                tot += (u64)((1u << l) - 1u) * (4ull + (u64)pl - (u64)(pl / 2));
            }
            if (tot < best) { best = tot; best_l = l; best_k0 = k0; }
        }
        if (lane == 0) { c.sc[0] = best_l; c.sc[1] = (best_l > F) ? 1u : 0u; c.sc[2] = best_k0; }
    }
    __syncthreads();
    const u32 po = c.sc[0], under = c.sc[1], k0 = c.sc[2];
    const u32 koff = (1u << po) - 1u;
    const u32 plen = n >> po;

    // ---- pass 2: exact size ----
    u32 maxk = 0;
    if (under) maxk = k0;
    else for (u32 p = tid; p < (1u << po); p += nt) maxk = max(maxk, (u32)c.karr[koff + p]);
    u64 bits = 0;
    if (lo < hi) {
        const u32 p0 = under ? 0u : lo / plen;
        if (under || (hi - 1) / plen == p0) {
            const u32 k = under ? k0 : (u32)c.karr[koff + p0];
            u32 acc = 0; // S * (2^32 >> k) can overflow 32 bits only for k == 0 with huge residuals: use 64
            u64 acc64 = 0;
            if (k >= 6) {
#pragma unroll
                for (int j = 0; j < S; j++) acc += zigzag(r[j]) >> k;
                acc64 = acc;
            } else {
#pragma unroll
                for (int j = 0; j < S; j++) acc64 += (u64)(zigzag(r[j]) >> k);
            }
            bits = acc64 + (u64)(hi - lo) * (1u + k);
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
    maxk = block_max_u32(maxk, c.red);
    bits = block_sum_u64(bits, c.red);
    out->po = po; out->under = under; out->k0 = k0;
    out->method = maxk > 14 ? 1u : 0u;
    out->bits = bits + 6ull + (u64)(1u << po) * (maxk > 14 ? 5ull : 4ull);
}

// loads the thread's run of candidate samples into registers (0 beyond n)
template <int S>
__device__ __forceinline__ void load_run(const uint8_t* __restrict__ pcm, u64 pcm_off, u32 base, u32 n, u32 cand,
                                         const bf_dev_params& P, int (&s)[S])
{
    if (P.stereo && P.bytes_ps == 2 && base + S <= n && (((pcm_off + base) * 4) & 15) == 0 && (S % 4) == 0) {
        const uint4* src = (const uint4*)(pcm + (pcm_off + base) * 4);
#pragma unroll
        for (int v = 0; v < S / 4; v++) {
            const uint4 w = __ldg(src + v);
            const u32 ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int L = (int)(short)(ws[e] & 0xFFFF), R = (int)(short)(ws[e] >> 16);
                s[v * 4 + e] = cand == 0 ? L : cand == 1 ? R : cand == 2 ? ((L + R) >> 1) : (L - R);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < S; j++) {
            const u32 i = base + j;
            s[j] = i < n ? ld_candidate(pcm, pcm_off + i, cand, P) : 0;
        }
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
    __shared__ u64 lvl_total[16];
    __shared__ u32 sc[8];
    __shared__ short s_q[BF_MAX_ORDER];

    const int tid = threadIdx.x, nt = blockDim.x;
    const u32 unit = blockIdx.x;
    const u32 frame = unit / P.K, cand = unit % P.K;
    const bf_frame_desc d = fd[frame];
    const u32 n = d.nsamp;
    const u32 bps = candidate_bps(cand, P);
    const u32 base = (u32)tid * S;

    AnalyzeCtx c;
    c.n = (int)n; c.red = red; c.lvl_total = lvl_total; c.sc = sc; c.resid = nullptr;
    unsigned char* sp = dyn_smem;
    c.samp = (int*)sp; sp += ((size_t)PADI(P.block_size) + 1) * 4;
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

    // ---- load, constant check, wasted bits ----
    int s[S];
    load_run<S>(pcm, d.pcm_off, base, n, cand, P, s);
    const int first = ld_candidate(pcm, d.pcm_off, cand, P);
    u32 orv = 0, diff = 0;
#pragma unroll
    for (int j = 0; j < S; j++) {
        orv |= (u32)s[j];
        diff |= (base + j < n) ? (u32)(s[j] ^ first) : 0u;
    }
    orv = block_or_u32(orv, red);
    diff = block_or_u32(diff, red);
    if (P.try_constant && diff == 0) {
        if (tid == 0) { plan.type = BF_CONSTANT; plan.bits = 8 + bps; plans[unit] = plan; }
        return;
    }
    const u32 wasted = orv ? (u32)(__ffs((int)orv) - 1) : 0u;
#pragma unroll
    for (int j = 0; j < S; j++) {
        s[j] >>= wasted;
        if (base + j < n) c.samp[PADI(base + j)] = s[j];
    }
    __syncthreads();
    const u32 sub_bps = bps - wasted;
    const u32 hdr_bits = 8 + wasted;
    int r[S];

    // ---- FIXED ----
    u64 fixed_bits = 0;
    u32 fixed_order = 0;
    RiceChoice rfix; rfix.po = 0; rfix.under = 0; rfix.k0 = 0; rfix.method = 0; rfix.bits = 0;
    if (P.try_fixed) {
        if (n > 4) {
            u64 e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
            const u32 lo = max(base, 4u), hi = min(base + (u32)S, n);
            if (lo < hi) {
                // differences of the samples before the run
                u32 a1 = 0, a2 = 0, a3 = 0, a4 = 0;
                if (base >= 1) a1 = (u32)c.samp[PADI(base - 1)]; // lo < hi <= n here, so base < n
                if (base >= 2) a2 = (u32)c.samp[PADI(base - 2)];
                if (base >= 3) a3 = (u32)c.samp[PADI(base - 3)];
                if (base >= 4) a4 = (u32)c.samp[PADI(base - 4)];
                u32 p1 = a1 - a2, p2 = p1 - (a2 - a3), p3 = p2 - ((a2 - a3) - (a3 - a4));
                u32 prev = a1;
                const bool full = (base >= 4) && (base + S <= n);
                if (full && sub_bps <= 18) {
                    // |d4| < 2^(bps+4): 32 terms fit 32 bits
                    u32 f0 = 0, f1 = 0, f2 = 0, f3 = 0, f4 = 0;
#pragma unroll
                    for (int j = 0; j < S; j++) {
                        const u32 x = (u32)s[j];
                        const u32 d1 = x - prev, d2 = d1 - p1, d3 = d2 - p2, d4 = d3 - p3;
                        f0 += (u32)abs((int)x); f1 += (u32)abs((int)d1); f2 += (u32)abs((int)d2);
                        f3 += (u32)abs((int)d3); f4 += (u32)abs((int)d4);
                        prev = x; p1 = d1; p2 = d2; p3 = d3;
                    }
                    e0 = f0; e1 = f1; e2 = f2; e3 = f3; e4 = f4;
                } else {
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
            e0 = block_sum_u64(e0, red); e1 = block_sum_u64(e1, red); e2 = block_sum_u64(e2, red);
            e3 = block_sum_u64(e3, red); e4 = block_sum_u64(e4, red);
            u64 best = e0;
            if (e1 < best) { best = e1; fixed_order = 1; }
            if (e2 < best) { best = e2; fixed_order = 2; }
            if (e3 < best) { best = e3; fixed_order = 3; }
            if (e4 < best) { best = e4; fixed_order = 4; }
        }
        fixed_residual_regs<S>(s, c.samp, base, n, fixed_order, r);
        mask_residuals<S>(r, base, fixed_order, n);
        rice_search_regs<S>(c, r, base, fixed_order, P, &rfix);
        save_rice(c, rfix, kfix);
        fixed_bits = hdr_bits + (u64)fixed_order * sub_bps + rfix.bits;
    }

    // ---- LPC ----
    u64 lpc_bits = 0;
    u32 lpc_order = 0, lpc_precision = 0;
    int lpc_shift = 0;
    RiceChoice rlpc; rlpc.po = 0; rlpc.under = 0; rlpc.k0 = 0; rlpc.method = 0; rlpc.bits = 0;
    if (P.try_lpc) {
        const bf_lpc_head head = heads[unit];
        const short* mycoef = coefs + (size_t)unit * P.model_stride;
        const u32 o_first = (P.exhaustive && !head.dummy) ? 1u : head.best_order;
        const u32 o_last = (P.exhaustive && !head.dummy) ? P.max_lpc_order : head.best_order;
        u64 best_bits = 0xFFFFFFFFull;
        bool have = false;
        for (u32 o = o_first; o <= o_last; o++) {
            __syncthreads();
            if (tid < BF_MAX_ORDER) s_q[tid] = tid < (int)o ? mycoef[(o * (o - 1)) / 2 + tid] : (short)0;
            __syncthreads();
            const int shift = head.shift[o - 1];
            // sum|q| * 2^(sub_bps-1) < 2^31  =>  the 64-bit accumulator of flac.c:1000-1005 stays inside int32
            u32 sumq = 0;
            for (u32 j = 0; j < o; j++) sumq += (u32)abs((int)s_q[j]);
            const bool narrow = ((u64)sumq << (sub_bps - 1)) < (1ull << 31);
            if (narrow) lpc_residual_dispatch<S, false>(s, c.samp, base, n, o, s_q, shift, r);
            else lpc_residual_dispatch<S, true>(s, c.samp, base, n, o, s_q, shift, r);
            mask_residuals<S>(r, base, o, n);
            RiceChoice rc;
            rice_search_regs<S>(c, r, base, o, P, &rc);
            const u64 bits = hdr_bits + (u64)o * sub_bps + 4 + 5 + (u64)o * head.precision + rc.bits;
            if (!have || (u32)bits < best_bits) {
                have = true;
                best_bits = (u32)bits;
                lpc_bits = bits; lpc_order = o; lpc_shift = shift; lpc_precision = head.precision;
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
            const short* mycoef = coefs + (size_t)unit * P.model_stride + (lpc_order * (lpc_order - 1)) / 2;
            for (u32 j = 0; j < lpc_order; j++) plan.coeffs[j] = mycoef[j];
        }
        for (u32 p = tid; p < (1u << rlpc.po); p += nt) my_rice[p] = klpc[p];
    } else {
        plan.type = BF_VERBATIM;
        plan.bits = hdr_bits + sub_bps * n;
    }
    if (tid == 0) plans[unit] = plan;
}
