// k_analyze.cuh -- integer model search, one CTA per unit (frame, candidate).
//
//   constant / wasted bits   flacenc_all_identical, flacenc_max_wasted_bits_per_sample
//                            flac.c:1606-1620, 1578-1604
//   FIXED orders 0..4        flacenc_write_fixed_subframe      flac.c:856-930
//   LPC residual             flacenc_encode_lpc_subframe       flac.c:961-1016
//   Rice partition search    flacenc_encode_residuals /
//                            flacenc_encode_residual_partitions flac.c:1326-1505
//   model choice             flacenc_write_subframe            flac.c:673-811
//   exhaustive order search  flacenc_best_lpc_coefficients     flac.c:1070-1120
//
// Output: one b200flac_plan per unit (+ its Rice parameters) carrying the EXACT
// bit size the reference's recorder would report (bits_written), so that the
// frame-level stereo choice (flac.c:581-652) and the packing offsets can be
// computed without writing a single candidate bit.
//
// Layout: the block's (wasted-shifted) samples and the current residual live in
// shared memory, skewed one word per 32 (PADI) so that each thread can own a
// contiguous run of S samples without bank conflicts.
#pragma once
#include "flac_common.cuh"

struct AnalyzeCtx {
    int n;             // samples in the block
    int* samp;         // [PADI(n)] wasted-shifted samples
    int* resid;        // [PADI(n)] residual of the model under test (index = sample index)
    u64* psum;         // partition-sum heap, level l at [(1<<l)-1, (2<<l)-1)
    uint8_t* karr;     // Rice parameter heap, same indexing
    u64* red;          // >= 40 u64 shared scratch
    u64* lvl_total;    // 16 u64 shared
    u32* sc;           // 8 u32 shared scalars
};

struct RiceChoice {
    u32 po;          // chosen partition order
    u32 under;       // 1: chosen level is a partition-length-underflow level
    u32 k0;          // Rice parameter of partition 0 when under
    u32 method;      // coding method (1 iff some k > 14)
    u64 bits;        // exact size of the residual block in bits
};

// Exact restatement of the per-partition part of flac.c:1461-1501
__device__ __forceinline__ u64 partition_estimate(u32 plength, u64 S, u32 max_rice, u32* k_out)
{
    u32 k = 0;
    while ((u64)(u32)(plength << k) < S) {   // 32-bit shift, then widened (H3)
        if (k < max_rice) k++; else break;
    }
    u64 est;
    if (k > 0) est = 4ull + (S >> (k - 1)) + (u64)(u32)((1u + k) * plength) - (u64)(plength / 2);
    else       est = 4ull + (S << 1) + (u64)plength - (u64)(plength / 2);
    *k_out = k;
    return est;
}

// Searches partition orders for the residual in ctx.resid (valid for sample
// indices order..n-1) exactly as flacenc_encode_residuals does, then counts the
// exact bits of the winner.  All threads of the CTA must call it.
template <int S>
__device__ void rice_search(const AnalyzeCtx& c, u32 order, const bf_dev_params& P, RiceChoice* out)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const u32 n = (u32)c.n;
    // partition orders tried: 0..po_eff (loop breaks at the first po with n % 2^po != 0, flac.c:1365,1389)
    const u32 po_eff = min(P.po_lim, (u32)(__ffs((int)n) - 1));
    // finest level whose partitions are at least `order` long; above it plength underflows (H3)
    u32 F = po_eff;
    while (F > 0 && (n >> F) < order) F--;
    const u32 heapn = (2u << F) - 1u;
    const u32 plenF = n >> F;

    for (u32 i = tid; i < heapn; i += nt) c.psum[i] = 0ull;
    if (tid < 16) c.lvl_total[tid] = 0ull;
    __syncthreads();

    // ---- sum |r| per finest partition ----
    for (u32 base = tid * S; base < n; base += nt * S) {
        const u32 lo = max(base, order), hi = min(base + S, n);
        if (lo < hi) {
            u32 p = lo / plenF;
            u32 next = (p + 1) * plenF;
            u64 run = 0;
            for (u32 i = lo; i < hi; i++) {
                if (i == next) {
                    atomicAdd(&c.psum[(1u << F) - 1u + p], run);
                    run = 0; p++; next += plenF;
                }
                const int r = c.resid[PADI(i)];
                run += (u64)(u32)(r < 0 ? (0u - (u32)r) : (u32)r);
            }
            atomicAdd(&c.psum[(1u << F) - 1u + p], run);
        }
    }
    __syncthreads();
    // ---- merge upward: a partition of order l is two partitions of order l+1 ----
    for (int l = (int)F - 1; l >= 0; l--) {
        for (u32 p = tid; p < (1u << l); p += nt)
            c.psum[(1u << l) - 1u + p] = c.psum[(2u << l) - 1u + 2 * p] + c.psum[(2u << l) - 1u + 2 * p + 1];
        __syncthreads();
    }
    // ---- Rice parameter + size estimate of every partition of every level ----
    for (u32 idx = tid; idx < heapn; idx += nt) {
        const u32 l = 31u - (u32)__clz((int)(idx + 1));
        const u32 p = idx + 1 - (1u << l);
        const u32 plength = (n >> l) - (p == 0 ? order : 0u);
        u32 k;
        const u64 est = partition_estimate(plength, c.psum[idx], P.max_rice, &k);
        c.karr[idx] = (uint8_t)k;
        atomicAdd(&c.lvl_total[l], est);
    }
    __syncthreads();
    if (tid == 0) {
        // underflow levels: partition 0 swallows every residual (the clamped split of
        // src/array.c:559-580), the others are empty but still cost 4 + plength - plength/2
        u32 k0_best = 0;
        u64 best = c.lvl_total[0];
        u32 best_l = 0;
        for (u32 l = 1; l <= po_eff; l++) {
            u64 tot;
            u32 k0 = 0;
            if (l <= F) {
                tot = c.lvl_total[l];
            } else {
                const u32 pl = n >> l;
                const u32 pl0 = pl - order; // wraps
                tot = partition_estimate(pl0, c.psum[0], P.max_rice, &k0);
                tot += (u64)((1u << l) - 1u) * (4ull + (u64)pl - (u64)(pl / 2));
            }
            if (tot < best) { best = tot; best_l = l; k0_best = k0; } // strict: first minimum wins
        }
        c.sc[0] = best_l;
        c.sc[1] = (best_l > F) ? 1u : 0u;
        c.sc[2] = k0_best;
    }
    __syncthreads();
    const u32 po = c.sc[0], under = c.sc[1], k0 = c.sc[2];
    const u32 koff = (1u << po) - 1u;
    const u32 plen = n >> po;

    // ---- exact size of the winner ----
    u32 maxk = 0;
    if (under) maxk = k0;
    else for (u32 p = tid; p < (1u << po); p += nt) maxk = max(maxk, (u32)c.karr[koff + p]);
    u64 bits = 0;
    for (u32 base = tid * S; base < n; base += nt * S) {
        const u32 lo = max(base, order), hi = min(base + S, n);
        if (lo < hi) {
            u32 p = under ? 0u : lo / plen;
            u32 next = under ? 0xFFFFFFFFu : (p + 1) * plen;
            u32 k = under ? k0 : (u32)c.karr[koff + p];
            for (u32 i = lo; i < hi; i++) {
                if (i == next) { p++; next += plen; k = c.karr[koff + p]; }
                const u32 u = zigzag(c.resid[PADI(i)]);
                bits += (u64)(u >> k) + 1u + k;
            }
        }
    }
    maxk = block_max_u32(maxk, c.red);
    bits = block_sum_u64(bits, c.red);
    out->po = po; out->under = under; out->k0 = k0;
    out->method = maxk > 14 ? 1u : 0u;
    out->bits = bits + 6ull + (u64)(1u << po) * (maxk > 14 ? 5ull : 4ull);
}

// copy the chosen level's Rice parameters out of the heap
__device__ __forceinline__ void save_rice(const AnalyzeCtx& c, const RiceChoice& rc, uint8_t* dst)
{
    const u32 np = 1u << rc.po, koff = np - 1u;
    for (u32 p = threadIdx.x; p < np; p += blockDim.x)
        dst[p] = rc.under ? (uint8_t)(p == 0 ? rc.k0 : 0u) : c.karr[koff + p];
}

// residual of FIXED order `order` at sample i (iterated differences of flac.c:918-930, closed form;
// identical modulo 2^32)
__device__ __forceinline__ int fixed_residual(const int* samp, u32 i, u32 order)
{
    const u32 s0 = (u32)samp[PADI(i)];
    if (order == 0) return (int)s0;
    const u32 s1 = (u32)samp[PADI(i - 1)];
    if (order == 1) return (int)(s0 - s1);
    const u32 s2 = (u32)samp[PADI(i - 2)];
    if (order == 2) return (int)(s0 - 2u * s1 + s2);
    const u32 s3 = (u32)samp[PADI(i - 3)];
    if (order == 3) return (int)(s0 - 3u * s1 + 3u * s2 - s3);
    const u32 s4 = (u32)samp[PADI(i - 4)];
    return (int)(s0 - 4u * s1 + 6u * s2 - 4u * s3 + s4);
}

// LPC residual at sample i, flac.c:999-1008: int64 accumulate, arithmetic >>, truncating cast (H1)
__device__ __forceinline__ int lpc_residual(const int* samp, u32 i, u32 order, const short* q, int shift)
{
    long long acc = 0;
    for (u32 j = 0; j < order; j++) acc = mad_wide((int)q[j], samp[PADI(i - 1 - j)], acc);
    acc >>= shift;
    return (int)((u32)samp[PADI(i)] - (u32)(int)acc);
}

// loads the unit's samples into c.samp, returns OR of all samples and whether all are identical
template <int S>
__device__ __forceinline__ void load_unit_samples(const AnalyzeCtx& c, const uint8_t* __restrict__ pcm,
                                                  u64 pcm_off, u32 cand, const bf_dev_params& P,
                                                  u32* or_out, u32* differ_out)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const u32 n = (u32)c.n;
    const int first = ld_candidate(pcm, pcm_off, cand, P);
    u32 orv = 0, diff = 0;
    for (u32 base = tid * S; base < n; base += nt * S) {
        const u32 hi = min(base + S, n);
        for (u32 i = base; i < hi; i++) {
            const int v = ld_candidate(pcm, pcm_off + i, cand, P);
            c.samp[PADI(i)] = v;
            orv |= (u32)v;
            diff |= (u32)(v ^ first);
        }
    }
    *or_out = block_or_u32(orv, c.red);
    *differ_out = block_or_u32(diff, c.red);
}

template <int S>
__global__ void k_analyze(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd,
                          bf_dev_params P, const bf_lpc_head* __restrict__ heads,
                          const short* __restrict__ coefs, b200flac_plan* __restrict__ plans,
                          uint8_t* __restrict__ rice_out, int* __restrict__ g_samples,
                          u64* __restrict__ g_heap, uint8_t* __restrict__ g_karr)
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

    // ---- carve buffers ----
    AnalyzeCtx c;
    c.n = (int)n; c.red = red; c.lvl_total = lvl_total; c.sc = sc;
    unsigned char* sp = dyn_smem;
    uint8_t *kfix, *klpc;
    if (P.samples_in_smem) {
        const u32 padn = PADI(P.block_size) + 1;
        c.samp = (int*)sp; sp += (size_t)padn * 4;
        c.resid = (int*)sp; sp += (size_t)padn * 4;
    } else {
        c.samp = g_samples + (size_t)unit * 2 * P.samp_stride;
        c.resid = c.samp + P.samp_stride;
    }
    if (P.heap_in_smem) {
        sp = (unsigned char*)(((uintptr_t)sp + 7) & ~(uintptr_t)7);
        c.psum = (u64*)sp; sp += (size_t)P.heap_entries * 8;
        c.karr = sp; sp += P.heap_entries;
        kfix = sp; sp += P.rice_stride;
        klpc = sp; sp += P.rice_stride;
    } else {
        c.psum = g_heap + (size_t)unit * P.heap_entries;
        c.karr = g_karr + (size_t)unit * (P.heap_entries + 2 * P.rice_stride);
        kfix = c.karr + P.heap_entries;
        klpc = kfix + P.rice_stride;
    }

    b200flac_plan plan;
    plan.type = BF_VERBATIM; plan.order = 0; plan.wasted = 0; plan.precision = 0; plan.shift = 0;
    plan.coding_method = 0; plan.partition_order = 0; plan.flags = 0; plan.bits = 0;
#pragma unroll
    for (int i = 0; i < BF_MAX_ORDER; i++) plan.coeffs[i] = 0;
    uint8_t* my_rice = rice_out + (size_t)unit * P.rice_stride;

    // ---- load, constant check, wasted bits ----
    u32 orv, differ;
    load_unit_samples<S>(c, pcm, d.pcm_off, cand, P, &orv, &differ);
    if (P.try_constant && differ == 0) {
        // flac.c:691-693, 813-830: CONSTANT, always written with wasted = 0 (H8)
        if (tid == 0) { plan.type = BF_CONSTANT; plan.bits = 8 + bps; plans[unit] = plan; }
        return;
    }
    const u32 wasted = orv ? (u32)(__ffs((int)orv) - 1) : 0u; // min ctz over non-zero samples = ctz(OR)
    if (wasted) {
        for (u32 base = tid * S; base < n; base += nt * S) {
            const u32 hi = min(base + S, n);
            for (u32 i = base; i < hi; i++) c.samp[PADI(i)] >>= wasted; // arithmetic (H10)
        }
    }
    __syncthreads();
    const u32 sub_bps = bps - wasted;
    const u32 hdr_bits = 8 + wasted; // 8-bit header; wasted>0 adds unary(wasted-1) = wasted bits

    // ---- FIXED ----
    u64 fixed_bits = 0;
    u32 fixed_order = 0;
    RiceChoice rfix; rfix.po = 0; rfix.under = 0; rfix.k0 = 0; rfix.method = 0; rfix.bits = 0;
    if (P.try_fixed) {
        u64 e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0;
        if (n > 4) {
            for (u32 base = tid * S; base < n; base += nt * S) {
                const u32 lo = max(base, 4u), hi = min(base + S, n);
                if (lo < hi) {
                    // differences of the four samples before lo
                    u32 a1 = (u32)c.samp[PADI(lo - 1)], a2 = (u32)c.samp[PADI(lo - 2)];
                    u32 a3 = (u32)c.samp[PADI(lo - 3)], a4 = (u32)c.samp[PADI(lo - 4)];
                    u32 p1 = a1 - a2;                      // d1[lo-1]
                    u32 p2 = p1 - (a2 - a3);               // d2[lo-1]
                    u32 p3 = p2 - ((a2 - a3) - (a3 - a4)); // d3[lo-1]
                    u32 prev = a1;
                    for (u32 i = lo; i < hi; i++) {
                        const u32 s = (u32)c.samp[PADI(i)];
                        const u32 d1 = s - prev, d2 = d1 - p1, d3 = d2 - p2, d4 = d3 - p3;
                        e0 += (u64)(long long)abs((int)s);
                        e1 += (u64)(long long)abs((int)d1);
                        e2 += (u64)(long long)abs((int)d2);
                        e3 += (u64)(long long)abs((int)d3);
                        e4 += (u64)(long long)abs((int)d4);
                        prev = s; p1 = d1; p2 = d2; p3 = d3;
                    }
                }
            }
            e0 = block_sum_u64(e0, red); e1 = block_sum_u64(e1, red); e2 = block_sum_u64(e2, red);
            e3 = block_sum_u64(e3, red); e4 = block_sum_u64(e4, red);
            // flac.c:877-893: first strict minimum over orders 0..4
            u64 best = e0;
            if (e1 < best) { best = e1; fixed_order = 1; }
            if (e2 < best) { best = e2; fixed_order = 2; }
            if (e3 < best) { best = e3; fixed_order = 3; }
            if (e4 < best) { best = e4; fixed_order = 4; }
        }
        for (u32 base = tid * S; base < n; base += nt * S) {
            const u32 lo = max(base, fixed_order), hi = min(base + S, n);
            for (u32 i = lo; i < hi; i++) c.resid[PADI(i)] = fixed_residual(c.samp, i, fixed_order);
        }
        __syncthreads();
        rice_search<S>(c, fixed_order, P, &rfix);
        save_rice(c, rfix, kfix);
        fixed_bits = hdr_bits + (u64)fixed_order * sub_bps + rfix.bits;
        __syncthreads();
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
        u64 best_bits = 0xFFFFFFFFull; // unsigned best_bits = UINT_MAX, flac.c:1079
        bool have = false;
        for (u32 o = o_first; o <= o_last; o++) {
            __syncthreads();
            if (tid < (int)o) s_q[tid] = mycoef[(o * (o - 1)) / 2 + tid];
            __syncthreads();
            const int shift = head.shift[o - 1];
            for (u32 base = tid * S; base < n; base += nt * S) {
                const u32 lo = max(base, o), hi = min(base + S, n);
                for (u32 i = lo; i < hi; i++) c.resid[PADI(i)] = lpc_residual(c.samp, i, o, s_q, shift);
            }
            __syncthreads();
            RiceChoice rc;
            rice_search<S>(c, o, P, &rc);
            const u64 bits = hdr_bits + (u64)o * sub_bps + 4 + 5 + (u64)o * head.precision + rc.bits;
            // exhaustive: first strict minimum over orders (flac.c:1102-1108); otherwise the only one
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
    const u32 vb = P.try_verbatim ? sub_bps * n : 0x7FFFFFFFu; // header NOT counted (H2)
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
        plan.bits = hdr_bits + sub_bps * n; // flac.c:832-854
    }
    if (tid == 0) plans[unit] = plan;
}
