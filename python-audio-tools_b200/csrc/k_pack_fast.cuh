// k_pack_fast.cuh -- bit packing with register-resident residuals and a shared-memory
// staging buffer, for blocks that fit one pass (n <= blockDim * S).
//
// Same bytes as k_pack_subframes (k_pack.cuh).  Differences:
//   * the chosen model's residual is recomputed into registers with the unrolled
//     multiply-accumulate of k_analyze_fast.cuh;
//   * every thread writes its run into a zeroed shared-memory image of the subframe
//     (atomicOr on shared memory); the image is then copied to its final position with
//     coalesced 32-bit stores -- only the first and last word, which may be shared with the
//     neighbouring subframe, go through a global atomicOr;
//   * a code whose unary part and binary part fit 32 bits together is emitted in one put.
#pragma once
#include "flac_common.cuh"
#include "k_analyze_fast.cuh"
#include "k_pack.cuh"

struct SmemSink {
    u32* words;   // shared-memory image, big-endian value per word
    u32 widx;
    u64 acc;
    u32 fill;
    __device__ __forceinline__ void init(u32* base, u32 bitpos) { words = base; widx = bitpos >> 5; fill = bitpos & 31; acc = 0; }
    __device__ __forceinline__ void put(u32 v, u32 nbits)
    {
        if (nbits == 0) return;
        acc |= (u64)v << (64 - fill - nbits);
        fill += nbits;
        if (fill >= 32) {
            const u32 w = (u32)(acc >> 32);
            if (w) atomicOr(words + widx, w);
            widx++; acc <<= 32; fill -= 32;
        }
    }
    __device__ __forceinline__ void put_signed(int v, u32 nbits) { put(nbits >= 32 ? (u32)v : ((u32)v & ((1u << nbits) - 1u)), nbits); }
    __device__ __forceinline__ void zeros(u32 n)
    {
        const u32 tot = fill + n;
        if (tot >= 32) {
            const u32 top = (u32)(acc >> 32);
            if (top) atomicOr(words + widx, top);
            widx += tot >> 5; acc = 0; fill = tot & 31;
        } else fill = tot;
    }
    __device__ __forceinline__ void flush() { if (fill) { const u32 top = (u32)(acc >> 32); if (top) atomicOr(words + widx, top); } }
};

__device__ __forceinline__ void put_subframe_header_s(SmemSink& bs, u32 type_bits, u32 wasted)
{
    bs.put(type_bits & 0x3F, 7);
    if (wasted) { bs.put(1, 1); bs.zeros(wasted - 1); bs.put(1, 1); }
    else bs.put(0, 1);
}

template <int S>
__global__ void __launch_bounds__(512)
k_pack_fast(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd, bf_dev_params P,
            const b200flac_plan* __restrict__ plans, const uint8_t* __restrict__ rice,
            const bf_frame_choice* __restrict__ choice, const u64* __restrict__ frame_off,
            u32* __restrict__ out_words, const u64* __restrict__ total, u64 capacity_bytes, u32 stage_words)
{
    if (*total + 16 > capacity_bytes) return;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ u64 red[40];
    __shared__ short s_q[BF_MAX_ORDER];

    const int tid = threadIdx.x, nt = blockDim.x;
    const u32 frame = blockIdx.x / P.channels, slot = blockIdx.x % P.channels;
    const bf_frame_choice ch = choice[frame];
    if (slot >= ch.n_sub) return;
    const bf_frame_desc d = fd[frame];
    const u32 n = d.nsamp;
    const u32 unit = ch.unit[slot];
    const u32 cand = unit % P.K;
    const b200flac_plan plan = plans[unit];
    const uint8_t* krice = rice + (size_t)unit * P.rice_stride;
    const u32 bps = candidate_bps(cand, P);
    const u64 frame_bit0 = frame_off[frame] * 8;
    const u64 start = frame_bit0 + ch.bitoff[slot];
    const u32 base = (u32)tid * S;

    int* samp = (int*)dyn_smem;
    u32* stage = (u32*)(dyn_smem + ((size_t)PADI(P.block_size) + 1) * 4);

    if (slot == 0 && tid == 0) {
        BitSink hs; hs.init(out_words, frame_bit0);
        put_frame_header(hs, d, P, ch.assignment);
        hs.flush();
    }
    const u32 wasted = plan.wasted;
    const u32 sub_bps = bps - wasted;
    if (plan.type == BF_CONSTANT) {
        if (tid == 0) {
            BitSink bs; bs.init(out_words, start);
            put_subframe_header(bs, 0, 0);
            bs.put_signed(ld_candidate(pcm, d.pcm_off, cand, P), bps);
            bs.flush();
        }
        return;
    }

    const u32 bit0 = (u32)(start & 31);             // position of the subframe inside staging word 0
    const u32 nwords = (bit0 + plan.bits + 31) >> 5; // words of the image (<= stage_words, checked by the host bound)
    if (nwords > stage_words) __trap(); // cannot happen while VERBATIM is a candidate (host only then picks this kernel)
    for (u32 w = tid; w < nwords; w += nt) stage[w] = 0;

    int s[S];
    load_run<S>(pcm, d.pcm_off, base, n, cand, P, s);
#pragma unroll
    for (int j = 0; j < S; j++) {
        s[j] >>= wasted;
        if (base + j < n) samp[PADI(base + j)] = s[j];
    }
    if (plan.type == BF_LPC && tid < BF_MAX_ORDER) s_q[tid] = tid < (int)plan.order ? plan.coeffs[tid] : (short)0;
    __syncthreads();

    if (plan.type == BF_VERBATIM) {
        if (tid == 0) {
            SmemSink bs; bs.init(stage, bit0);
            put_subframe_header_s(bs, 1, wasted);
            bs.flush();
        }
        if (base < n) {
            SmemSink bs; bs.init(stage, bit0 + 8 + wasted + base * sub_bps);
#pragma unroll
            for (int j = 0; j < S; j++) if (base + j < n) bs.put_signed(s[j], sub_bps);
            bs.flush();
        }
    } else {
        const u32 order = plan.order;
        const u32 po = plan.partition_order, under = plan.flags & 1u;
        const u32 plen = n >> po;
        const u32 kbits = plan.coding_method ? 5u : 4u;
        u32 hdr_end = bit0 + 8 + wasted + order * sub_bps;
        if (plan.type == BF_LPC) hdr_end += 4 + 5 + order * plan.precision;
        if (tid == 0) {
            SmemSink bs; bs.init(stage, bit0);
            if (plan.type == BF_FIXED) put_subframe_header_s(bs, 0x8 | order, wasted);
            else put_subframe_header_s(bs, 0x20 | (order - 1), wasted);
            for (u32 i = 0; i < order; i++) bs.put_signed(samp[PADI(i)], sub_bps);
            if (plan.type == BF_LPC) {
                bs.put(plan.precision - 1, 4);
                bs.put_signed(plan.shift, 5);
                for (u32 i = 0; i < order; i++) bs.put_signed(plan.coeffs[i], plan.precision);
            }
            bs.put(plan.coding_method, 2);
            bs.put(po, 4);
            bs.flush();
        }
        const u32 res0 = hdr_end + 6;
        int r[S];
        if (plan.type == BF_FIXED) {
            fixed_residual_regs<S>(s, samp, base, n, order, r);
        } else {
            u32 sumq = 0;
            for (u32 j = 0; j < order; j++) sumq += (u32)abs((int)s_q[j]);
            const bool narrow = ((u64)sumq << (sub_bps - 1)) < (1ull << 31);
            if (narrow) lpc_residual_dispatch<S, false>(s, samp, base, n, order, s_q, plan.shift, r);
            else lpc_residual_dispatch<S, true>(s, samp, base, n, order, s_q, plan.shift, r);
        }
        const u32 lo = max(base, order), hi = min(base + (u32)S, n);
        u32 mybits = 0;
        u32 p_first = 0;
        if (lo < hi) {
            u32 p = under ? 0u : lo / plen;
            p_first = p;
            u32 next = under ? 0xFFFFFFFFu : (p + 1) * plen - base;
            u32 k = krice[p];
            if (lo == order) mybits += kbits * (under ? 1u : (lo / plen + 1u));
            else if (!under && lo == p * plen) mybits += kbits;
#pragma unroll
            for (int j = 0; j < S; j++) {
                const u32 i = base + j;
                if ((u32)j == next) { p++; next += plen; if (i < hi) { k = krice[p]; mybits += kbits; } }
                if (i >= lo && i < hi) mybits += (zigzag(r[j]) >> k) + 1u + k;
            }
        }
        u32 totalbits;
        const u32 off = block_exscan_u32(mybits, red, &totalbits);
        if (lo < hi) {
            SmemSink bs; bs.init(stage, res0 + off);
            u32 p = p_first;
            u32 next = under ? 0xFFFFFFFFu : (p + 1) * plen - base;
            u32 k = krice[p];
            if (lo == order) {
                const u32 lead = under ? 1u : (lo / plen + 1u);
                for (u32 q = 0; q < lead; q++) bs.put(krice[q], kbits);
            } else if (!under && lo == p * plen) {
                bs.put(k, kbits);
            }
#pragma unroll
            for (int j = 0; j < S; j++) {
                const u32 i = base + j;
                if ((u32)j == next) { p++; next += plen; if (i < hi) { k = krice[p]; bs.put(k, kbits); } }
                if (i >= lo && i < hi) {
                    const u32 u = zigzag(r[j]);
                    const u32 msb = u >> k;
                    const u32 code = (1u << k) | (u & ((1u << k) - 1u));
                    if (msb + k + 1 <= 32) bs.put(code, msb + k + 1);
                    else { bs.zeros(msb); bs.put(code, k + 1); }
                }
            }
            bs.flush();
        }
        if (tid == 0) {
            u32 first_trailing;
            if (n == order) first_trailing = 0;
            else if (under) first_trailing = 1;
            else first_trailing = 1u << po;
            if (first_trailing < (1u << po)) {
                SmemSink bs; bs.init(stage, res0 + totalbits);
                for (u32 p = first_trailing; p < (1u << po); p++) bs.put(krice[p], kbits);
                bs.flush();
            }
        }
    }
    __syncthreads();
    // ---- copy the image to its final position ----
    u32* dst = out_words + (start >> 5);
    for (u32 w = tid; w < nwords; w += nt) {
        const u32 v = stage[w];
        if (w == 0 || w == nwords - 1) { if (v) atomicOr(dst + w, __byte_perm(v, 0, 0x0123)); }
        else dst[w] = __byte_perm(v, 0, 0x0123);
    }
}
