// k_pack_v2.cuh -- bit packing for every block that fits shared memory and one pass of
// <= 512 threads (anything else falls back to k_pack_subframes in k_pack.cuh).
//
// Same bytes as k_pack_subframes.  One CTA per (frame, subframe slot):
//   * coalesced PCM load into shared memory, candidate signal formed on the way in;
//   * the chosen model's residual is recomputed in chunks of 8 (rolled loops, history window in
//     registers -- see k_analyze_v2.cuh), zig-zag folded and kept in shared memory;
//   * per-thread bit counts -> block scan -> every thread writes its run into a zeroed
//     shared-memory image of the subframe: words wholly inside the thread's bit range are plain
//     stores, its first and last word are merged with one shared atomicOr each;
//   * the image is copied to its final position with coalesced 32-bit stores; only its first and
//     last word (possibly shared with the neighbouring subframe) go through a global atomicOr.
#pragma once
#include "flac_common.cuh"
#include "k_analyze_v2.cuh"
#include "k_pack.cuh"

// MSB-first writer into the shared image.  The first word the thread touches is kept in
// `first_w` (merged later); every further completed word is exclusively the thread's.
struct RunSink {
    u32* words;
    u32 widx, widx0;
    u64 acc;
    u32 fill;
    u32 first_w;
    __device__ __forceinline__ void init(u32* base, u32 bitpos)
    {
        words = base; widx = bitpos >> 5; widx0 = widx; fill = bitpos & 31; acc = 0; first_w = 0;
    }
    __device__ __forceinline__ void put(u32 v, u32 nbits) // 1 <= nbits <= 32, v < 2^nbits
    {
        acc |= (u64)v << (64 - fill - nbits);
        fill += nbits;
        if (fill >= 32) {
            const u32 w = (u32)(acc >> 32);
            if (widx == widx0) first_w = w; else words[widx] = w;
            widx++; acc <<= 32; fill -= 32;
        }
    }
    __device__ __forceinline__ void zeros(u32 nz)
    {
        const u32 tot = fill + nz;
        if (tot >= 32) {
            const u32 w = (u32)(acc >> 32);
            if (widx == widx0) first_w = w; else words[widx] = w; // following all-zero words are already zero
            widx += tot >> 5; acc = 0; fill = tot & 31;
        } else fill = tot;
    }
    // merge the two boundary words
    __device__ __forceinline__ void finish()
    {
        if (widx == widx0) {
            const u32 w = (u32)(acc >> 32);
            if (w) atomicOr(words + widx0, w);
        } else {
            if (first_w) atomicOr(words + widx0, first_w);
            const u32 w = (u32)(acc >> 32);
            if (fill && w) atomicOr(words + widx, w);
        }
    }
};

// generic small writer (headers): everything through shared atomics
struct SmemSink {
    u32* words;
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

template <int NTMAX, int MINB>
__global__ void __launch_bounds__(NTMAX, MINB)
k_pack_v2(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd, bf_dev_params P, u32 S,
          const b200flac_plan* __restrict__ plans, const uint8_t* __restrict__ rice,
          const bf_frame_choice* __restrict__ choice, const u64* __restrict__ frame_off,
          u32* __restrict__ out_words, const u64* __restrict__ total, u64 capacity_bytes, u32 stage_words)
{
    if (*total + 16 > capacity_bytes) return;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ u64 red[40];
    __shared__ short s_q[BF_MAX_ORDER];
    __shared__ b200flac_plan s_plan;
    __shared__ bf_frame_choice s_choice;

    const int tid = threadIdx.x, nt = blockDim.x;
    const u32 frame = blockIdx.x / P.channels, slot = blockIdx.x % P.channels;
    if (tid == 0) s_choice = choice[frame];
    __syncthreads();
    if (slot >= s_choice.n_sub) return;
    const bf_frame_desc d = fd[frame];
    const u32 n = d.nsamp;
    const u32 unit = s_choice.unit[slot];
    const u32 cand = unit % P.K;
    if (tid == 0) s_plan = plans[unit];
    const uint8_t* krice = rice + (size_t)unit * P.rice_stride;
    const u32 bps = candidate_bps(cand, P);
    const u64 frame_bit0 = frame_off[frame] * 8;
    const u64 start = frame_bit0 + s_choice.bitoff[slot];

    V2Ctx c;
    c.n = n; c.S = S; c.base = (u32)tid * S; c.rs = nullptr; c.fine = nullptr; c.karr = nullptr;
    const size_t padn = (size_t)PADI(P.block_size) + 1;
    c.samp = (int*)dyn_smem;
    c.resid = c.samp + padn;
    u32* stage = (u32*)(c.resid + padn);
    const u32 base = c.base;

    u32 orv, diff;
    load_unit_v2(c, pcm, d.pcm_off, cand, P, &orv, &diff); // does not depend on the plan
    __syncthreads();
    const u32 ptype = s_plan.type, wasted = s_plan.wasted, order = s_plan.order;
    const u32 pbits = s_plan.bits;

    if (slot == 0 && tid == 0) {
        BitSink hs; hs.init(out_words, frame_bit0);
        put_frame_header(hs, d, P, s_choice.assignment);
        hs.flush();
    }
    const u32 sub_bps = bps - wasted;
    if (ptype == BF_CONSTANT) {
        if (tid == 0) {
            BitSink bs; bs.init(out_words, start);
            put_subframe_header(bs, 0, 0);
            bs.put_signed(ld_candidate(pcm, d.pcm_off, cand, P), bps);
            bs.flush();
        }
        return;
    }

    const u32 bit0 = (u32)(start & 31);          // position of the subframe inside image word 0
    const u32 nwords = (bit0 + pbits + 31) >> 5; // words of the image
    if (nwords > stage_words) __trap();          // cannot happen while VERBATIM is a candidate (host only then picks this kernel)
    for (u32 w = tid; w < nwords; w += nt) stage[w] = 0;
    if (wasted) for (u32 i = tid; i < n; i += nt) c.samp[PADI(i)] >>= wasted;
    if (ptype == BF_LPC && tid < BF_MAX_ORDER) s_q[tid] = tid < (int)order ? s_plan.coeffs[tid] : (short)0;
    __syncthreads();

    const u32 end = min(base + S, n);
    if (ptype == BF_VERBATIM) {
        if (tid == 0) {
            SmemSink bs; bs.init(stage, bit0);
            put_subframe_header_s(bs, 1, wasted);
            bs.flush();
        }
        if (base < n) {
            RunSink bs; bs.init(stage, bit0 + 8 + wasted + base * sub_bps);
            const u32 mask = sub_bps >= 32 ? 0xFFFFFFFFu : ((1u << sub_bps) - 1u);
            for (u32 i = base; i < end; i++) bs.put((u32)c.samp[PADI(i)] & mask, sub_bps);
            bs.finish();
        }
    } else {
        const u32 po = s_plan.partition_order, under = s_plan.flags & 1u, narrow = s_plan.flags & 2u;
        const u32 plen = n >> po;
        const u32 kbits = s_plan.coding_method ? 5u : 4u;
        u32 hdr_end = bit0 + 8 + wasted + order * sub_bps;
        if (ptype == BF_LPC) hdr_end += 4 + 5 + order * s_plan.precision;
        if (tid == 0) {
            SmemSink bs; bs.init(stage, bit0);
            if (ptype == BF_FIXED) put_subframe_header_s(bs, 0x8 | order, wasted);
            else put_subframe_header_s(bs, 0x20 | (order - 1), wasted);
            for (u32 i = 0; i < order; i++) bs.put_signed(c.samp[PADI(i)], sub_bps);
            if (ptype == BF_LPC) {
                bs.put(s_plan.precision - 1, 4);
                bs.put_signed(s_plan.shift, 5);
                for (u32 i = 0; i < order; i++) bs.put_signed(s_plan.coeffs[i], s_plan.precision);
            }
            bs.put(s_plan.coding_method, 2);
            bs.put(po, 4);
            bs.flush();
        }
        const u32 res0 = hdr_end + 6;
        const int shift = s_plan.shift;
        if (ptype == BF_FIXED) fixed_residual_v2(c, order);
        else if (order <= 12) { if (narrow) lpc_residual_v2<12, false>(c, s_q, shift); else lpc_residual_v2<12, true>(c, s_q, shift); }
        else { if (narrow) lpc_residual_v2<32, false>(c, s_q, shift); else lpc_residual_v2<32, true>(c, s_q, shift); }

        // ---- bit count of the run; the residual is zig-zag folded in place ----
        const u32 lo = max(base, order), hi = end;
        const bool have = lo < hi;
        const u32 p_first = (have && !under) ? lo / plen : 0u;
        u32 mybits = 0, lead = 0;
        if (have) {
            if (lo == order) lead = under ? 1u : (lo / plen + 1u);      // partitions 0..p_first (an empty leading one too)
            else if (!under && lo == p_first * plen) lead = 1;
            mybits = lead * kbits;
            u32 p = p_first;
            u32 next = under ? 0xFFFFFFFFu : (p + 1) * plen;
            u32 k = krice[p];
            for (u32 i0 = lo & ~(V2_CH - 1u); i0 < hi; i0 += V2_CH) {
                int* rp = CHP(c.resid, i0);
                if (i0 >= lo && i0 + V2_CH <= hi && i0 + V2_CH <= next) {
                    u32 a = 0;
#pragma unroll
                    for (int j = 0; j < V2_CH; j++) {
                        const u32 u = zigzag(rp[j]);
                        rp[j] = (int)u;
                        a += u >> k;
                    }
                    mybits += a + V2_CH * (1u + k);
                } else {
#pragma unroll
                    for (int j = 0; j < V2_CH; j++) {
                        const u32 i = i0 + j;
                        if (i >= lo && i < hi) {
                            if (i == next) { p++; next += plen; k = krice[p]; mybits += kbits; }
                            const u32 u = zigzag(rp[j]);
                            rp[j] = (int)u;
                            mybits += (u >> k) + 1u + k;
                        }
                    }
                }
                if (i0 + V2_CH == next && i0 + V2_CH < hi) { p++; next += plen; k = krice[p]; mybits += kbits; }
            }
        }
        u32 totalbits;
        const u32 off = block_exscan_u32(mybits, red, &totalbits);
        if (have) {
            RunSink bs; bs.init(stage, res0 + off);
            if (lo == order) { for (u32 q = 0; q < lead; q++) bs.put(krice[q], kbits); }
            else if (lead) bs.put(krice[p_first], kbits);
            u32 p = p_first;
            u32 next = under ? 0xFFFFFFFFu : (p + 1) * plen;
            u32 k = krice[p];
            for (u32 i0 = lo & ~(V2_CH - 1u); i0 < hi; i0 += V2_CH) {
                const int* rp = CHP(c.resid, i0);
                if (i0 >= lo && i0 + V2_CH <= hi && i0 + V2_CH <= next) {
                    const u32 kmask = (1u << k) - 1u, kone = 1u << k;
#pragma unroll
                    for (int j = 0; j < V2_CH; j++) {
                        const u32 u = (u32)rp[j];
                        const u32 msb = u >> k;
                        const u32 code = kone | (u & kmask);
                        if (msb + k + 1 <= 32) bs.put(code, msb + k + 1);
                        else { bs.zeros(msb); bs.put(code, k + 1); }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < V2_CH; j++) {
                        const u32 i = i0 + j;
                        if (i >= lo && i < hi) {
                            if (i == next) { p++; next += plen; k = krice[p]; bs.put(k, kbits); }
                            const u32 u = (u32)rp[j];
                            const u32 msb = u >> k;
                            const u32 code = (1u << k) | (u & ((1u << k) - 1u));
                            if (msb + k + 1 <= 32) bs.put(code, msb + k + 1);
                            else { bs.zeros(msb); bs.put(code, k + 1); }
                        }
                    }
                }
                if (i0 + V2_CH == next && i0 + V2_CH < hi) { p++; next += plen; k = krice[p]; bs.put(k, kbits); }
            }
            bs.finish();
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
