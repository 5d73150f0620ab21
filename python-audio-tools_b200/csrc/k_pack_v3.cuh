// k_pack_v3.cuh -- frame packing: one CTA builds a WHOLE frame in shared memory and writes it out.
//
// Same bytes as k_pack_v2 + k_frame_crc16 (and the reference: frame header + CRC-8 flac.c:412-518,
// subframes :813-854, 897-915, 977-1015, Rice coding :1406-1434, CRC-16 :530,668-670).  What differs:
//   * the CTA is two thread groups, each packing one subframe at a time into a shared-memory image
//     of the frame at its final bit offset (the subframe sizes are known from the analysis);
//   * the chosen model's residual is computed IN PLACE over the samples (history first, then a group
//     barrier), zig-zag folded and bit-counted in the same pass;
//   * the frame's CRC-16 is computed from the image: 60-byte chunks (15 words: consecutive threads
//     hit different banks), one table CRC per thread, chunk results moved to their position with one
//     GF(2) multiply by a tabulated power of x, XOR-reduced;
//   * the finished frame goes to its byte offset in the output with coalesced 32-bit stores (a byte
//     permute handles both the byte order and the misalignment), head and tail bytes singly: no
//     zeroing of the output, no global atomics, no separate CRC kernel.
#pragma once
#include "flac_common.cuh"
#include "k_pack.cuh"
#include "k_pack_v2.cuh"
#include "k_analyze_v3.cuh"

#define P3_CHUNK_WORDS 15      // odd: consecutive threads' chunks start in different banks; ~1 chunk per thread
#define P3_CHUNK_BYTES 60

__device__ __forceinline__ void p3_group_bar(u32 g, u32 gt)
{
    asm volatile("bar.sync %0, %1;\n" ::"r"(g + 1), "r"(gt) : "memory");
}

// exclusive prefix sum of one u32 per thread over a group of gt threads; red: >= 16 words of the group
__device__ __forceinline__ u32 p3_group_exscan(u32 v, u32* red, u32 g, u32 gt, u32 gtid, u32* total)
{
    const u32 lane = gtid & 31, warp = gtid >> 5, nw = gt >> 5;
    u32 inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= (u32)o) inc += t;
    }
    if (lane == 31) red[warp] = inc;
    p3_group_bar(g, gt);
    u32 off = 0, tot = 0;
    for (u32 w = 0; w < nw; w++) { const u32 t = red[w]; off += w < warp ? t : 0u; tot += t; }
    *total = tot;              // (the caller alternates between two `red` buffers: no second barrier)
    return off + inc - v;
}

// MSB-first writer of a thread's run into the shared image, pending bits RIGHT-aligned in a 64-bit
// register: appending is a shift and an OR, a completed word is one funnel shift.  Bits above `fill`
// in acc are stale and never read.  Every completed word is OR-merged (shared atomicOr on a
// zero-initialised image), so the word a run shares with its predecessor needs no special case.
struct RunSink3 {
    u32* words;
    u32 widx;
    u64 acc;
    u32 fill;
    __device__ __forceinline__ void init(u32* base, u32 bitpos)
    {
        words = base; widx = bitpos >> 5; fill = bitpos & 31; acc = 0;
    }
    __device__ __forceinline__ void put(u32 v, u32 nbits) // 1 <= nbits <= 32, v < 2^nbits
    {
        acc = (acc << nbits) | (u64)v;
        fill += nbits;
        if (fill >= 32) {
            fill -= 32;
            atomicOr(words + widx, (u32)(acc >> fill));
            widx++;
        }
    }
    __device__ __forceinline__ void zeros(u32 nz)
    {
        if (fill + nz < 64) {
            acc <<= nz;
            fill += nz;
            if (fill >= 32) { fill -= 32; atomicOr(words + widx, (u32)(acc >> fill)); widx++; }
        } else {
            // the pending bits complete a word with zeros; whole zero words follow
            if (fill) atomicOr(words + widx, (u32)(acc << (32 - fill)));
            const u32 tot = fill + nz;
            widx += tot >> 5; fill = tot & 31; acc = 0;
        }
    }
    __device__ __forceinline__ void finish()
    {
        if (fill) atomicOr(words + widx, (u32)(acc << (32 - fill)));
    }
};

// one field of 1..32 bits (v < 2^nbits) at a known bit position of the zero-initialised image.  A single,
// not inlined copy serves every header field, warm-up sample and coefficient: the inlined writers these
// replace were ~150 instructions of hot code executed by a handful of threads.
#ifndef P3_FIELD_INLINE
#define P3_FIELD_INLINE 1      /* (a real call cost 0.28 ms per hour: measured) */
#endif
#if P3_FIELD_INLINE
__device__ __forceinline__
#else
__device__ __noinline__
#endif
void p3_field(u32* img, u32 bitpos, u32 v, u32 nbits)
{
    const u32 w = bitpos >> 5, o = bitpos & 31;
    const u64 x = (u64)v << (64 - o - nbits);
    if ((u32)(x >> 32)) atomicOr(img + w, (u32)(x >> 32));
    if ((u32)x) atomicOr(img + w + 1, (u32)x);
}
__device__ __forceinline__ void p3_field_signed(u32* img, u32 bitpos, int v, u32 nbits)
{
    p3_field(img, bitpos, nbits >= 32 ? (u32)v : ((u32)v & ((1u << nbits) - 1u)), nbits);
}
// subframe header (flac.c:820-826 / 897-905 / 977-985): 0 | type (6 bits) | wasted flag, then the unary
// count of wasted bits -- zeros that are already there and a one
__device__ __forceinline__ void p3_subframe_header(u32* img, u32 bit0, u32 type_bits, u32 wasted)
{
    p3_field(img, bit0, ((type_bits & 0x3F) << 1) | (wasted ? 1u : 0u), 8);
    if (wasted) p3_field(img, bit0 + 8 + wasted - 1, 1, 1);
}

// a Rice code longer than 32 bits: a run of zeros, then the stop bit and the k low bits (rare)
__device__ __forceinline__ void p3_put_long(RunSink3& bs, u32 msb, u32 code, u32 k)
{
    bs.zeros(msb);
    bs.put(code, k + 1);
}

// a * b modulo the CRC-16 polynomial: 16 shift-and-add steps, then the upper half is folded back with
// the byte tables (the CRC of the two-byte message `hi` is hi * x^16 mod P)
__device__ __forceinline__ u32 p3_gf16_mul(u32 a, u32 b, const unsigned short* tab)
{
    u32 r = 0;
#pragma unroll 1
    for (int i = 0; i < 16; i++) r ^= ((b >> i) & 1u) ? (a << i) : 0u;       // (once per thread: rolled, small code)
    const u32 hi = r >> 16;
    return (r & 0xFFFFu) ^ (u32)tab[256 + (hi >> 8)] ^ (u32)tab[hi & 0xFF];
}

struct P3Shared {
    b200flac_plan plan[2];
    bf_frame_choice choice;
    u32 red[2][2][16];
    short q[2][BF_MAX_ORDER];
    u32 crc_part[32];
    u32 crc_last;
};

// residual of the chosen model over the thread's run, in place, zig-zag folded; returns the run's
// bit count.  The caller has already read what it needs from the samples that other threads overwrite.
struct P3Count {
    u32 lo, hi, plen, kbits, under, order;
    const uint8_t* krice;
};

__device__ __forceinline__ void p3_count_chunk(const u32 (&u)[8], const int* __restrict__ buf, u32 i0, const P3Count& c,
                                               u32& p, u32& next, u32& k, u32& mybits)
{
    if (i0 >= c.lo && i0 + 8 <= c.hi && i0 + 8 <= next) {
        u32 a = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) a += u[j] >> k;
        mybits += a + 8 * (1u + k);
    } else {
        // rare (warm-up samples, a partition boundary inside the chunk, the block's tail): a rolled loop
        // over the folded residuals just stored, to keep the hot code small
#pragma unroll 1
        for (u32 i = max(i0, c.lo); i < min(i0 + 8, c.hi); i++) {
            if (i == next) { p++; next += c.plen; k = c.krice[p]; mybits += c.kbits; }
            mybits += ((u32)buf[V3_SK(i)] >> k) + 1u + k;
        }
    }
    if (i0 + 8 == next && i0 + 8 < c.hi) { p++; next += c.plen; k = c.krice[p]; mybits += c.kbits; }
}

template <int OG, bool WIDE>
__device__ __forceinline__ u32 p3_lpc_inplace(int* __restrict__ buf, u32 base, u32 end, const short* q_sm, int shift,
                                              const P3Count& c, u32 p, u32 next, u32 k, u32 g, u32 gt)
{
    int q[OG];
#pragma unroll
    for (int t = 0; t < OG; t++) q[t] = q_sm[t];
    int w[OG + 8];
#pragma unroll
    for (int t = 0; t < OG; t += 4) {
        int4 h = make_int4(0, 0, 0, 0);
        if (base < end && base >= (u32)(OG - t)) h = *(const int4*)(buf + V3_SK(base - (OG - t)));
        w[t] = h.x; w[t + 1] = h.y; w[t + 2] = h.z; w[t + 3] = h.w;
    }
    p3_group_bar(g, gt);          // every thread holds its history: the samples may be overwritten now
    u32 mybits = 0;
    for (u32 i0 = base; i0 < end; i0 += 8) {
        const int4 va = *(const int4*)(buf + V3_SK(i0));
        const int4 vb = *(const int4*)(buf + V3_SK(i0) + 4);
        w[OG + 0] = va.x; w[OG + 1] = va.y; w[OG + 2] = va.z; w[OG + 3] = va.w;
        w[OG + 4] = vb.x; w[OG + 5] = vb.y; w[OG + 6] = vb.z; w[OG + 7] = vb.w;
        u32 u[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
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
            u[j] = zigzag((int)((u32)w[OG + j] - (u32)pred));
        }
        *(uint4*)(buf + V3_SK(i0)) = make_uint4(u[0], u[1], u[2], u[3]);
        *(uint4*)(buf + V3_SK(i0) + 4) = make_uint4(u[4], u[5], u[6], u[7]);
        p3_count_chunk(u, buf, i0, c, p, next, k, mybits);
#pragma unroll
        for (int t = 0; t < OG; t++) w[t] = w[t + 8];
    }
    return mybits;
}

__device__ __forceinline__ u32 p3_fixed_inplace(int* __restrict__ buf, u32 base, u32 end, u32 order,
                                                const P3Count& c, u32 p, u32 next, u32 k, u32 g, u32 gt)
{
    // closed form of the iterated differences (v3_fixed_coef): one loop for the five orders
    const V3FixedCoef fc = v3_fixed_coef(order);
    int4 h = make_int4(0, 0, 0, 0);                       // samples base-4 .. base-1
    if (base && base < end) h = *(const int4*)(buf + V3_SK(base - 4));
    p3_group_bar(g, gt);
    u32 mybits = 0;
    for (u32 i0 = base; i0 < end; i0 += 8) {
        const int4 va = *(const int4*)(buf + V3_SK(i0));
        const int4 vb = *(const int4*)(buf + V3_SK(i0) + 4);
        const int w[12] = {h.x, h.y, h.z, h.w, va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
        u32 u[8];
#pragma unroll
        for (int j = 0; j < 8; j++)
            u[j] = zigzag(w[j + 4] + fc.c1 * w[j + 3] + fc.c2 * w[j + 2] + fc.c3 * w[j + 1] + fc.c4 * w[j]);
        *(uint4*)(buf + V3_SK(i0)) = make_uint4(u[0], u[1], u[2], u[3]);
        *(uint4*)(buf + V3_SK(i0) + 4) = make_uint4(u[4], u[5], u[6], u[7]);
        p3_count_chunk(u, buf, i0, c, p, next, k, mybits);
        h = vb;
    }
    return mybits;
}

// host and device agree on the dynamic shared memory through this
__host__ __device__ inline size_t p3_smem_bytes(u32 block_size, u32 img_words)
{
    const size_t padn = ((size_t)V3_SK(block_size) + 8 + 3) & ~(size_t)3;
    return 2 * padn * 4 + (size_t)(img_words + 12) * 4 + 2048 + 32;
}

// blockDim.x = 2 * gt; gt * S >= block_size; S a multiple of 8.
//   crc_tab[4][256]: CRC-16 of one byte followed by 0..3 zero bytes; crc_pow[0..CHUNK] = x^(8 r), crc_pow[CHUNK + 1 + j] = x^(8 * CHUNK * j), CHUNK = 60 bytes, mod the
//   CRC-16 polynomial (built by the host).
template <int NTMAX, int MINB, int SC>
__global__ void __launch_bounds__(NTMAX, MINB)
k_pack_v3(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd, bf_dev_params P, u32 S_rt,
          const b200flac_plan* __restrict__ plans, const uint8_t* __restrict__ rice,
          const bf_frame_choice* __restrict__ choice, const u64* __restrict__ frame_off,
          uint8_t* __restrict__ out, const u64* __restrict__ total, u64 capacity_bytes, u32 img_words,
          const unsigned short* __restrict__ crc_tab, const unsigned short* __restrict__ crc_pow)
{
    if (*total + 16 > capacity_bytes) return;
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ P3Shared sh;
    const u32 S = SC ? (u32)SC : S_rt;       // samples per thread: a compile-time 32 for the common shapes

    const u32 tid = threadIdx.x, nt = blockDim.x, gt = nt >> 1;
    const u32 g = tid >= gt ? 1u : 0u, gtid = tid - g * gt;
    const u32 frame = blockIdx.x;
    const size_t padn = ((size_t)V3_SK(P.block_size) + 8 + 3) & ~(size_t)3;     // keeps the image 16-byte aligned
    int* buf = (int*)dyn_smem + (size_t)g * padn;
    u32* img = (u32*)((int*)dyn_smem + 2 * padn);
    unsigned short* tab = (unsigned short*)(img + ((img_words + 8 + 3) & ~3u));   // past the zeroing's overshoot; 4 x 256 entries

    if (tid == 0) sh.choice = choice[frame];
#pragma unroll 1
    for (u32 t = tid; t < 512; t += nt) ((u32*)tab)[t] = ((const u32*)crc_tab)[t];      // 4 x 256 entries
    const bf_frame_desc d = fd[frame];
    const u32 n = d.nsamp;
    // the whole image is cleared (not just this frame's extent) so that one barrier covers everything
    for (u32 w = tid * 4; w < img_words + 4; w += nt * 4) *(uint4*)(img + w) = make_uint4(0, 0, 0, 0);
    __syncthreads();
    const u32 frame_bytes = sh.choice.frame_bytes, n_sub = sh.choice.n_sub;
    const u32 nwords = (frame_bytes + 3) >> 2;
    if (nwords + 2 > img_words + 4) __trap();    // cannot happen: the image is sized for the largest frame
    if (tid < 4 && sh.choice.header_words[tid]) atomicOr(img + tid, sh.choice.header_words[tid]);   // frame header + CRC-8 (k_frame_select)

    const u32 base = gtid * S;
    const u32 end = min(base + S, n);
    for (u32 slot0 = 0; slot0 < n_sub; slot0 += 2) {
        const u32 slot = slot0 + g;
        if (slot >= n_sub) break;                // this group is done (group-uniform)
        const u32 unit = sh.choice.unit[slot];
        const u32 cand = P.K == 4 ? unit & 3u : unit % P.K;
        const u32 bps = candidate_bps(cand, P);
        const u32 bit0 = sh.choice.bitoff[slot];
        const uint8_t* krice = rice + (size_t)unit * P.rice_stride;
        if (gtid < 19) ((u32*)&sh.plan[g])[gtid] = ((const u32*)(plans + unit))[gtid];     // 76 bytes as 19 words
        if (gtid < BF_MAX_ORDER) {
            const b200flac_plan* gp = plans + unit;
            sh.q[g][gtid] = (gp->type == BF_LPC && gtid < gp->order) ? gp->coeffs[gtid] : (short)0;
        }

        // ---- samples of the candidate, coalesced ----
        if (P.stereo && P.bytes_ps == 2) {
            const int coef = cand == 0 ? 0x0001 : cand == 1 ? 0x0100 : cand == 2 ? 0x0101 : 0xFF01;
            const int shv = cand == 2 ? 1 : 0;
            const uint8_t* src = pcm + d.pcm_off * 4;
            if ((((uintptr_t)src) & 15) == 0 && (n & 3) == 0) {
                const uint4* s4 = (const uint4*)src;
                const u32 stride = gt * 4;
                u32 i = gtid * 4;
                for (; i + 3 * stride < n; i += 4 * stride) {
                    uint4 w4[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) w4[q] = __ldg(s4 + ((i + q * stride) >> 2));
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        int4 v;
                        v.x = __dp2a_lo((int)w4[q].x, coef, 0) >> shv; v.y = __dp2a_lo((int)w4[q].y, coef, 0) >> shv;
                        v.z = __dp2a_lo((int)w4[q].z, coef, 0) >> shv; v.w = __dp2a_lo((int)w4[q].w, coef, 0) >> shv;
                        *(int4*)(buf + V3_SK(i + q * stride)) = v;
                    }
                }
                for (; i < n; i += stride) {
                    const uint4 w1 = __ldg(s4 + (i >> 2));
                    int4 v;
                    v.x = __dp2a_lo((int)w1.x, coef, 0) >> shv; v.y = __dp2a_lo((int)w1.y, coef, 0) >> shv;
                    v.z = __dp2a_lo((int)w1.z, coef, 0) >> shv; v.w = __dp2a_lo((int)w1.w, coef, 0) >> shv;
                    *(int4*)(buf + V3_SK(i)) = v;
                }
            } else {
                const u32* s1 = (const u32*)src;
                for (u32 i = gtid; i < n; i += gt) buf[V3_SK(i)] = __dp2a_lo((int)__ldg(s1 + i), coef, 0) >> shv;
            }
        } else if (P.stereo && P.bytes_ps == 3 && (n & 7u) == 0 && (((uintptr_t)(pcm + d.pcm_off * 6)) & 15) == 0) {
            u32 o_ = 0, d_ = 0;
            load_stereo24_skewed(pcm + d.pcm_off * 6, buf, n, cand, gtid, gt, 0, &o_, &d_);
        } else {
            for (u32 i = gtid; i < n; i += gt) buf[V3_SK(i)] = ld_candidate(pcm, d.pcm_off + i, cand, P);
        }
        p3_group_bar(g, gt);
        const u32 ptype = sh.plan[g].type, wasted = sh.plan[g].wasted, order = sh.plan[g].order;
        const u32 sub_bps = bps - wasted;

        if (ptype == BF_CONSTANT) {
            if (gtid == 0) {
                p3_subframe_header(img, bit0, 0, 0);
                p3_field_signed(img, bit0 + 8, buf[0], bps);
            }
        } else {
            if (wasted) {
                for (u32 i = gtid; i < n; i += gt) buf[V3_SK(i)] >>= wasted;
                p3_group_bar(g, gt);
            }
            if (ptype == BF_VERBATIM) {
                if (gtid == 0) p3_subframe_header(img, bit0, 1, wasted);
                if (base < n) {
                    RunSink3 bs; bs.init(img, bit0 + 8 + wasted + base * sub_bps);
                    const u32 mask = sub_bps >= 32 ? 0xFFFFFFFFu : ((1u << sub_bps) - 1u);
                    for (u32 i = base; i < end; i++) bs.put((u32)buf[V3_SK(i)] & mask, sub_bps);
                    bs.finish();
                }
            } else {
                const u32 po = sh.plan[g].partition_order, under = sh.plan[g].flags & 1u, narrow = sh.plan[g].flags & 2u;
                const u32 plen = n >> po;
                const u32 kbits = sh.plan[g].coding_method ? 5u : 4u;
                u32 hdr_end = bit0 + 8 + wasted + order * sub_bps;
                if (ptype == BF_LPC) hdr_end += 4 + 5 + order * sh.plan[g].precision;
                // header, warm-up samples and coefficients, every field at its known bit position: lane i
                // writes warm-up sample i and coefficient i (read before anyone passes the barrier below)
                if (gtid < order) {
                    p3_field_signed(img, bit0 + 8 + wasted + gtid * sub_bps, buf[V3_SK(gtid)], sub_bps);
                    if (ptype == BF_LPC) {
                        const u32 prec = sh.plan[g].precision;
                        p3_field_signed(img, bit0 + 8 + wasted + order * sub_bps + 9 + gtid * prec, sh.plan[g].coeffs[gtid], prec);
                    }
                }
                if (gtid == 0) {
                    p3_subframe_header(img, bit0, ptype == BF_FIXED ? (0x8 | order) : (0x20 | (order - 1)), wasted);
                    if (ptype == BF_LPC)    // 4 bits of precision - 1, 5 bits of signed shift
                        p3_field(img, bit0 + 8 + wasted + order * sub_bps, ((sh.plan[g].precision - 1) << 5) | ((u32)sh.plan[g].shift & 31u), 9);
                    p3_field(img, hdr_end, (sh.plan[g].coding_method << 4) | po, 6);
                }
                const u32 res0 = hdr_end + 6;
                const int shift = sh.plan[g].shift;

                // ---- residual in place + bit count of the run ----
                P3Count c;
                c.lo = max(base, order); c.hi = end; c.plen = plen; c.kbits = kbits; c.under = under; c.order = order;
                c.krice = krice;
                const bool have = c.lo < c.hi;
                const u32 p_first = (have && !under) ? c.lo / plen : 0u;
                u32 lead = 0;
                if (have) {
                    if (c.lo == order) lead = under ? 1u : (p_first + 1u);        // partitions 0..p_first (an empty leading one too)
                    else if (!under && c.lo == p_first * plen) lead = 1;
                }
                const u32 next0 = under ? 0xFFFFFFFFu : (p_first + 1) * plen;
                const u32 k0 = have ? (u32)krice[p_first] : 0u;
                const u32 rend = base < n ? end : base;   // threads past the block only take part in the barriers
                u32 mybits;
                if (ptype == BF_FIXED) mybits = p3_fixed_inplace(buf, base, rend, order, c, p_first, next0, k0, g, gt);
                else if (order <= 12) mybits = narrow ? p3_lpc_inplace<12, false>(buf, base, rend, sh.q[g], shift, c, p_first, next0, k0, g, gt)
                                                      : p3_lpc_inplace<12, true>(buf, base, rend, sh.q[g], shift, c, p_first, next0, k0, g, gt);
                else mybits = narrow ? p3_lpc_inplace<32, false>(buf, base, rend, sh.q[g], shift, c, p_first, next0, k0, g, gt)
                                     : p3_lpc_inplace<32, true>(buf, base, rend, sh.q[g], shift, c, p_first, next0, k0, g, gt);
                mybits = have ? mybits + lead * kbits : 0u;

                u32 totalbits;
                const u32 off = p3_group_exscan(mybits, sh.red[g][(slot0 >> 1) & 1], g, gt, gtid, &totalbits);
                if (have) {
                    RunSink3 bs; bs.init(img, res0 + off);
                    if (c.lo == order) { for (u32 q = 0; q < lead; q++) bs.put(krice[q], kbits); }
                    else if (lead) bs.put(krice[p_first], kbits);
                    u32 p = p_first;
                    u32 next = next0;
                    u32 k = k0;
                    for (u32 i0 = c.lo & ~7u; i0 < c.hi; i0 += 8) {
                        const uint4 ua = *(const uint4*)(buf + V3_SK(i0));
                        const uint4 ub = *(const uint4*)(buf + V3_SK(i0) + 4);
                        const u32 us[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
                        if (i0 >= c.lo && i0 + 8 <= c.hi && i0 + 8 <= next) {
                            const u32 kmask = (1u << k) - 1u, kone = 1u << k;
#pragma unroll
                            for (int j = 0; j < 8; j++) {
                                const u32 u = us[j];
                                const u32 msb = u >> k;
                                const u32 code = kone | (u & kmask);
                                if (msb + k + 1 <= 32) bs.put(code, msb + k + 1);
                                else p3_put_long(bs, msb, code, k);
                            }
                        } else {
#pragma unroll 1
                            for (u32 i = max(i0, c.lo); i < min(i0 + 8, c.hi); i++) {
                                if (i == next) { p++; next += plen; k = krice[p]; bs.put(k, kbits); }
                                const u32 u = (u32)buf[V3_SK(i)];
                                const u32 msb = u >> k;
                                const u32 code = (1u << k) | (u & ((1u << k) - 1u));
                                if (msb + k + 1 <= 32) bs.put(code, msb + k + 1);
                                else p3_put_long(bs, msb, code, k);
                            }
                        }
                        if (i0 + 8 == next && i0 + 8 < c.hi) { p++; next += plen; k = krice[p]; bs.put(k, kbits); }
                    }
                    bs.finish();
                }
                if (gtid == 0) {
                    u32 first_trailing;
                    if (n == order) first_trailing = 0;
                    else if (under) first_trailing = 1;
                    else first_trailing = 1u << po;
#pragma unroll 1
                    for (u32 p = first_trailing; p < (1u << po); p++) p3_field(img, res0 + totalbits + (p - first_trailing) * kbits, krice[p], kbits);
                }
            }
        }
        if (slot0 + 2 < n_sub) p3_group_bar(g, gt);      // the group's buffer and plan are free for its next subframe
    }
    __syncthreads();

    // ---- CRC-16 of the frame (every byte before it), flac.c:530,668-670 ----
    const u32 nb = frame_bytes - 2;
    const u32 T = (nb + P3_CHUNK_BYTES - 1) / P3_CHUNK_BYTES;
    u32 acc = 0;
    for (u32 c = tid; c < T; c += nt) {
        const u32 b0 = c * P3_CHUNK_BYTES, b1 = min(b0 + P3_CHUNK_BYTES, nb);
        const u32* wp = img + c * P3_CHUNK_WORDS;
        u32 crc = 0;
        u32 b = b0;
#pragma unroll 1
        for (; b + 4 <= b1; b += 4) {
            // four bytes at once: tab[256 n + x] = CRC of byte x followed by n zero bytes, so the four lookups
            // of a word are independent of each other (one shared-memory latency per word on the serial chain)
            const u32 a = *wp++ ^ (crc << 16);
            crc = (u32)tab[768 + (a >> 24)] ^ (u32)tab[512 + ((a >> 16) & 0xFF)] ^ (u32)tab[256 + ((a >> 8) & 0xFF)] ^ (u32)tab[a & 0xFF];
        }
        if (b < b1) {
            const u32 w = *wp;
            for (u32 i = 0; b < b1; b++, i++) crc = ((crc << 8) & 0xFFFF) ^ tab[(crc >> 8) ^ ((w >> (24 - 8 * i)) & 0xFF)];
        }
        if (c + 1 < T) acc ^= p3_gf16_mul(crc, (u32)crc_pow[P3_CHUNK_BYTES + 1 + (T - 2 - c)], tab);
        else sh.crc_last = crc;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc ^= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if ((tid & 31) == 0) sh.crc_part[tid >> 5] = acc;
    __syncthreads();
    // ---- the frame to its byte offset in the output; thread 0 finishes the CRC and writes its two
    // bytes itself, everybody else moves the nb bytes before it ----
    uint8_t* dst = out + frame_off[frame];
    if (tid == 0) {
        u32 a = 0;
        for (u32 w = 0; w < (nt >> 5); w++) a ^= sh.crc_part[w];
        const u32 r = nb - (T - 1) * P3_CHUNK_BYTES;          // bytes of the last chunk, 1..60
        const u32 crc = p3_gf16_mul(a, (u32)crc_pow[r], tab) ^ sh.crc_last;
        dst[nb] = (uint8_t)(crc >> 8);
        dst[nb + 1] = (uint8_t)crc;
    }
    const u32 h = (4u - (u32)((uintptr_t)dst & 3)) & 3u;      // bytes before the first aligned word
    const u32 hb = min(h, nb);
    if (tid < hb) dst[tid] = (uint8_t)(img[0] >> (24 - 8 * tid));
    const u32 body = (nb - hb) >> 2;
    u32* dw = (u32*)(dst + hb);
    // output byte i of word j is image byte h + 4 j + i: bytes h..3 of img[j], then 0..h-1 of img[j + 1]
    const u32 sel = h == 0 ? 0x0123u : h == 1 ? 0x7012u : h == 2 ? 0x6701u : 0x5670u;
#ifndef P3_COPY_UNROLL
#define P3_COPY_UNROLL 1
#endif
#define P3_PRAGMA_(x) _Pragma(#x)
#define P3_UNROLL_(n) P3_PRAGMA_(unroll n)
    P3_UNROLL_(P3_COPY_UNROLL)
    for (u32 j = tid; j < body; j += nt) dw[j] = __byte_perm(img[j], img[j + 1], sel);
    const u32 done = hb + 4 * body;
    if (tid < nb - done) {
        const u32 b = done + tid;
        dst[b] = (uint8_t)(img[b >> 2] >> (24 - 8 * (b & 3)));
    }
}
