// k_pack.cuh -- frame assembly: stereo decision, offsets, and the generic bit packing / CRC kernels
// (k_pack_v3.cuh packs whole frames with the CRC-16 fused; it uses k_frame_select and k_scan_offsets
// from here and leaves the rest to the shapes it does not take).
//
//   k_frame_select   stereo assignment by exact bit counts + frame sizes
//                    flacenc_write_frame                  flac.c:532-666
//   k_scan_offsets   exclusive prefix sum of frame sizes  (current_offset, flac.c:266)
//   k_zero_output    clears the exact output extent (the bit writer ORs into zeros)
//   k_pack_subframes one CTA per (frame, subframe slot): header, warm-up, LPC
//                    parameters, Rice-coded residual at its final bit position
//                    flac.c:412-518 (frame header), 813-854, 897-915, 977-1015, 1406-1434
//   k_frame_crc16    CRC-16 of every frame, one warp per frame (flac.c:530,668-670)
#pragma once
#include "flac_common.cuh"
#include "k_analyze.cuh"

// UTF-8 style frame-number length, flac.c:1531-1566
__device__ __forceinline__ u32 utf8_bytes(u32 v)
{
    if (v <= 0x7F) return 1;
    if (v <= 0x7FF) return 2;
    if (v <= 0xFFFF) return 3;
    if (v <= 0x1FFFFF) return 4;
    if (v <= 0x3FFFFFF) return 5;
    return 6;
}

// block-size code, flac.c:427-449
__device__ __forceinline__ u32 block_size_code(u32 bs)
{
    switch (bs) {
    case 192: return 1; case 576: return 2; case 1152: return 3; case 2304: return 4;
    case 4608: return 5; case 256: return 8; case 512: return 9; case 1024: return 10;
    case 2048: return 11; case 4096: return 12; case 8192: return 13; case 16384: return 14;
    case 32768: return 15;
    default: return bs < 256 ? 6 : (bs < 65536 ? 7 : 0);
    }
}

// the bytes of a frame header, flac.c:488-517, CRC-8 included; returns their number (6..16)
__device__ inline u32 build_frame_header(uint8_t (&h)[16], const bf_frame_desc& d, const bf_dev_params& P, u32 assignment)
{
    u32 nb = 0;
    const u32 bsc = block_size_code(d.nsamp);
    h[nb++] = 0xFF;
    h[nb++] = 0xF8;                                   // sync, reserved 0, blocking strategy 0
    h[nb++] = (uint8_t)((bsc << 4) | P.sr_code);
    h[nb++] = (uint8_t)((assignment << 4) | (P.bps_code << 1));
    const u32 v = d.frame_number;
    if (v <= 0x7F) {
        h[nb++] = (uint8_t)v;
    } else {
        const u32 tb = utf8_bytes(v);
        int shift = (int)(tb - 1) * 6;
        h[nb++] = (uint8_t)(((0xFFu << (8 - tb)) & 0xFF) | ((v >> shift) & ((1u << (7 - tb)) - 1u)));
        for (shift -= 6; shift >= 0; shift -= 6) h[nb++] = (uint8_t)(0x80 | ((v >> shift) & 0x3F));
    }
    if (bsc == 6) h[nb++] = (uint8_t)(d.nsamp - 1);
    else if (bsc == 7) { h[nb++] = (uint8_t)((d.nsamp - 1) >> 8); h[nb++] = (uint8_t)(d.nsamp - 1); }
    if (P.sr_code == 0xC) h[nb++] = (uint8_t)(P.sample_rate / 1000);
    else if (P.sr_code == 0xD) { h[nb++] = (uint8_t)(P.sample_rate >> 8); h[nb++] = (uint8_t)P.sample_rate; }
    else if (P.sr_code == 0xE) { h[nb++] = (uint8_t)((P.sample_rate / 10) >> 8); h[nb++] = (uint8_t)(P.sample_rate / 10); }
    u32 crc = 0;
    for (u32 i = 0; i < nb; i++) crc = crc8_byte(crc, h[i]);
    h[nb++] = (uint8_t)crc;
    return nb;
}

__global__ void k_frame_select(const bf_frame_desc* __restrict__ fd, u32 n_frames, bf_dev_params P,
                               const b200flac_plan* __restrict__ plans, bf_frame_choice* __restrict__ choice,
                               u32* __restrict__ frame_bytes)
{
    const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const bf_frame_desc d = fd[f];
    bf_frame_choice ch;
    for (int i = 0; i < B200FLAC_MAX_CHANNELS; i++) { ch.unit[i] = 0; ch.bitoff[i] = 0; }
    ch.side_slot = 0xFF; ch.pad = 0;
    const u32 u0 = f * P.K;
    if (P.stereo) {
        const u32 lb = plans[u0].bits, rb = plans[u0 + 1].bits, ab = plans[u0 + 2].bits, db = plans[u0 + 3].bits;
        u32 first, second, asg;
        if (P.mid_side) { // flac.c:581-629 (mid_side wins when both flags are set)
            if ((lb + rb) < min(min(lb + db, db + rb), ab + db)) { asg = 0x1; first = 0; second = 1; }
            else if (lb < min(rb, ab)) { asg = 0x8; first = 0; second = 3; ch.side_slot = 1; }
            else if (rb < ab) { asg = 0x9; first = 3; second = 1; ch.side_slot = 0; }
            else { asg = 0xA; first = 2; second = 3; ch.side_slot = 1; }
        } else if ((lb + rb) < (ab + db)) { asg = 0x1; first = 0; second = 1; } // flac.c:630-641
        else { asg = 0xA; first = 2; second = 3; ch.side_slot = 1; }
        ch.assignment = (uint8_t)asg; ch.n_sub = 2;
        ch.unit[0] = u0 + first; ch.unit[1] = u0 + second;
    } else {
        ch.assignment = (uint8_t)(P.channels - 1); ch.n_sub = (uint8_t)P.channels;
        for (u32 c = 0; c < P.channels; c++) ch.unit[c] = u0 + c;
    }
    // header: 4 fixed bytes + frame number + optional block size / sample rate + CRC-8
    const u32 bsc = block_size_code(d.nsamp);
    u32 hb = 4 + utf8_bytes(d.frame_number) + (bsc == 6 ? 1 : bsc == 7 ? 2 : 0) +
             (P.sr_code == 0xC ? 1 : (P.sr_code == 0xD || P.sr_code == 0xE) ? 2 : 0) + 1;
    ch.header_bytes = hb;
    u64 bits = (u64)hb * 8;
    for (u32 s = 0; s < ch.n_sub; s++) {
        ch.bitoff[s] = (u32)bits;
        bits += plans[ch.unit[s]].bits;
    }
    ch.frame_bytes = (u32)((bits + 7) >> 3) + 2;
    {
        // the header's bytes are made here, once per frame by one thread of a tiny kernel, not by thread 0 of the
        // packing CTA (where they were ~150 instructions of an instruction-fetch-bound kernel's hot code)
        uint8_t h[16];
        for (int i = 0; i < 16; i++) h[i] = 0;
        build_frame_header(h, d, P, ch.assignment);
        for (int w = 0; w < 4; w++)
            ch.header_words[w] = ((u32)h[4 * w] << 24) | ((u32)h[4 * w + 1] << 16) | ((u32)h[4 * w + 2] << 8) | (u32)h[4 * w + 3];
    }
    choice[f] = ch;
    frame_bytes[f] = ch.frame_bytes;
}

// single-CTA exclusive scan (n_frames is at most a few 10^5).  carry_in: where the frames of this range start
// in the output (the running total left by the previous chunk of the batch; NULL = 0); the new running total
// goes to *total and, when given, *total2.
__global__ void k_scan_offsets(const u32* __restrict__ frame_bytes, u32 n_frames, u64* __restrict__ frame_off,
                               u64* __restrict__ total, const u64* __restrict__ carry_in = nullptr,
                               u64* __restrict__ total2 = nullptr)
{
    __shared__ u64 red[40];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = carry_in ? *carry_in : 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (u32 base = 0; base < n_frames; base += blockDim.x) {
        const u32 i = base + threadIdx.x;
        const u64 v = i < n_frames ? (u64)frame_bytes[i] : 0ull;
        u64 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) red[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            u64 w = lane < nw ? red[lane] : 0ull, winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u64 t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
                if (lane >= o) winc += t;
            }
            red[lane] = winc - w;
            if (lane == 31) red[32] = winc;
        }
        __syncthreads();
        if (i < n_frames) frame_off[i] = carry + red[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += red[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) { *total = carry; if (total2) *total2 = carry; }
}

__global__ void k_zero_output(uint4* __restrict__ out, const u64* __restrict__ total, u64 capacity_bytes)
{
    u64 nbytes = *total + 16;
    if (nbytes > capacity_bytes) nbytes = capacity_bytes;
    const u64 n16 = (nbytes + 15) >> 4;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (u64)gridDim.x * blockDim.x) out[i] = z;
}

// ---------------------------------------------------------------------------
// subframe header of flac.c:820-826 / 897-905 / 977-985
__device__ __forceinline__ void put_subframe_header(BitSink& bs, u32 type_bits, u32 wasted)
{
    bs.put(type_bits & 0x3F, 7);       // pad bit 0 + 6 type bits
    if (wasted) { bs.put(1, 1); bs.zeros(wasted - 1); bs.put(1, 1); }
    else bs.put(0, 1);
}

// frame header, flac.c:488-517, with its CRC-8
template <class Sink>
__device__ void put_frame_header(Sink& bs, const bf_frame_desc& d, const bf_dev_params& P, u32 assignment)
{
    uint8_t h[16];
    const u32 nb = build_frame_header(h, d, P, assignment);
    for (u32 i = 0; i < nb; i++) bs.put(h[i], 8);
}

template <int S>
__global__ void k_pack_subframes(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd,
                                 bf_dev_params P, const b200flac_plan* __restrict__ plans,
                                 const uint8_t* __restrict__ rice, const bf_frame_choice* __restrict__ choice,
                                 const u64* __restrict__ frame_off, u32* __restrict__ out_words,
                                 int* __restrict__ g_samples, const u64* __restrict__ total, u64 capacity_bytes)
{
    if (*total + 16 > capacity_bytes) return; // batch does not fit: the host reports the error
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

    AnalyzeCtx c;
    c.n = (int)n; c.red = red; c.resid = nullptr; c.psum = nullptr; c.karr = nullptr; c.lvl_total = nullptr; c.sc = nullptr;
    if (P.samples_in_smem) c.samp = (int*)dyn_smem;
    else c.samp = g_samples + (size_t)blockIdx.x * 2 * P.samp_stride; // at most n_frames*channels <= n_units blocks

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

    // samples, shifted by the wasted bits
    u32 orv, differ;
    load_unit_samples<S>(c, pcm, d.pcm_off, cand, P, &orv, &differ);
    if (wasted) {
        for (u32 base = tid * S; base < n; base += nt * S) {
            const u32 hi = min(base + S, n);
            for (u32 i = base; i < hi; i++) c.samp[PADI(i)] >>= wasted;
        }
    }
    __syncthreads();

    if (plan.type == BF_VERBATIM) {
        const u64 body = start + 8 + wasted;
        if (tid == 0) {
            BitSink bs; bs.init(out_words, start);
            put_subframe_header(bs, 1, wasted);
            bs.flush();
        }
        for (u32 base = tid * S; base < n; base += nt * S) {
            const u32 hi = min(base + S, n);
            BitSink bs; bs.init(out_words, body + (u64)base * sub_bps);
            for (u32 i = base; i < hi; i++) bs.put_signed(c.samp[PADI(i)], sub_bps);
            bs.flush();
        }
        return;
    }

    // FIXED / LPC
    const u32 order = plan.order;
    const u32 po = plan.partition_order, under = plan.flags & 1u;
    const u32 plen = n >> po;
    const u32 kbits = plan.coding_method ? 5u : 4u;
    u64 hdr_end; // bit position where the residual block starts
    if (plan.type == BF_FIXED) hdr_end = start + 8 + wasted + (u64)order * sub_bps;
    else hdr_end = start + 8 + wasted + (u64)order * sub_bps + 4 + 5 + (u64)order * plan.precision;
    if (plan.type == BF_LPC && tid < (int)order) s_q[tid] = plan.coeffs[tid];
    __syncthreads();

    if (tid == 0) {
        BitSink bs; bs.init(out_words, start);
        if (plan.type == BF_FIXED) put_subframe_header(bs, 0x8 | order, wasted);
        else put_subframe_header(bs, 0x20 | (order - 1), wasted);
        for (u32 i = 0; i < order; i++) bs.put_signed(c.samp[PADI(i)], sub_bps); // warm-up
        if (plan.type == BF_LPC) {
            bs.put(plan.precision - 1, 4);
            bs.put_signed(plan.shift, 5);
            for (u32 i = 0; i < order; i++) bs.put_signed(plan.coeffs[i], plan.precision);
        }
        bs.put(plan.coding_method, 2);
        bs.put(po, 4);
        bs.flush();
    }
    const u64 res0 = hdr_end + 6;

    // The residual block is laid out as in flac.c:1409-1434: each partition's Rice parameter,
    // then its residuals.  Partition p starts at sample st(p) = max(p*plen, order) (only partition 0
    // when `under`); its parameter field precedes the first residual with index >= st(p).
    // Per pass of blockDim*S samples: (1) bits of each thread's run, (2) block scan, (3) emit.
    u64 run_pos = res0;
    for (u32 pbase = 0; pbase < n; pbase += nt * S) {
        const u32 base = pbase + tid * S;
        const u32 lo = max(base, order), hi = min(base + S, n);
        int r_local[S];
        u32 mybits = 0;
        if (lo < hi) {
            u32 p = under ? 0u : lo / plen;
            u32 next = under ? 0xFFFFFFFFu : (p + 1) * plen;
            u32 k = krice[p];
            // parameter fields owned by this run: partitions whose start falls at its first sample
            if (lo == order) mybits += kbits * (under ? 1u : (lo / plen + 1u)); // 0..p, empty leading one too
            else if (!under && lo == p * plen) mybits += kbits;
#pragma unroll
            for (int j = 0; j < S; j++) {
                const u32 i = lo + j;
                if (i < hi) {
                    if (i == next) { p++; next += plen; k = krice[p]; mybits += kbits; }
                    const int r = (plan.type == BF_FIXED) ? fixed_residual(c.samp, i, order)
                                                          : lpc_residual(c.samp, i, order, s_q, plan.shift);
                    r_local[j] = r;
                    mybits += (zigzag(r) >> k) + 1u + k;
                }
            }
        }
        u32 total;
        const u32 off = block_exscan_u32(mybits, red, &total);
        if (lo < hi) {
            BitSink bs; bs.init(out_words, run_pos + off);
            u32 p = under ? 0u : lo / plen;
            u32 next = under ? 0xFFFFFFFFu : (p + 1) * plen;
            u32 k = krice[p];
            if (lo == order) {
                const u32 lead = under ? 1u : (lo / plen + 1u);
                for (u32 q = 0; q < lead; q++) bs.put(krice[q], kbits);
            } else if (!under && lo == p * plen) {
                bs.put(k, kbits);
            }
#pragma unroll
            for (int j = 0; j < S; j++) {
                const u32 i = lo + j;
                if (i < hi) {
                    if (i == next) { p++; next += plen; k = krice[p]; bs.put(k, kbits); }
                    const u32 u = zigzag(r_local[j]);
                    bs.zeros(u >> k);
                    bs.put((1u << k) | (u & ((1u << k) - 1u)), k + 1);
                }
            }
            bs.flush();
        }
        run_pos += total;
        __syncthreads();
    }
    const u64 total_end = run_pos;
    // parameter fields of partitions that start at n (no residual follows): only when the block has
    // no residuals at all (n == order) or for the empty partitions of an underflow level
    if (tid == 0) {
        u32 first_trailing;
        if (n == order) first_trailing = 0;
        else if (under) first_trailing = 1;
        else first_trailing = 1u << po; // none
        if (first_trailing < (1u << po)) {
            BitSink bs; bs.init(out_words, total_end);
            for (u32 p = first_trailing; p < (1u << po); p++) bs.put(krice[p], kbits);
            bs.flush();
        }
    }
}

// one warp per frame: CRC-16 over every byte before it, appended big-endian.
// Byte-table CRC per lane over a contiguous chunk (aligned 32-bit loads), then
// crc(A||B) = crc(A) * x^(8|B|) + crc(B) to combine the 32 chunks.
__global__ void k_frame_crc16(const u64* __restrict__ frame_off, const u32* __restrict__ frame_bytes,
                              u32 n_frames, uint8_t* __restrict__ out, const u64* __restrict__ total,
                              u64 capacity_bytes)
{
    __shared__ unsigned short tab[256];
    for (u32 t = threadIdx.x; t < 256; t += blockDim.x) tab[t] = (unsigned short)crc16_byte(0, t);
    __syncthreads();
    if (*total + 16 > capacity_bytes) return;
    const u32 f = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (f >= n_frames) return;
    const uint8_t* p = out + frame_off[f];
    const u32 nb = frame_bytes[f] - 2;
    const u32 per = ((nb + 31) / 32 + 3) & ~3u;
    const u32 b0 = min(lane * per, nb), b1 = min(b0 + per, nb);
    u32 crc = 0;
    u32 i = b0;
    // head bytes up to 4-byte alignment, then whole words, then the tail
    while (i < b1 && ((uintptr_t)(p + i) & 3)) { crc = ((crc << 8) & 0xFFFF) ^ tab[(crc >> 8) ^ p[i]]; i++; }
    for (; i + 4 <= b1; i += 4) {
        const u32 w = *(const u32*)(p + i);
        crc = ((crc << 8) & 0xFFFF) ^ tab[(crc >> 8) ^ (w & 0xFF)];
        crc = ((crc << 8) & 0xFFFF) ^ tab[(crc >> 8) ^ ((w >> 8) & 0xFF)];
        crc = ((crc << 8) & 0xFFFF) ^ tab[(crc >> 8) ^ ((w >> 16) & 0xFF)];
        crc = ((crc << 8) & 0xFFFF) ^ tab[(crc >> 8) ^ (w >> 24)];
    }
    for (; i < b1; i++) crc = ((crc << 8) & 0xFFFF) ^ tab[(crc >> 8) ^ p[i]];
    crc = gf16_mul(crc, gf16_xpow8(nb - b1));
#pragma unroll
    for (int o = 16; o; o >>= 1) crc ^= __shfl_xor_sync(0xFFFFFFFFu, crc, o);
    if (lane == 0) {
        uint8_t* q = out + frame_off[f] + nb;
        q[0] = (uint8_t)(crc >> 8);
        q[1] = (uint8_t)crc;
    }
}
