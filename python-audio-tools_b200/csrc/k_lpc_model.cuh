// k_lpc_model.cuh -- the floating-point stage of the encoder, one THREAD per
// unit (frame, candidate):
//
//   window apply          flacenc_window_signal           flac.c:1129-1167
//   autocorrelation       flacenc_autocorrelate           flac.c:1169-1188
//   Levinson-Durbin       flacenc_compute_lp_coefficients flac.c:1190-1231
//   order estimate        flacenc_estimate_best_lpc_order flac.c:1233-1268
//   quantisation          flacenc_quantize_coefficients   flac.c:1270-1324
//
// Why one thread per unit: the reference sums each autocorrelation lag
// sequentially in double precision, one rounding per multiply and one per add.
// Any tree or warp-shuffle reduction changes the rounding and, now and then,
// a quantised coefficient -- and with it the file bytes.  A thread that walks
// its block in the reference's order with IEEE mul/add (no FMA contraction)
// reproduces every bit; the parallelism comes from the ~10^5 independent units
// of a batch instead.  The (max_lpc_order+1) lag sums are independent
// dependency chains inside the thread, which hides the FP64 latency.
//
// A warp owns 32 consecutive units = all candidates of ~32/K consecutive
// frames.  PCM is staged through shared memory in tiles of TS PCM frames per
// frame row with coalesced 32-bit loads, then every lane walks its own row.
#pragma once
#include "flac_common.cuh"

#define LPC_WARPS 4
#define LPC_STAGE_WORDS 1344         // words per staging buffer and warp (two buffers per warp)
#define LPC_MAX_ROWS 34              // frames a warp's 32 units can span (K = 1: 32)

// (int)round(x) the way x86-64 cvttsd2si does it: NaN / out of range -> INT_MIN
// (CUDA's own conversion would give 0 / saturate)
__device__ __forceinline__ int x86_d2i(double x)
{
    if (!(x > -2147483649.0 && x < 2147483648.0)) return (int)0x80000000;
    return (int)x;
}

// flac.c:1270-1324
__device__ void quantize_coefficients(const double* c, u32 order, u32 precision, short* qlp, int* shift_out)
{
    double l = 2.2250738585072014e-308; // DBL_MIN
    for (u32 i = 0; i < order; i++) {
        const double a = fabs(c[i]);
        l = (a > l) ? a : l;
    }
    int log2cmax;
    frexp(l, &log2cmax);
    int shift = (int)(precision - 1) - (log2cmax - 1) - 1;
    shift = max(shift, -16);
    shift = min(shift, 15);
    const int qmax = (1 << (precision - 1)) - 1, qmin = -(1 << (precision - 1));
    double error = 0.0;
    if (shift >= 0) {
        const double scale = (double)(1 << shift);
        for (u32 i = 0; i < order; i++) {
            error = __dadd_rn(error, __dmul_rn(c[i], scale));
            const int ei = x86_d2i(round(error));
            qlp[i] = (short)min(max(ei, qmin), qmax);
            error = __dsub_rn(error, (double)ei);
        }
    } else {
        const double scale = (double)(1 << -shift);
        for (u32 i = 0; i < order; i++) {
            error = __dadd_rn(error, __ddiv_rn(c[i], scale));
            const int ei = x86_d2i(round(error));
            qlp[i] = (short)min(max(ei, qmin), qmax);
            error = __dsub_rn(error, (double)ei);
        }
        shift = 0;
    }
    *shift_out = shift;
}

// ---- cp.async (LDGSTS) helpers: 4-byte copies work for any sample width / alignment ----
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// per-warp staging area (dynamic shared memory): two tile buffers + the rows' descriptors
struct LpcWarpStage {
    u32 buf[2][LPC_STAGE_WORDS];
    u64 row_byte0[LPC_MAX_ROWS];   // byte offset of the row's first PCM frame
    u32 row_n[LPC_MAX_ROWS];       // PCM frames in the row (0: no such frame)
};

// One pass over the warp's units accumulating lags LB .. LB+NL-1.
// hist is a ring of H = LB+NL windowed samples with compile-time indexing; tiles of TS = H*M
// PCM frames per row are staged with cp.async, double buffered, so the copy of tile t+1 is in
// flight while tile t is consumed.
template <int LB, int NL>
__device__ __forceinline__ void autoc_pass(const uint8_t* __restrict__ pcm, const double* __restrict__ windows,
                                           const bf_dev_params& P, u32 nrows, u32 nmax,
                                           bool valid, u32 myrow, u32 cand, u32 n, u32 woff,
                                           LpcWarpStage* st, double* acc, u32* or_out)
{
    constexpr int H = LB + NL;
    const int lane = threadIdx.x & 31;
    const u32 C = P.channels, B = P.bytes_ps;
    const u32 rowbytes = C * B;
    // tile length: as many ring rotations as fit the staging buffer (at most 8)
    u32 M = ((LPC_STAGE_WORDS / nrows - 3) * 4) / (H * rowbytes);
    M = M < 1 ? 1 : (M > 8 ? 8 : M);
    const u32 TS = H * M;
    const u32 row_words = ((TS * rowbytes + 4 + 3) >> 2) + 1;

    double hist[H];
    double a_[NL];          // private accumulators: stay in registers
#pragma unroll
    for (int i = 0; i < H; i++) hist[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NL; i++) a_[i] = 0.0;
    u32 orv = 0;

    auto issue_tile = [&](u32 i0, u32 b) {
        for (u32 r = 0; r < nrows; r++) {
            const u32 rn = st->row_n[r];
            if (i0 < rn) {
                const u32 take = min(TS, rn - i0);
                const u64 a = st->row_byte0[r] + (u64)i0 * rowbytes;
                const u32 mis = (u32)(a & 3);
                const u32 nwords = (mis + take * rowbytes + 3) >> 2;
                const u32* src = (const u32*)(pcm + (a - mis));
                u32* dst = st->buf[b] + r * row_words;
                for (u32 w = lane; w < nwords; w += 32) cp_async4(dst + w, src + w);
            }
        }
        cp_async_commit();
    };

    const u64 my_byte0 = st->row_byte0[myrow];
    // per-lane constants of the fast paths
    const bool st16 = P.stereo && B == 2;
    const int cl = cand == 1 ? 0 : 1, cr = cand == 0 ? 0 : (cand == 3 ? -1 : 1), sh = cand == 2 ? 1 : 0;
    const u32 nmin = __reduce_min_sync(0xFFFFFFFFu, valid ? n : nmax);
    // flat part of the Tukey window of length n (flac.c:1140-1152): window1 < i <= window2
    const u32 flat_lo = valid ? ((u32)(0.5 * (n - 1)) / 2 + 1) : 0u;
    const u32 flat_hi = valid ? (u32)((n - 1) * (1.0 - (0.5 / 2.0))) : 0xFFFFFFFFu;
    issue_tile(0, 0);
    u32 b = 0;
    for (u32 i0 = 0; i0 < nmax; i0 += TS, b ^= 1) {
        if (i0 + TS < nmax) { issue_tile(i0 + TS, b ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
        const uint8_t* row = (const uint8_t*)(st->buf[b] + myrow * row_words) + (u32)((my_byte0 + (u64)i0 * rowbytes) & 3);
#pragma unroll 1
        for (u32 m = 0; m < M; m++) {
            const u32 tbase = m * H;
            const u32 ig = i0 + tbase;                 // first sample of this ring rotation
            if (ig >= nmax) break;
            // whole rotation in range for every lane, stereo 16-bit rows (always 4-byte aligned)?
            const bool full = st16 && (ig + H <= nmin);
            // ... and inside the flat part of every lane's Tukey window, partners included: samples are
            // exact integers there and their products are exact in double, so fma(x, y, acc) rounds
            // exactly like the reference's multiply-then-add
            const bool flat = full && __all_sync(0xFFFFFFFFu, ig >= flat_lo + (H - 1) && ig + H - 1 <= flat_hi);
            if (flat) {
                const u32* rw = (const u32*)row + tbase;
#pragma unroll
                for (int u = 0; u < H; u++) {
                    const u32 pr = rw[u];
                    const int sv = (cl * (int)(short)(pr & 0xFFFF) + cr * ((int)pr >> 16)) >> sh;
                    orv |= (u32)sv;
                    const double x = (double)sv;
                    hist[u] = x;
#pragma unroll
                    for (int l = 0; l < NL; l++) {
                        const int partner = (u - (LB + l) + 2 * H) % H;
                        a_[l] = __fma_rn(x, hist[partner], a_[l]);
                    }
                }
            } else if (full) {
                const u32* rw = (const u32*)row + tbase;
                const double* wp = windows + woff + ig;
#pragma unroll
                for (int u = 0; u < H; u++) {
                    const u32 pr = rw[u];
                    const int sv = (cl * (int)(short)(pr & 0xFFFF) + cr * ((int)pr >> 16)) >> sh;
                    orv |= (u32)sv;
                    const double x = __dmul_rn((double)sv, wp[u]);
                    hist[u] = x;
#pragma unroll
                    for (int l = 0; l < NL; l++) {
                        const int partner = (u - (LB + l) + 2 * H) % H;
                        a_[l] = __dadd_rn(a_[l], __dmul_rn(x, hist[partner]));
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < H; u++) {
                    const u32 t = tbase + u;
                    const u32 i = i0 + t;
                    double x = 0.0;
                    if (valid && i < n) {
                        int sv;
                        if (!P.stereo) {
                            sv = ld_pcm(row, t * C + cand, B);
                        } else {
                            const int L = ld_pcm(row, t * 2, B), R = ld_pcm(row, t * 2 + 1, B);
                            sv = cand == 0 ? L : cand == 1 ? R : cand == 2 ? ((L + R) >> 1) : (L - R);
                        }
                        orv |= (u32)sv;
                        x = __dmul_rn((double)sv, windows[woff + i]);
                    }
                    hist[u] = x;
#pragma unroll
                    for (int l = 0; l < NL; l++) {
                        const int partner = (u - (LB + l) + 2 * H) % H;
                        a_[l] = __dadd_rn(a_[l], __dmul_rn(x, hist[partner]));
                    }
                }
            }
        }
        __syncwarp(); // everyone is done with buf[b] before it is refilled two tiles later
    }
#pragma unroll
    for (int i = 0; i < NL; i++) acc[i] = a_[i];
    *or_out = orv;
}

// MAXL: compile-time bound on max_lpc_order handled by this instantiation (8, 12, 16 or 32)
template <int MAXL>
__global__ void __launch_bounds__(LPC_WARPS * 32)
k_lpc_model(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd, u32 n_frames,
            const double* __restrict__ windows, bf_dev_params P,
            bf_lpc_head* __restrict__ heads, short* __restrict__ coefs)
{
    extern __shared__ __align__(16) unsigned char lpc_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    LpcWarpStage* st = (LpcWarpStage*)lpc_smem + warp;

    const u32 K = P.K;
    const u32 U = n_frames * K;
    const u32 u0 = (blockIdx.x * LPC_WARPS + warp) * 32;
    if (u0 >= U) return;
    const u32 unit = u0 + lane;
    const bool valid = unit < U;
    const u32 f0 = u0 / K;
    const u32 ulast = min(u0 + 31, U - 1);
    const u32 nrows = ulast / K - f0 + 1;
    const u32 frame = valid ? unit / K : f0;
    const u32 cand = valid ? unit % K : 0;
    const bf_frame_desc d = fd[frame];
    const u32 n = valid ? d.nsamp : 0;
    const u32 nmax = __reduce_max_sync(0xFFFFFFFFu, n);
    const u32 L = P.max_lpc_order;

    // row descriptors, once
    for (u32 r = lane; r < nrows; r += 32) {
        const bf_frame_desc rd = fd[f0 + r];
        st->row_byte0[r] = rd.pcm_off * (u64)(P.channels * P.bytes_ps);
        st->row_n[r] = rd.nsamp;
    }
    __syncwarp();

    double autoc[MAXL + 1];
    u32 orv = 0;
    if constexpr (MAXL <= 16) {
        autoc_pass<0, MAXL + 1>(pcm, windows, P, nrows, nmax, valid, frame - f0, cand, n, d.window_off, st, autoc, &orv);
    } else {
        u32 dummy;
        autoc_pass<0, 17>(pcm, windows, P, nrows, nmax, valid, frame - f0, cand, n, d.window_off, st, autoc, &orv);
        __syncwarp();
        autoc_pass<17, 16>(pcm, windows, P, nrows, nmax, valid, frame - f0, cand, n, d.window_off, st, autoc + 17, &dummy);
    }
    if (!valid) return;

    bf_lpc_head head;
    head.best_order = 0; head.precision = (uint8_t)P.precision; head.dummy = 0; head.pad = 0;
    for (int i = 0; i < BF_MAX_ORDER; i++) head.shift[i] = 0;
    short* mycoef = coefs + (size_t)unit * P.model_stride;

    if (!(n > L + 1)) {
        // flac.c:1121-1126: dummy coefficients
        head.best_order = 1; head.precision = 2; head.dummy = 1; head.shift[0] = 0;
        mycoef[0] = 1;
        heads[unit] = head;
        return;
    }

    // wasted bits: the reference windows the samples AFTER shifting them right by
    // `wasted` (flac.c:700-724).  s>>w is exact, so every product and partial sum
    // is the unshifted one scaled by 2^-2w, also exactly: scale once at the end.
    const u32 wasted = orv ? (u32)(__ffs((int)orv) - 1) : 0u;
    if (wasted) {
        const double sc = ldexp(1.0, -2 * (int)wasted);
        for (u32 i = 0; i <= L; i++) autoc[i] = autoc[i] * sc;
    }

    // Levinson-Durbin, flac.c:1190-1231 (operation order preserved, no FMA)
    double lp[MAXL * (MAXL + 1) / 2];   // order o at lp[o*(o-1)/2 ..]
    double err[MAXL];
    {
        double k = __ddiv_rn(autoc[1], autoc[0]);
        lp[0] = k;
        err[0] = __dmul_rn(autoc[0], __dsub_rn(1.0, __dmul_rn(k, k)));
        for (u32 i = 1; i < L; i++) {
            const double* prev = lp + (i * (i - 1)) / 2;
            double* cur = lp + (i * (i + 1)) / 2;
            double q = autoc[i + 1];
            for (u32 j = 0; j < i; j++) q = __dsub_rn(q, __dmul_rn(prev[j], autoc[i - j]));
            k = __ddiv_rn(q, err[i - 1]);
            for (u32 j = 0; j < i; j++) cur[j] = __dsub_rn(prev[j], __dmul_rn(k, prev[i - j - 1]));
            cur[i] = k;
            err[i] = __dmul_rn(err[i - 1], __dsub_rn(1.0, __dmul_rn(k, k)));
        }
    }

    if (!P.exhaustive) {
        // flac.c:1233-1268; bps is the candidate's bps BEFORE removing wasted bits (H11)
        const u32 bps = candidate_bps(cand, P);
        const double error_scale = __ddiv_rn(0.6931471805599453 * 0.6931471805599453, __dmul_rn((double)n, 2.0));
        u32 best_order = 0;
        double best_bits = 1.7976931348623157e308;
        for (u32 i = 0; i < L; i++) {
            const u32 order = i + 1;
            if (err[i] > 0.0) {
                const u32 header_bits = order * (bps + P.precision);
                double bpr = __ddiv_rn(log(__dmul_rn(err[i], error_scale)), 0.6931471805599453 * 2);
                if (!(bpr > 0.0)) bpr = 0.0;
                const double est = __dadd_rn((double)header_bits, __dmul_rn(bpr, (double)(n - order)));
                if (est < best_bits) { best_order = order; best_bits = est; }
            } else {
                best_order = order;
                break;
            }
        }
        if (best_order == 0) best_order = 1; // unreachable for finite input (the reference asserts it)
        int shift;
        quantize_coefficients(lp + (best_order * (best_order - 1)) / 2, best_order, P.precision,
                              mycoef + (best_order * (best_order - 1)) / 2, &shift);
        head.best_order = (uint8_t)best_order;
        head.shift[best_order - 1] = (int8_t)shift;
    } else {
        for (u32 o = 1; o <= L; o++) {
            int shift;
            quantize_coefficients(lp + (o * (o - 1)) / 2, o, P.precision, mycoef + (o * (o - 1)) / 2, &shift);
            head.shift[o - 1] = (int8_t)shift;
        }
        head.best_order = (uint8_t)L;
    }
    heads[unit] = head;
}
