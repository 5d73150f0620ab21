// k_lpc_model.cuh -- the floating-point stage of the encoder, one THREAD per
// unit (frame, candidate) and lag group:
//
//   k_lpc_autoc   window apply          flacenc_window_signal           flac.c:1129-1167
//                 autocorrelation       flacenc_autocorrelate           flac.c:1169-1188
//   k_lpc_finish  Levinson-Durbin       flacenc_compute_lp_coefficients flac.c:1190-1231
//                 order estimate        flacenc_estimate_best_lpc_order flac.c:1233-1268
//                 quantisation          flacenc_quantize_coefficients   flac.c:1270-1324
//
// Why one thread per chain: the reference sums each autocorrelation lag
// sequentially in double precision, one rounding per multiply and one per add.
// Any tree or warp-shuffle reduction changes the rounding and, now and then,
// a quantised coefficient -- and with it the file bytes.  A thread that walks
// its block in the reference's order with IEEE mul/add (no FMA contraction)
// reproduces every bit; the parallelism comes from the ~10^5 independent units
// of a batch, and, for small batches, from splitting a unit's independent lag
// sums over G threads.  The lag sums a thread owns are independent dependency
// chains, which hides the FP64 latency.
//
// k_lpc_autoc is a persistent kernel of one-warp CTAs.  A TASK is (a run of up to
// 32/K consecutive frames OF THE SAME LENGTH -- the host cuts the runs, bf_lpc_task --
// with all their candidates, lag group); warps draw tasks from a global ticket counter, so the last
// tasks of a batch land on whichever SMs are free (a static grid's last wave
// was packed onto a few SMs by the block scheduler and ran at a third of the
// speed).  A task's PCM rows (the frames its units belong to) are staged through
// shared memory in tiles, double buffered, together with the matching tile of the
// Tukey window: one bulk asynchronous copy (cp.async.bulk, the TMA engine) per row
// and tile, completing on the buffer's mbarrier -- 9 copy instructions per tile
// where per-lane cp.async took ~20 issue slots per lane; every lane then walks its own row.
#pragma once
#include "flac_common.cuh"

#ifndef LPC_MIN_CTAS
#define LPC_MIN_CTAS 1
#endif
#define LPC_MAX_ROWS 32              // frames of a task (K = 1: 32)

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
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- bulk asynchronous copies (TMA engine, UBLKCP in SASS) completing on a shared-memory mbarrier ----
// One instruction moves a whole staged row: 16-byte aligned source and destination, size a multiple of 16.
__device__ __forceinline__ void mbar_init(u64* bar, u32 count)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, u32 bytes, u64* bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}

// staging area of a warp (dynamic shared memory):
//   hdr | window tiles [2][WT] (double) | PCM tiles [2][rows][row_words] (u32)
struct LpcStageHdr {
    u64 row_byte0[LPC_MAX_ROWS];   // byte offset of the row's first PCM frame
    u32 row_n[LPC_MAX_ROWS];       // PCM frames in the row (0: no such frame)
    u64 bar[2];                    // mbarriers of the two tile buffers (bulk staging)
};

// tile geometry, shared by host and device: T = PCM frames per staged tile (>= the longest ring),
// rows = frames a task can span, row_words = u32 words per staged row.  Bulk staging (rows <= 8, i.e. K >= 4):
// a row is one bulk copy from the 16-byte aligned address at or below its first byte, so rows are 16-byte
// aligned and 4 x odd words apart (the <= 8 rows of a task then start in different banks); otherwise rows are
// an odd number of words apart and filled with 4-byte cp.async.
struct LpcGeom { u32 T, WT, rows, row_words, smem_bytes, bulk; };   // WT: doubles per window tile (even: 16-byte aligned tiles)
__host__ __device__ inline LpcGeom lpc_geometry(u32 K, u32 rowbytes, u32 ring)
{
    LpcGeom g;
    g.rows = 32 / K;
    g.bulk = g.rows <= 8 ? 1u : 0u;
    u32 T = 4608 / (g.rows * rowbytes);           // ~4.5 KB per PCM tile: 20 one-warp CTAs per SM fit at any sample width
    if (T > 160) T = 160;
    if (T < ring) T = ring;
    g.T = T;
    g.WT = (T + 3) & ~1u;
    if (g.bulk) {
        u32 w = (T * rowbytes + 15 + 15 + 3) >> 2;   // <= 15 bytes before the first frame, size rounded up to 16
        w = (w + 3) & ~3u;
        if (((w >> 2) & 1u) == 0) w += 4;            // 4 x odd
        g.row_words = w;
    } else {
        g.row_words = (((T * rowbytes + 6) >> 2) + 1) | 1u;
    }
    g.smem_bytes = (u32)sizeof(LpcStageHdr) + 2 * g.WT * 8 + 2 * g.rows * g.row_words * 4 + 16;
    return g;
}

// candidate signal of a 16-bit stereo pair (flacenc_average_difference, flac.c:1507-1529) as ONE
// dot product: (cl*L + cr*R) >> sh with (cl, cr, sh) = (1,0,0) left, (0,1,0) right, (1,1,1) average,
// (1,-1,0) difference; coef packs cl and cr as two signed bytes (IDP.2A.LO.S16.S8).  One code path
// for the four candidates: lanes of a warp hold different candidates.
__device__ __forceinline__ int stereo16_candidate(u32 pr, int coef, int sh)
{
    return __dp2a_lo((int)pr, coef, 0) >> sh;
}
__device__ __forceinline__ int stereo16_coef(u32 cand) { return cand == 0 ? 0x0001 : cand == 1 ? 0x0100 : cand == 2 ? 0x0101 : 0xFF01; }

// One task: lags LB .. LB+NL-1 of the 32 units u0 .. u0+31 (one unit per lane).
// hist is a ring of H = LB+NL windowed samples with compile-time indexing; tiles of TS = H*M
// PCM frames per row are staged with cp.async, double buffered, so the copy of tile t+1 is in
// flight while tile t is consumed.
// FAST: 16-bit stereo (one 32-bit load and a dot product per sample); otherwise samples are
// assembled from the staged bytes.
template <int LB, int NL, bool FAST, bool BULK>
__device__ __forceinline__ void autoc_task(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd,
                                           const double* __restrict__ windows, const bf_dev_params& P,
                                           const LpcGeom& g, u32 f0, u32 nrows, unsigned char* smem,
                                           double* __restrict__ autoc_out, u32 autoc_stride, uint8_t* __restrict__ wasted_out,
                                           u32 (&phase)[2])
{
    constexpr int H = LB + NL;
    const int lane = threadIdx.x & 31;
    LpcStageHdr* hdr = (LpcStageHdr*)smem;
    double* wbuf = (double*)(smem + sizeof(LpcStageHdr));     // two tiles of WT doubles
    u32* tbuf = (u32*)(wbuf + 2 * g.WT);
    const u32 K = P.K, C = P.channels, B = P.bytes_ps;
    const u32 rowbytes = C * B;
    const u32 row_words = g.row_words;
    const u32 M = g.T / H;
    const u32 TS = H * M;

    // lane -> (row, candidate); every frame of the task has the same length n, hence one window
    const bool valid = (u32)lane < nrows * K;
    const u32 myrow = valid ? (u32)lane / K : 0;
    const u32 cand = valid ? (u32)lane % K : 0;
    const u32 frame = f0 + myrow;
    const u32 unit = frame * K + cand;
    const bf_frame_desc d0 = fd[f0];
    const u32 n = d0.nsamp, nmax = n;
    const u32 woff0 = d0.window_off;

    __syncwarp();                           // the previous task is done with the staging area
    for (u32 r = lane; r < nrows; r += 32) {
        const bf_frame_desc rd = fd[f0 + r];
        hdr->row_byte0[r] = rd.pcm_off * (u64)rowbytes;
        hdr->row_n[r] = rd.nsamp;
    }
    __syncwarp();

    double hist[H];
    double a_[NL];          // private accumulators: stay in registers
#pragma unroll
    for (int i = 0; i < H; i++) hist[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NL; i++) a_[i] = 0.0;
    u32 orv = 0;
    const int coef = stereo16_coef(cand), sh = cand == 2 ? 1 : 0;
    const int cl24 = cand == 1 ? 0 : 1, cr24 = cand == 0 ? 0 : cand == 3 ? -1 : 1;     // (L, R) weights of the candidate

    auto issue_tile = [&](u32 i0, u32 b) {
        u32* base = tbuf + (size_t)b * g.rows * row_words;
        if (BULK) {
            // lane r < nrows: one bulk copy of row r's tile, from the aligned address at or below its first byte;
            // lane 31: the window tile.  Lane 0 arms the buffer's mbarrier with the byte total first.
            const void* src = nullptr;
            void* dst = nullptr;
            u32 bytes = 0;
            if ((u32)lane < nrows) {
                const u32 rn = hdr->row_n[lane];
                if (i0 < rn) {
                    const u32 take = min(TS, rn - i0);
                    const uintptr_t a = (uintptr_t)pcm + hdr->row_byte0[lane] + (u64)i0 * rowbytes;
                    const u32 mis = (u32)(a & 15);
                    src = (const void*)(a - mis);
                    dst = base + (u32)lane * row_words;
                    bytes = (mis + take * rowbytes + 15) & ~15u;
                }
            } else if (lane == 31) {
                const u32 take = min(TS, nmax - i0);
                const uintptr_t a = (uintptr_t)(windows + woff0 + i0);
                const u32 mis = (u32)(a & 15);
                src = (const void*)(a - mis);
                dst = wbuf + b * g.WT;
                bytes = (mis + take * 8 + 15) & ~15u;
            }
            const u32 total = __reduce_add_sync(0xFFFFFFFFu, bytes);
            // the buffer's previous contents were read through the generic proxy; the copies write through the async one
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            if (lane == 0) mbar_expect_tx(&hdr->bar[b], total);
            __syncwarp();
            if (bytes) bulk_g2s(dst, src, bytes, &hdr->bar[b]);
        } else {
        for (u32 r = 0; r < nrows; r++) {
            const u32 rn = hdr->row_n[r];
            if (i0 < rn) {
                const u32 take = min(TS, rn - i0);
                const u64 a = hdr->row_byte0[r] + (u64)i0 * rowbytes;
                const u32 mis = (u32)(a & 3);
                const u32 nwords = (mis + take * rowbytes + 3) >> 2;
                const u32* src = (const u32*)(pcm + (a - mis));
                u32* dst = base + r * row_words;
                for (u32 w = lane; w < nwords; w += 32) cp_async4(dst + w, src + w);
            }
        }
        const u32 take = min(TS, nmax - i0);
        for (u32 w = lane; w < take; w += 32) cp_async8(wbuf + b * g.WT + w, windows + woff0 + i0 + w);
        cp_async_commit();
        }
    };
    // wait until tile buffer b holds its data; `more`: another tile is in flight behind it
    auto wait_tile = [&](u32 b, bool more) {
        if (BULK) { mbar_wait(&hdr->bar[b], phase[b]); phase[b] ^= 1u; }
        else if (more) cp_async_wait<1>();
        else cp_async_wait<0>();
    };

    const u64 my_byte0 = hdr->row_byte0[myrow];
    // flat part of the Tukey window of length n (flac.c:1140-1152): window1 < i <= window2
    const u32 flat_lo = (u32)(0.5 * (n - 1)) / 2 + 1;
    const u32 flat_hi = (u32)((n - 1) * (1.0 - (0.5 / 2.0)));
    const u32 flat_first = flat_lo + (H - 1);   // first rotation start whose partners are flat too
    const u32 flat_last = flat_hi;              // need ig + H - 1 <= flat_last
    issue_tile(0, 0);
    u32 b = 0;
    for (u32 i0 = 0; i0 < nmax; i0 += TS, b ^= 1) {
        const bool more = i0 + TS < nmax;
        if (more) issue_tile(i0 + TS, b ^ 1);
        wait_tile(b, more);
        __syncwarp();
        const u32 amask = BULK ? 15u : 3u;
        const uint8_t* row = (const uint8_t*)(tbuf + ((size_t)b * g.rows + myrow) * row_words) +
                             (u32)(((BULK ? (uintptr_t)pcm : (uintptr_t)0) + my_byte0 + (u64)i0 * rowbytes) & amask);
        const u32 row_sa = (u32)__cvta_generic_to_shared((const void*)((uintptr_t)row & ~(uintptr_t)3));   // (24-bit stereo fetch)
        const u32 row24 = (u32)((uintptr_t)row & 3);
        const double* wt = wbuf + b * g.WT +
                           (BULK ? (u32)(((uintptr_t)(windows + woff0 + i0) & 15) >> 3) : 0u);
#pragma unroll 1
        for (u32 m = 0; m < M; m++) {
            const u32 tbase = m * H;
            const u32 ig = i0 + tbase;                 // first sample of this ring rotation
            if (ig >= nmax) break;
            // whole rotation in range?
            const bool full = ig + H <= n;
            // ... and inside the flat part of the Tukey window, partners included: samples are exact
            // integers there and their products are exact in double, so fma(x, y, acc) rounds exactly
            // like the reference's multiply-then-add
            const bool flat = full && ig >= flat_first && ig + H - 1 <= flat_last;
            // candidate sample t of this lane's row in the staged tile
            auto fetch = [&](u32 t) -> int {
                if (FAST) return stereo16_candidate(((const u32*)row)[t], coef, sh);
                if (!P.stereo) return ld_pcm(row, t * C + cand, B);
                if (B == 3) {
                    // a 24-bit stereo frame is six bytes at an even offset (4-byte aligned buffer, offsets that are
                    // multiples of six): both samples lie in two aligned words, one PRMT each (byte select + sign).
                    // The candidate is (cl * L + cr * R) >> sh like the 16-bit pair's dot product: the lanes of a
                    // warp hold different candidates, and `cand == 0 ? L : ...` compiled to four divergent branches
                    // per sample (a sixth of the kernel's instructions at 24 bits).  Explicit shared-memory loads:
                    // the staging pointer reaches this function as a generic one.
                    const u32 a = row24 + 6u * t;
                    u32 w0, w1;
                    asm("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(row_sa + (a & ~3u)));
                    asm("ld.shared.u32 %0, [%1 + 4];" : "=r"(w1) : "r"(row_sa + (a & ~3u)));
                    const int L = s24_from_words(w0, w1, (int)(a & 3u)), R = s24_from_words(w0, w1, (int)(a & 3u) + 3);
                    return (cl24 * L + cr24 * R) >> sh;
                }
                const int L = ld_pcm(row, t * 2, B), R = ld_pcm(row, t * 2 + 1, B);
                return cand == 0 ? L : cand == 1 ? R : cand == 2 ? ((L + R) >> 1) : (L - R);
            };
            // Anything but 16-bit stereo: the rotation's samples are unpacked ONCE, ahead of the two variants of the
            // loop below -- unpacking inside both put 29 KB of hot code into sixteen independent warps per SM
            // (`no_instruction` 1.5 stalls per issue at 24 bits).  The 16-bit pair's two instructions stay where they were.
            int pre[FAST ? 1 : H];
            if constexpr (!FAST) {
#pragma unroll
                for (int u = 0; u < H; u++) pre[u] = (full || (ig + u < n)) ? fetch(tbase + u) : 0;
            }
            if (flat) {
#pragma unroll
                for (int u = 0; u < H; u++) {
                    const int sv = FAST ? fetch(tbase + u) : pre[FAST ? 0 : u];
                    if (LB == 0) orv |= (u32)sv;
                    const double x = int2double_exact(sv);
                    hist[u] = x;
#pragma unroll
                    for (int l = 0; l < NL; l++) {
                        const int partner = (u - (LB + l) + 2 * H) % H;
                        a_[l] = __fma_rn(x, hist[partner], a_[l]);
                    }
                }
            } else {
                // tapered part of the window, and the block's last, partial rotation: out-of-range
                // samples count as zero, which leaves the sums untouched
                const double* wp = wt + tbase;
#pragma unroll
                for (int u = 0; u < H; u++) {
                    const bool in = full || (ig + u < n);
                    const int sv = FAST ? (in ? fetch(tbase + u) : 0) : pre[FAST ? 0 : u];
                    if (LB == 0) orv |= (u32)sv;
                    const double x = in ? __dmul_rn(int2double_exact(sv), wp[u]) : 0.0;
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
    if (valid && n > 0) {
        double* o = autoc_out + (size_t)unit * autoc_stride + LB;
#pragma unroll
        for (int i = 0; i < NL; i++) o[i] = a_[i];
        if (LB == 0) wasted_out[unit] = (uint8_t)(orv ? (u32)(__ffs((int)orv) - 1) : 0u);
    }
}

// MAXL: compile-time bound on max_lpc_order handled by this instantiation (8, 12, 16 or 32).
// G: lag groups a unit is split into (1 or 2; MAXL = 32 always uses 2 -- 33 sums and their ring do
// not fit the register file).  Large batches use G = 1 (the unpack/window work is not duplicated),
// small ones G = 2 (half the sequential work per thread, twice the warps).
// MODE: 0 = 16-bit stereo (one 32-bit load and a dot product per sample), 1 = any other shape with bulk staging,
// 2 = per-lane staging (K < 4).  One kernel per mode: each gets its own register allocation -- the unpacking of the
// general shapes costs 30 registers the 16-bit kernel must not pay in resident warps -- and a third of the code.
template <int MAXL, int G, int MODE>
__global__ void __launch_bounds__(32, LPC_MIN_CTAS)
k_lpc_autoc(const uint8_t* __restrict__ pcm, const bf_frame_desc* __restrict__ fd, u32 n_frames,
            const double* __restrict__ windows, bf_dev_params P, const bf_lpc_task* __restrict__ tasks,
            u32 n_tasks, u32* __restrict__ ticket,
            double* __restrict__ autoc_out, uint8_t* __restrict__ wasted_out)
{
    extern __shared__ __align__(16) unsigned char lpc_smem[];
    constexpr int NL0 = (G == 1) ? MAXL + 1 : (MAXL + 2) / 2, NL1 = MAXL + 1 - NL0;
    const int lane = threadIdx.x & 31;
    const LpcGeom g = lpc_geometry(P.K, P.channels * P.bytes_ps, MAXL + 1);
    u32 phase[2] = {0u, 0u};        // parity of the next completion of each tile buffer's mbarrier
    if (MODE != 2) {
        LpcStageHdr* hdr = (LpcStageHdr*)lpc_smem;
        if (lane == 0) { mbar_init(&hdr->bar[0], 1); mbar_init(&hdr->bar[1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        __syncwarp();
    }
#define AUTOC_TASK(LB_, NL_) \
    do { \
        autoc_task<LB_, NL_, MODE == 0, MODE != 2>(pcm, fd, windows, P, g, tk.first_frame, tk.n_frames, lpc_smem, autoc_out, MAXL + 1, wasted_out, phase); \
    } while (0)
    for (;;) {
        u32 t = 0;
        if (lane == 0) t = atomicAdd(ticket, 1u);
        t = __shfl_sync(0xFFFFFFFFu, t, 0);
        if (t >= n_tasks) break;
        const bf_lpc_task tk = tasks[t / G];
        if (G == 1 || (t % G) == 0) AUTOC_TASK(0, NL0);
        else if constexpr (G == 2) AUTOC_TASK(NL0, NL1);
    }
#undef AUTOC_TASK
}

// Levinson-Durbin, order estimate and quantisation: one thread per unit.
template <int MAXL>
__global__ void __launch_bounds__(128)
k_lpc_finish(const bf_frame_desc* __restrict__ fd, u32 n_frames, bf_dev_params P,
             const double* __restrict__ autoc_in, const uint8_t* __restrict__ wasted_in,
             bf_lpc_head* __restrict__ heads, short* __restrict__ coefs)
{
    const u32 unit = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 K = P.K;
    if (unit >= n_frames * K) return;
    const u32 frame = unit / K, cand = unit % K;
    const u32 n = fd[frame].nsamp;
    const u32 L = P.max_lpc_order;
    double autoc[MAXL + 1];
#pragma unroll
    for (int i = 0; i <= MAXL; i++) autoc[i] = autoc_in[(size_t)unit * (MAXL + 1) + i];
    const u32 wasted = wasted_in[unit];

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
