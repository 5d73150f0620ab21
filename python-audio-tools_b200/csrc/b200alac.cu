// b200alac.cu -- the ALAC encoder of the engine (include/b200alac.h), sm_100a.
//
// The reference encodes one frameset at a time and, inside it, tries every interlacing leftweight by encoding the
// whole frame into a recorder (src/encoders/alac.c:438-545).  Both of ALAC's inner loops are serial chains along
// a channel -- calculate_residuals nudges its predictor coefficients after every sample by the sign of the error
// (:936-1018) and encode_residuals carries an adaptive history through the block (:1034-1100) -- but every
// (frameset, channel pair, leftweight, channel, order) is a chain of its own, and an hour of stereo has 775,200 of
// them.  So the work is laid out as one THREAD per chain:
//
//   k_alac_model   one thread per (frame, leftweight, channel): correlate_channels (:678-718) fused into the
//                  sample fetch, Tukey window (:779-818), the nine autocorrelation sums walked in the reference's
//                  order with one rounding per multiply and per add (:820-838), Levinson-Durbin (:840-883),
//                  quantisation at orders 4 and 8 (:885-909).
//   k_alac_size    one thread per (frame, leftweight, channel, order): the adaptive residual and the adaptive
//                  Golomb code fused sample by sample, counting bits only (what the reference reads back from its
//                  recorders with bits_written).
//   k_alac_select  one thread per frameset: order 4 unless order 8 saves 64 bits (:752-765), the first strict
//                  minimum over the leftweights (:455-479), frame and frameset sizes; k_alac_scan: byte offsets.
//   k_alac_emit_residuals   one thread per (frame, channel) of the winning (leftweight, order): the same fused
//                  walk, now writing the code bits at their final position -- a channel's residual block is one
//                  contiguous run of the frame (:666-675), so it needs no prefix sum, only the sizes already known.
//   k_alac_emit_fixed       one CTA per frame: frame header, subframe headers, the uncompressed low bytes of
//                  24-bit samples, uncompressed frames (:405-436), the '111' that ends a frameset (:370-371).
//
// Same bytes as the reference: tests/test_alac_gpu.py compares with the tests' CPU checker, which is pinned to the
// compiled reference encoder.  No CPU fallback.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <vector>

#include "../../include/b200alac.h"

typedef unsigned long long u64;
typedef unsigned int u32;

static thread_local char a_err[512] = "";
extern "C" const char* b200alac_last_error(void) { return a_err; }
extern "C" void b200alac_free(void* p) { free(p); }
static int afail(const char* msg) { snprintf(a_err, sizeof(a_err), "%s", msg); return 1; }
#define ACK(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            snprintf(a_err, sizeof(a_err), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
            rc = 1;                                                                                      \
            goto done;                                                                                   \
        }                                                                                                \
    } while (0)

#define ALAC_SHIFT 2            // INTERLACING_SHIFT, alac.c:26
#define ALAC_MAX_LW 8           // leftweight candidates handled (the reference default is 0..4)

struct AFrameset {
    u64 pcm_off;     // first PCM frame
    u32 n;           // PCM frames
    u32 win_off;     // offset of this length's Tukey window (doubles)
};

struct AGroup { unsigned char nch, c0, c1, pad; };

struct ADev {
    u32 C, B, bps, lsb_bits, block_size, init_hist, hist_mult, max_k, lw_min, nlw, ng, upf;   // upf: units per frame = 2 * nlw
    AGroup grp[8];
};

struct AModel {      // per unit (frame, leftweight, channel)
    short q4[4];
    short q8[8];
    u32 zero;        // autocorrelation[0] == 0: coefficients 0,0,0,0 (alac.c:767-776)
    u32 pad;
};

struct AChoice {     // per frame (frameset, channel group)
    u32 compressed, lw, order[2];
    u32 res_bits[2];
    u32 bits;        // the frame's size without its 3-bit channel count
    u32 bit_off;     // bit offset of the channel count inside the frameset
};

// ---- sample access: de-interleave, strip the low byte of 24-bit samples, correlate ----
__device__ __forceinline__ int alac_ld(const uint8_t* __restrict__ p, u32 B)
{
    if (B == 2) return (int)(*(const short*)p);
    return ((int)((u32)p[0] << 8 | (u32)p[1] << 16 | (u32)p[2] << 24)) >> 8;
}

struct ASrc {
    const uint8_t* base;   // first PCM frame of the frameset
    u32 stride, o0, o1, B, lsb_bits, nch, lw, ch;
    __device__ __forceinline__ int at(u32 i) const
    {
        const uint8_t* p = base + (size_t)i * stride;
        const int a = alac_ld(p + o0, B) >> lsb_bits;
        if (nch == 1) return a;
        const int b = alac_ld(p + o1, B) >> lsb_bits;
        if (lw == 0) return ch ? b : a;
        // correlate_channels, alac.c:703-712
        return ch ? a - b : b + (int)(((long long)(a - b) * (long long)lw) >> ALAC_SHIFT);
    }
};

__device__ __forceinline__ ASrc alac_src(const uint8_t* pcm, const AFrameset& fs, const ADev& D, u32 g, u32 lw, u32 ch)
{
    ASrc s;
    s.base = pcm + fs.pcm_off * D.C * D.B;
    s.stride = D.C * D.B;
    s.o0 = D.grp[g].c0 * D.B;
    s.o1 = D.grp[g].c1 * D.B;
    s.B = D.B; s.lsb_bits = D.lsb_bits; s.nch = D.grp[g].nch; s.lw = lw; s.ch = ch;
    return s;
}

// (int)round(x) the way x86-64 does it (cvttsd2si: NaN / out of range -> INT_MIN)
__device__ __forceinline__ int alac_d2i(double x)
{
    if (!(x > -2147483649.0 && x < 2147483648.0)) return (int)0x80000000;
    return (int)x;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
k_alac_model(const uint8_t* __restrict__ pcm, const AFrameset* __restrict__ fsets, u32 n_units, ADev D,
             const double* __restrict__ windows, AModel* __restrict__ models)
{
    const u32 unit = blockIdx.x * blockDim.x + threadIdx.x;
    if (unit >= n_units) return;
    const u32 fg = unit / D.upf, r = unit % D.upf, lwi = r >> 1, ch = r & 1;
    const u32 f = fg / D.ng, g = fg % D.ng;
    if (D.grp[g].nch == 1 && r != 0) return;
    const AFrameset fs = fsets[f];
    const u32 n = fs.n;
    if (n < 10) return;                                    // uncompressed (alac.c:387)
    const ASrc src = alac_src(pcm, fs, D, g, D.lw_min + lwi, ch);
    const double* win = windows + fs.win_off;
    // window_signal + autocorrelate, alac.c:779-838: every lag's sum in the reference's order
    double h[9], acc[9];
#pragma unroll
    for (int l = 0; l < 9; l++) { h[l] = 0.0; acc[l] = 0.0; }
    for (u32 i = 0; i < n; i++) {
        const double x = __dmul_rn((double)src.at(i), win[i]);
#pragma unroll
        for (int l = 8; l > 0; l--) h[l] = h[l - 1];
        h[0] = x;
#pragma unroll
        for (int l = 0; l < 9; l++) if ((u32)l <= i) acc[l] = __dadd_rn(acc[l], __dmul_rn(h[l], x));
    }
    AModel m;
    memset(&m, 0, sizeof(m));
    if (acc[0] != 0.0) {
        // compute_lp_coefficients, alac.c:840-883
        double lp[8][8], err[8];
        double k = acc[1] / acc[0];
        lp[0][0] = k;
        err[0] = acc[0] * (1.0 - (k * k));
        for (int i = 1; i < 8; i++) {
            double q = acc[i + 1];
            for (int j = 0; j < i; j++) q -= lp[i - 1][j] * acc[i - j];
            k = q / err[i - 1];
            for (int j = 0; j < i; j++) lp[i][j] = lp[i - 1][j] - (k * lp[i - 1][i - j - 1]);
            lp[i][i] = k;
            err[i] = err[i - 1] * (1.0 - (k * k));
        }
        // quantize_coefficients, alac.c:885-909
        for (int pass = 0; pass < 2; pass++) {
            const int order = pass ? 8 : 4;
            double error = 0.0;
            for (int i = 0; i < order; i++) {
                error += lp[order - 1][i] * 512.0;
                const int ei = alac_d2i(round(error));
                const int q = min(max(ei, -32768), 32767);
                if (pass) m.q8[i] = (short)q; else m.q4[i] = (short)q;
                error -= (double)ei;
            }
        }
    } else {
        m.zero = 1;
    }
    models[unit] = m;
}

// ---- the fused residual + Golomb walk of one channel ----
struct ACount {      // sink that only counts
    u64 bits;
    __device__ __forceinline__ void put(u32, u32 n) { bits += n; }
};

struct AWrite {      // MSB-first writer into zeroed big-endian words of global memory (OR-merged)
    u32* words;
    u64 widx;
    u64 acc;         // pending bits, left-aligned
    u32 fill;
    __device__ __forceinline__ void init(u32* base, u64 bitpos) { words = base; widx = bitpos >> 5; fill = (u32)(bitpos & 31); acc = 0; }
    __device__ __forceinline__ void put(u32 v, u32 nbits)      // nbits <= 32
    {
        if (!nbits) return;
        acc |= (u64)v << (64 - fill - nbits);
        fill += nbits;
        if (fill >= 32) {
            const u32 w = (u32)(acc >> 32);
            if (w) atomicOr(words + widx, __byte_perm(w, 0, 0x0123));
            widx++; acc <<= 32; fill -= 32;
        }
    }
    __device__ __forceinline__ void finish() { if (fill) { const u32 w = (u32)(acc >> 32); if (w) atomicOr(words + widx, __byte_perm(w, 0, 0x0123)); } }
};

__device__ __forceinline__ u32 alac_log2(u32 v) { return v ? 31u - (u32)__clz((int)v) : 0xFFFFFFFFu; }   // LOG2, alac.c:1020-1031

// value / (2^k - 1) and the remainder for the values ALAC codes (below 2^25): a float quotient, corrected by one
__device__ __forceinline__ void alac_divmod(u32 value, u32 d, u32& q, u32& r)
{
    q = (u32)(__fdividef((float)value, (float)d));
    int rem = (int)value - (int)(q * d);
    if (rem < 0) { q--; rem += (int)d; }
    else if (rem >= (int)d) { q++; rem -= (int)d; }
    r = (u32)rem;
}

// write_residual, alac.c:1102-1122
__device__ __forceinline__ void alac_write_residual(ACount& s, u32 value, u32 k, u32 sample_size)
{
    // (counting only: no branches -- the lanes of a warp walk different channels)
    u32 msb, lsb;
    alac_divmod(value, (1u << k) - 1u, msb, lsb);
    const u32 tail = k > 1 ? (lsb > 0 ? k : k - 1) : 0u;
    s.bits += msb > 8 ? 9u + sample_size : msb + 1u + tail;
}

template <class Sink>
__device__ __forceinline__ void alac_write_residual(Sink& s, u32 value, u32 k, u32 sample_size)
{
    u32 msb, lsb;
    alac_divmod(value, (1u << k) - 1u, msb, lsb);
    if (msb > 8) {
        s.put(0x1FF, 9);
        s.put(value, sample_size);
    } else {
        s.put(((1u << msb) - 1u) << 1, msb + 1);          // msb ones, then a zero
        if (k > 1) {
            if (lsb > 0) s.put(lsb + 1, k);
            else s.put(0, k - 1);
        }
    }
}

// calculate_residuals (alac.c:936-1018) and encode_residuals (:1034-1100) of one channel, fused: every residual is
// coded as soon as it exists.  Returns true on the reference's residual overflow.
template <int CC, class Sink>
__device__ __forceinline__ bool alac_chain(const ASrc& src, u32 n, u32 sample_size, const short* q, const ADev& D, Sink& sink)
{
    int coef[CC];
#pragma unroll
    for (int j = 0; j < CC; j++) coef[j] = q[j];
    int w[CC + 1];               // w[t] = s[i - 1 - t]
#pragma unroll
    for (int t = 0; t <= CC; t++) w[t] = 0;
    const int smask = (1 << sample_size) - 1, sbit = 1 << (sample_size - 1);
    int history = (int)D.init_hist;
    u32 sign_modifier = 0, zeroes = 0, run_k = 0;
    bool in_run = false, overflow = false;
    const u32 max_unsigned = 1u << sample_size;
    for (u32 i = 0; i < n; i++) {
        const int s = src.at(i);
        int res;
        if (i == 0) {
            res = s;
        } else if (i < (u32)CC + 1) {
            const int t = (s - w[0]) & smask;                       // TRUNCATE_BITS, alac.c:924-934
            res = (t & sbit) ? t - (1 << sample_size) : t;
        } else {
            const int base = w[CC];
            long long sum = 1 << 8;
#pragma unroll
            for (int j = 0; j < CC; j++) sum += (long long)coef[j] * (long long)(w[j] - base);
            sum >>= 9;
            const int t = (s - base - (int)sum) & smask;
            int error = (t & sbit) ? t - (1 << sample_size) : t;
            res = error;
            // the coefficients follow the sign of the error (alac.c:986-1008); its two mirrored branches as one
            // loop over se = sign(error): coefficient -= se * sign(diff), error -= ((diff * se * sign(diff)) >> 9) * (j + 1),
            // until the error has changed sign -- lanes of a warp take different paths, so no branch here
            {
                const int se = (error > 0) - (error < 0);
                bool live = se != 0;
#pragma unroll
                for (int j = 0; j < CC; j++) {
                    const int diff = base - w[CC - 1 - j];              // base - s[i - CC + j]
                    const int sg = ((diff > 0) - (diff < 0)) * se;
                    if (live) {
                        coef[CC - j - 1] -= sg;
                        error -= ((diff * sg) >> 9) * (j + 1);
                        if (error * se <= 0) live = false;
                    }
                }
            }
        }
#pragma unroll
        for (int t = CC; t > 0; t--) w[t] = w[t - 1];
        w[0] = s;

        // ---- encode_residuals as a state machine over the residuals as they come ----
        if (in_run) {
            if (res == 0) { zeroes++; continue; }
            alac_write_residual(sink, zeroes, run_k, 16);
            if (zeroes < 0xFFFF) sign_modifier = 1;
            history = 0;
            in_run = false;
        }
        const u32 u = res >= 0 ? (u32)(res << 1) : (u32)(-res << 1) - 1u;
        if (u >= max_unsigned) overflow = true;                     // alac.c:1059-1063 (the frame goes out uncompressed)
        u32 k = alac_log2((u32)((history >> 9) + 3));
        k = min(k, D.max_k);
        alac_write_residual(sink, u - sign_modifier, k, sample_size);
        sign_modifier = 0;
        if (u <= 0xFFFF) {
            history += ((int)(u * D.hist_mult) - ((history * (int)D.hist_mult) >> 9));
            if (history < 128 && i + 1 < n) {
                in_run = true;
                zeroes = 0;
                run_k = 7u - alac_log2((u32)history) + (u32)((history + 16) >> 6);
                run_k = min(run_k, D.max_k);
            }
        } else {
            history = 0xFFFF;
        }
    }
    if (in_run) alac_write_residual(sink, zeroes, run_k, 16);
    return overflow;
}

__global__ void __launch_bounds__(64)
k_alac_size(const uint8_t* __restrict__ pcm, const AFrameset* __restrict__ fsets, u32 n_chains, ADev D,
            const AModel* __restrict__ models, u32* __restrict__ bits, u32* __restrict__ overflow)
{
    // blockIdx.y is the order (4 or 8): a warp runs ONE instantiation of the walk (with the two orders in
    // alternate lanes every warp executed both, each with half its lanes)
    const u32 unit = blockIdx.x * blockDim.x + threadIdx.x, oi = blockIdx.y;
    if (unit * 2 >= n_chains) return;
    const u32 chain = unit * 2 + oi;
    const u32 fg = unit / D.upf, r = unit % D.upf, lwi = r >> 1, ch = r & 1;
    const u32 f = fg / D.ng, g = fg % D.ng;
    if (D.grp[g].nch == 1 && r != 0) return;
    const AFrameset fs = fsets[f];
    if (fs.n < 10) return;
    const AModel m = models[unit];
    if (m.zero && oi) return;                              // all-zero input: order 4 with zero coefficients only
    const ASrc src = alac_src(pcm, fs, D, g, D.lw_min + lwi, ch);
    const u32 sample_size = D.bps - D.lsb_bits + (D.grp[g].nch == 2 ? 1u : 0u);
    ACount c; c.bits = 0;
    bool ov;
    if (oi) ov = alac_chain<8>(src, fs.n, sample_size, m.q8, D, c);
    else ov = alac_chain<4>(src, fs.n, sample_size, m.q4, D, c);
    bits[chain] = c.bits > 0xFFFFFFFFull ? 0xFFFFFFFFu : (u32)c.bits;
    if (ov) atomicOr(overflow + fg, 1u);
}

__device__ __forceinline__ u32 alac_head_bits(u32 n, u32 block_size) { return 16 + 1 + 2 + 1 + (n != block_size ? 32u : 0u); }

// one thread per frameset: the choices of its frames and its size in bytes
__global__ void k_alac_select(const AFrameset* __restrict__ fsets, u32 n_fsets, ADev D, const AModel* __restrict__ models,
                              const u32* __restrict__ bits, const u32* __restrict__ overflow, AChoice* __restrict__ choice,
                              u32* __restrict__ fset_bytes)
{
    const u32 f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_fsets) return;
    const u32 n = fsets[f].n;
    u64 pos = 0;
    for (u32 g = 0; g < D.ng; g++) {
        const u32 fg = f * D.ng + g, nch = D.grp[g].nch;
        AChoice c;
        memset(&c, 0, sizeof(c));
        c.bit_off = (u32)pos;
        const u32 head = alac_head_bits(n, D.block_size);
        if (n >= 10 && !overflow[fg]) {
            c.compressed = 1;
            u64 best = ~0ull;
            const u32 tries = nch == 2 ? D.nlw : 1u;
            for (u32 lwi = 0; lwi < tries; lwi++) {
                u64 total = head + 16 + (u64)D.lsb_bits * n * nch;
                u32 ord[2] = {0, 0}, rb[2] = {0, 0};
                for (u32 ch = 0; ch < nch; ch++) {
                    const u32 unit = fg * D.upf + lwi * 2 + ch;
                    const u32 b4 = bits[unit * 2], b8 = bits[unit * 2 + 1];
                    // alac.c:752-765: order 4 unless order 8 is more than 64 bits shorter; all-zero input: order 4
                    const bool four = models[unit].zero || ((u64)b4 < (u64)b8 + 64);
                    ord[ch] = four ? 4 : 8;
                    rb[ch] = four ? b4 : b8;
                    total += 16 + 16 * ord[ch] + rb[ch];
                }
                if (total < best) {        // first strict minimum, alac.c:470-476
                    best = total;
                    c.lw = D.lw_min + lwi; c.order[0] = ord[0]; c.order[1] = ord[1]; c.res_bits[0] = rb[0]; c.res_bits[1] = rb[1];
                }
            }
            if (nch == 1) c.lw = 0;
            c.bits = (u32)best;
        } else {
            c.compressed = 0;                  // write_uncompressed_frame, alac.c:405-436
            c.bits = head + D.bps * n * nch;
        }
        choice[fg] = c;
        pos += 3 + c.bits;
    }
    pos += 3;                                   // the frameset's closing '111' (alac.c:370-371)
    fset_bytes[f] = (u32)((pos + 7) >> 3);
}

__global__ void __launch_bounds__(1024)
k_alac_scan(const u32* __restrict__ sizes, u32 n, u64* __restrict__ off, u64* __restrict__ total)
{
    __shared__ u64 red[33];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (u32 base = 0; base < n; base += blockDim.x) {
        const u32 i = base + threadIdx.x;
        const u64 v = i < n ? (u64)sizes[i] : 0ull;
        u64 inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= (u32)o) inc += t;
        }
        if (lane == 31) red[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            u64 w = red[lane], winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u64 t = __shfl_up_sync(0xFFFFFFFFu, winc, o);
                if (lane >= (u32)o) winc += t;
            }
            red[lane] = winc - w;
            if (lane == 31) red[32] = winc;
        }
        __syncthreads();
        if (i < n) off[i] = carry + red[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += red[32];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void k_alac_zero(uint4* __restrict__ out, const u64* __restrict__ total, u64 capacity)
{
    u64 nbytes = *total + 16;
    if (nbytes > capacity) nbytes = capacity;
    const u64 n16 = (nbytes + 15) >> 4;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (u64)gridDim.x * blockDim.x)
        out[i] = make_uint4(0, 0, 0, 0);
}

// where a compressed frame's pieces start, relative to the frame's channel count
struct ALayout { u64 head, sub0, sub1, lsb, res0, res1; };
__device__ __forceinline__ ALayout alac_layout(const AChoice& c, u32 n, u32 nch, const ADev& D)
{
    ALayout L;
    L.head = 3;
    L.sub0 = L.head + alac_head_bits(n, D.block_size) + 16;
    L.sub1 = L.sub0 + 16 + 16 * c.order[0];
    L.lsb = nch == 2 ? L.sub1 + 16 + 16 * c.order[1] : L.sub1;
    L.res0 = L.lsb + (u64)D.lsb_bits * n * nch;
    L.res1 = L.res0 + c.res_bits[0];
    return L;
}

// one thread per (frame, channel): the residual block of the chosen leftweight and order, written in place
__global__ void __launch_bounds__(64)
k_alac_emit_residuals(const uint8_t* __restrict__ pcm, const AFrameset* __restrict__ fsets, u32 n_fg, ADev D,
                      const AModel* __restrict__ models, const AChoice* __restrict__ choice, const u64* __restrict__ fset_off,
                      uint8_t* __restrict__ out, const u64* __restrict__ total, u64 capacity)
{
    if (*total + 16 > capacity) return;
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_fg * 2) return;
    const u32 fg = t >> 1, ch = t & 1;
    const u32 f = fg / D.ng, g = fg % D.ng, nch = D.grp[g].nch;
    if (ch >= nch) return;
    const AChoice c = choice[fg];
    if (!c.compressed) return;
    const AFrameset fs = fsets[f];
    const u32 lwi = nch == 2 ? c.lw - D.lw_min : 0u;
    const AModel m = models[fg * D.upf + lwi * 2 + ch];
    const ASrc src = alac_src(pcm, fs, D, g, c.lw, ch);
    const u32 sample_size = D.bps - D.lsb_bits + (nch == 2 ? 1u : 0u);
    const ALayout L = alac_layout(c, fs.n, nch, D);
    AWrite w;
    w.init((u32*)out, fset_off[f] * 8 + c.bit_off + (ch ? L.res1 : L.res0));
    if (c.order[ch] == 8) alac_chain<8>(src, fs.n, sample_size, m.q8, D, w);
    else alac_chain<4>(src, fs.n, sample_size, m.q4, D, w);
    w.finish();
}

// a field of 1..32 bits at an absolute bit position of the zeroed output
__device__ __forceinline__ void alac_field(u32* words, u64 bitpos, u32 v, u32 nbits)
{
    const u64 wi = bitpos >> 5;
    const u32 o = (u32)(bitpos & 31);
    const u64 x = (u64)(nbits >= 32 ? v : (v & ((1u << nbits) - 1u))) << (64 - o - nbits);
    if ((u32)(x >> 32)) atomicOr(words + wi, __byte_perm((u32)(x >> 32), 0, 0x0123));
    if ((u32)x) atomicOr(words + wi + 1, __byte_perm((u32)x, 0, 0x0123));
}

// one CTA per frame: everything that sits at a position known from the choice alone
__global__ void __launch_bounds__(128)
k_alac_emit_fixed(const uint8_t* __restrict__ pcm, const AFrameset* __restrict__ fsets, ADev D, const AModel* __restrict__ models,
                  const AChoice* __restrict__ choice, const u64* __restrict__ fset_off, uint8_t* __restrict__ out,
                  const u64* __restrict__ total, u64 capacity)
{
    if (*total + 16 > capacity) return;
    const u32 fg = blockIdx.x, f = fg / D.ng, g = fg % D.ng, nch = D.grp[g].nch;
    const AFrameset fs = fsets[f];
    const AChoice c = choice[fg];
    const u32 n = fs.n;
    u32* words = (u32*)out;
    const u64 p0 = fset_off[f] * 8 + c.bit_off;
    const uint8_t* base = pcm + fs.pcm_off * D.C * D.B;
    const u32 stride = D.C * D.B, o0 = D.grp[g].c0 * D.B, o1 = D.grp[g].c1 * D.B;
    if (threadIdx.x == 0) {
        alac_field(words, p0, nch - 1, 3);
        // 16 unused bits | has-size | uncompressed LSB bytes | not-compressed | [size]   (alac.c:411-423, 563-575)
        u64 p = p0 + 3 + 16;
        alac_field(words, p, ((n != D.block_size ? 1u : 0u) << 3) | ((c.compressed ? D.lsb_bits / 8 : 0u) << 1) | (c.compressed ? 0u : 1u), 4);
        p += 4;
        if (n != D.block_size) { alac_field(words, p, n, 32); p += 32; }
        if (c.compressed) {
            alac_field(words, p, nch == 2 ? ALAC_SHIFT : 0u, 8);
            alac_field(words, p + 8, nch == 2 ? c.lw : 0u, 8);
        }
        if (g + 1 == D.ng) alac_field(words, p0 + 3 + c.bits, 7, 3);         // the frameset's closing '111'
    }
    if (c.compressed) {
        const ALayout L = alac_layout(c, n, nch, D);
        const u32 lwi = nch == 2 ? c.lw - D.lw_min : 0u;
        // write_subframe_header, alac.c:1125-1139: prediction type 0, shift 9, Rice modifier 4, order, coefficients
        for (u32 ch = 0; ch < nch; ch++) {
            const u64 sp = p0 + (ch ? L.sub1 : L.sub0);
            const AModel* m = models + (fg * D.upf + lwi * 2 + ch);
            const u32 order = c.order[ch];
            if (threadIdx.x == 0) alac_field(words, sp, (9u << 8) | (4u << 5) | order, 16);
            if (threadIdx.x < order) alac_field(words, sp + 16 + 16 * threadIdx.x, (u32)(unsigned short)(order == 8 ? m->q8[threadIdx.x] : m->q4[threadIdx.x]), 16);
        }
        if (D.lsb_bits) {
            const u32 mask = (1u << D.lsb_bits) - 1u;
            for (u32 e = threadIdx.x; e < n * nch; e += blockDim.x) {
                const u32 i = e / nch, ch = e % nch;
                const int s = alac_ld(base + (size_t)i * stride + (ch ? o1 : o0), D.B);
                alac_field(words, p0 + L.lsb + (u64)e * D.lsb_bits, (u32)s & mask, D.lsb_bits);
            }
        }
    } else {
        const u64 sp = p0 + 3 + alac_head_bits(n, D.block_size);
        for (u32 e = threadIdx.x; e < n * nch; e += blockDim.x) {
            const u32 i = e / nch, ch = e % nch;
            const int s = alac_ld(base + (size_t)i * stride + (ch ? o1 : o0), D.B);
            alac_field(words, sp + (u64)e * D.bps, (u32)s, D.bps);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static int alac_layout_groups(u32 channels, AGroup* grp)     // write_frameset, alac.c:288-373
{
    static const int L[9][5][2] = {
        {{0, 0}}, {{0, -1}}, {{0, 1}},
        {{2, -1}, {0, 1}}, {{2, -1}, {0, 1}, {3, -1}}, {{2, -1}, {0, 1}, {3, 4}},
        {{2, -1}, {0, 1}, {4, 5}, {3, -1}}, {{2, -1}, {0, 1}, {4, 5}, {6, -1}, {3, -1}},
        {{2, -1}, {6, 7}, {0, 1}, {4, 5}, {3, -1}}};
    static const int N[9] = {0, 1, 1, 2, 3, 3, 4, 5, 5};
    for (int g = 0; g < N[channels]; g++) {
        grp[g].c0 = (unsigned char)L[channels][g][0];
        grp[g].nch = L[channels][g][1] < 0 ? 1 : 2;
        grp[g].c1 = grp[g].nch == 2 ? (unsigned char)L[channels][g][1] : grp[g].c0;
        grp[g].pad = 0;
    }
    return N[channels];
}

static int check_params(const b200alac_params* p)
{
    if (!p) return afail("params is NULL");
    if (p->channels < 1 || p->channels > 8) return afail("unsupported channel count");
    if (p->bits_per_sample != 16 && p->bits_per_sample != 24) return afail("bits per sample must be 16 or 24");
    if (p->block_size < 1 || p->block_size > (1u << 24)) return afail("unsupported block size");
    if (p->maximum_k < 1 || p->maximum_k > 31) return afail("maximum_k must be 1..31");
    if (p->maximum_interlacing_leftweight < p->minimum_interlacing_leftweight ||
        p->maximum_interlacing_leftweight - p->minimum_interlacing_leftweight + 1 > ALAC_MAX_LW)
        return afail("unsupported interlacing leftweight range");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) { cudaGetLastError(); return afail("no CUDA device available: the B200 ALAC engine has no CPU fallback"); }
    return 0;
}

extern "C" uint64_t b200alac_output_bound(const b200alac_params* p, uint64_t n_pcm_frames, uint32_t n_framesets)
{
    // escape codes cost 9 bits more than the sample; headers, coefficient tables and padding per frameset
    return (n_pcm_frames * p->channels * (p->bits_per_sample + 1 + 9) + 7) / 8 + (u64)n_framesets * (p->channels * 64 + 16) + 64;
}

static void tukey(u32 N, double* w)      // window_signal, alac.c:787-810 (host libm, as the reference)
{
    const double alpha = 0.5;
    const unsigned window1 = (unsigned)(alpha * (N - 1)) / 2;
    const unsigned window2 = (unsigned)((N - 1) * (1.0 - (alpha / 2.0)));
    for (unsigned n = 0; n < N; n++) {
        if (n <= window1) w[n] = 0.5 * (1.0 + cos(M_PI * (((2 * n) / (alpha * (N - 1))) - 1.0)));
        else if (n <= window2) w[n] = 1.0;
        else w[n] = 0.5 * (1.0 + cos(M_PI * (((2.0 * n) / (alpha * (N - 1))) - (2.0 / alpha) + 1.0)));
    }
}

static int encode_core(const b200alac_params* p, const uint8_t* d_pcm, uint64_t n_pcm_frames, const uint32_t* lengths,
                       uint32_t n_lengths, int device, uint8_t* d_out, uint64_t out_capacity, uint8_t** host_out,
                       uint64_t* out_bytes, std::vector<uint32_t>* sizes, float* kernel_ms)
{
    int rc = 0;
    std::vector<AFrameset> fs;
    std::vector<double> win;
    std::map<u32, u32> woff;
    {
        u64 pos = 0;
        u32 li = 0;
        while (pos < n_pcm_frames) {
            u32 n;
            if (lengths) {
                if (li >= n_lengths || lengths[li] == 0) return afail("frame lengths do not add up to the PCM frame count");
                n = lengths[li++];
            } else n = (u32)((n_pcm_frames - pos) < p->block_size ? (n_pcm_frames - pos) : p->block_size);
            if (pos + n > n_pcm_frames) return afail("frame lengths do not add up to the PCM frame count");
            AFrameset f; f.pcm_off = pos; f.n = n; f.win_off = 0;
            if (n >= 10) {
                auto it = woff.find(n);
                if (it == woff.end()) {
                    it = woff.emplace(n, (u32)win.size()).first;
                    win.resize(win.size() + n);
                    tukey(n, win.data() + it->second);
                }
                f.win_off = it->second;
            }
            fs.push_back(f);
            pos += n;
        }
        if (lengths && li != n_lengths) return afail("frame lengths do not add up to the PCM frame count");
    }
    const u32 nf = (u32)fs.size();
    if (out_bytes) *out_bytes = 0;
    if (sizes) sizes->clear();
    if (nf == 0) { if (host_out) *host_out = (uint8_t*)malloc(1); return 0; }
    ADev D;
    memset(&D, 0, sizeof(D));
    D.C = p->channels; D.B = p->bits_per_sample / 8; D.bps = p->bits_per_sample;
    D.lsb_bits = p->bits_per_sample <= 16 ? 0 : p->bits_per_sample - 16;
    D.block_size = p->block_size; D.init_hist = p->initial_history; D.hist_mult = p->history_multiplier; D.max_k = p->maximum_k;
    D.lw_min = p->minimum_interlacing_leftweight;
    D.nlw = p->maximum_interlacing_leftweight - p->minimum_interlacing_leftweight + 1;
    D.upf = 2 * D.nlw;
    D.ng = (u32)alac_layout_groups(p->channels, D.grp);
    const u32 n_fg = nf * D.ng, n_units = n_fg * D.upf, n_chains = n_units * 2;
    AFrameset* d_fs = nullptr;
    double* d_win = nullptr;
    AModel* d_models = nullptr;
    u32 *d_bits = nullptr, *d_ov = nullptr, *d_fbytes = nullptr;
    AChoice* d_choice = nullptr;
    u64 *d_off = nullptr, *d_total = nullptr;
    uint8_t* d_own = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    u64 total = 0;
    ACK(cudaSetDevice(device));
    for (auto& e : ev) ACK(cudaEventCreate(&e));
    ACK(cudaMalloc((void**)&d_fs, nf * sizeof(AFrameset)));
    ACK(cudaMalloc((void**)&d_win, (win.size() + 2) * sizeof(double)));
    ACK(cudaMalloc((void**)&d_models, (size_t)n_units * sizeof(AModel)));
    ACK(cudaMalloc((void**)&d_bits, (size_t)n_chains * sizeof(u32)));
    ACK(cudaMalloc((void**)&d_ov, (size_t)n_fg * sizeof(u32)));
    ACK(cudaMalloc((void**)&d_choice, (size_t)n_fg * sizeof(AChoice)));
    ACK(cudaMalloc((void**)&d_fbytes, (size_t)nf * sizeof(u32)));
    ACK(cudaMalloc((void**)&d_off, (size_t)nf * sizeof(u64)));
    ACK(cudaMalloc((void**)&d_total, 64));
    ACK(cudaMemcpy(d_fs, fs.data(), nf * sizeof(AFrameset), cudaMemcpyHostToDevice));
    if (!win.empty()) ACK(cudaMemcpy(d_win, win.data(), win.size() * sizeof(double), cudaMemcpyHostToDevice));
    ACK(cudaMemset(d_models, 0, (size_t)n_units * sizeof(AModel)));
    ACK(cudaMemset(d_bits, 0, (size_t)n_chains * sizeof(u32)));
    ACK(cudaMemset(d_ov, 0, (size_t)n_fg * sizeof(u32)));
    ACK(cudaEventRecord(ev[0]));
    k_alac_model<<<(n_units + 63) / 64, 64>>>(d_pcm, d_fs, n_units, D, d_win, d_models);
    ACK(cudaGetLastError());
    ACK(cudaEventRecord(ev[1]));
    k_alac_size<<<dim3((n_units + 63) / 64, 2), 64>>>(d_pcm, d_fs, n_chains, D, d_models, d_bits, d_ov);
    ACK(cudaGetLastError());
    ACK(cudaEventRecord(ev[2]));
    k_alac_select<<<(nf + 127) / 128, 128>>>(d_fs, nf, D, d_models, d_bits, d_ov, d_choice, d_fbytes);
    k_alac_scan<<<1, 1024>>>(d_fbytes, nf, d_off, d_total);
    ACK(cudaGetLastError());
    ACK(cudaMemcpy(&total, d_total, sizeof(u64), cudaMemcpyDeviceToHost));
    if (!d_out) {
        out_capacity = ((total + 16 + 15) & ~15ull) + 16;
        ACK(cudaMalloc((void**)&d_own, (size_t)out_capacity));
        d_out = d_own;
    } else if (total + 16 > out_capacity) {
        rc = afail("encoded framesets exceed the output buffer");
        goto done;
    }
    ACK(cudaEventRecord(ev[3]));
    k_alac_zero<<<148 * 4, 256>>>((uint4*)d_out, d_total, out_capacity & ~15ull);
    k_alac_emit_residuals<<<(n_fg * 2 + 63) / 64, 64>>>(d_pcm, d_fs, n_fg, D, d_models, d_choice, d_off, d_out, d_total, out_capacity & ~15ull);
    k_alac_emit_fixed<<<n_fg, 128>>>(d_pcm, d_fs, D, d_models, d_choice, d_off, d_out, d_total, out_capacity & ~15ull);
    ACK(cudaGetLastError());
    ACK(cudaEventRecord(ev[4]));
    ACK(cudaDeviceSynchronize());
    if (kernel_ms) for (int i = 0; i < 4; i++) ACK(cudaEventElapsedTime(&kernel_ms[i], ev[i], ev[i + 1]));
    if (sizes) {
        sizes->resize(nf);
        ACK(cudaMemcpy(sizes->data(), d_fbytes, nf * sizeof(u32), cudaMemcpyDeviceToHost));
    }
    if (host_out) {
        *host_out = (uint8_t*)malloc((size_t)total + 1);
        if (!*host_out) { rc = afail("out of memory"); goto done; }
        ACK(cudaMemcpy(*host_out, d_out, (size_t)total, cudaMemcpyDeviceToHost));
    }
    if (out_bytes) *out_bytes = total;
done:
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    cudaFree(d_fs); cudaFree(d_win); cudaFree(d_models); cudaFree(d_bits); cudaFree(d_ov); cudaFree(d_choice);
    cudaFree(d_fbytes); cudaFree(d_off); cudaFree(d_total); cudaFree(d_own);
    return rc;
}

extern "C" int b200alac_encode_device(const b200alac_params* params, const void* d_pcm, uint64_t n_pcm_frames, int device,
                                      void* d_out, uint64_t out_capacity, uint64_t* out_bytes, uint32_t* frame_sizes,
                                      uint32_t* n_frames, float* kernel_ms)
{
    if (check_params(params)) return 1;
    if (!d_pcm || !d_out || ((uintptr_t)d_out & 15)) return afail("d_pcm/d_out is NULL or d_out is not 16-byte aligned");
    std::vector<uint32_t> sizes;
    if (encode_core(params, (const uint8_t*)d_pcm, n_pcm_frames, nullptr, 0, device, (uint8_t*)d_out, out_capacity, nullptr,
                    out_bytes, &sizes, kernel_ms)) return 1;
    if (frame_sizes) memcpy(frame_sizes, sizes.data(), sizes.size() * sizeof(uint32_t));
    if (n_frames) *n_frames = (uint32_t)sizes.size();
    return 0;
}

extern "C" int b200alac_encode_framesets(const b200alac_params* params, const uint8_t* pcm, uint64_t n_pcm_frames,
                                         const uint32_t* frame_lengths, uint32_t n_lengths, int device,
                                         uint8_t** out, uint64_t* out_bytes, uint32_t** frame_sizes, uint32_t* n_frames,
                                         float* kernel_ms)
{
    if (check_params(params)) return 1;
    if (!out || (!pcm && n_pcm_frames)) return afail("pcm/out is NULL");
    int rc = 0;
    uint8_t* d_pcm = nullptr;
    std::vector<uint32_t> sizes;
    const size_t nbytes = (size_t)n_pcm_frames * params->channels * (params->bits_per_sample / 8);
    *out = nullptr;
    ACK(cudaSetDevice(device));
    ACK(cudaMalloc((void**)&d_pcm, nbytes + 64));
    if (nbytes) ACK(cudaMemcpy(d_pcm, pcm, nbytes, cudaMemcpyHostToDevice));
    rc = encode_core(params, d_pcm, n_pcm_frames, frame_lengths, n_lengths, device, nullptr, 0, out, out_bytes, &sizes, kernel_ms);
    if (!rc) {
        if (frame_sizes) {
            *frame_sizes = (uint32_t*)malloc((sizes.size() ? sizes.size() : 1) * sizeof(uint32_t));
            memcpy(*frame_sizes, sizes.data(), sizes.size() * sizeof(uint32_t));
        }
        if (n_frames) *n_frames = (uint32_t)sizes.size();
    }
done:
    cudaFree(d_pcm);
    return rc;
}

extern "C" int b200alac_encode_mdat(const char* filename, const b200alac_params* params, const uint8_t* pcm,
                                    uint64_t n_pcm_frames, int device)
{
    if (!filename) return afail("filename is NULL");
    uint8_t* frames = nullptr;
    uint64_t nbytes = 0;
    if (b200alac_encode_framesets(params, pcm, n_pcm_frames, nullptr, 0, device, &frames, &nbytes, nullptr, nullptr, nullptr)) return 1;
    int rc = 0;
    FILE* f = fopen(filename, "wb");
    if (!f) {
        snprintf(a_err, sizeof(a_err), "cannot open \"%.300s\" for writing", filename);
        rc = 1;
    } else {
        const uint32_t size = (uint32_t)(nbytes + 8);       // alac.c:185-189
        const uint8_t head[8] = {(uint8_t)(size >> 24), (uint8_t)(size >> 16), (uint8_t)(size >> 8), (uint8_t)size, 'm', 'd', 'a', 't'};
        if (fwrite(head, 1, 8, f) != 8 || (nbytes && fwrite(frames, 1, (size_t)nbytes, f) != nbytes)) rc = afail("write error");
        if (fclose(f) != 0 && !rc) rc = afail("write error");
    }
    free(frames);
    return rc;
}
