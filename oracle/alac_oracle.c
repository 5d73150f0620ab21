/* alac_oracle.c -- CPU restatement of the reference's ALAC encoder.
 *
 * TEST INFRASTRUCTURE ONLY (like oracle/flac_oracle.c): imported by tests/ and nothing else; the product
 * path (python-audio-tools_b200/) never links, loads or executes it.
 *
 * Follows /root/reference/src/encoders/alac.c function by function:
 *   write_frameset :288-373 (channel pairing per channel count)      write_frame :375-403 (>= 10 frames: compressed,
 *   write_uncompressed_frame :405-436                                  a residual overflow falls back to uncompressed)
 *   write_compressed_frame :438-545 (leftweights 0..4, first strict minimum of the recorded bits; 24-bit: low byte raw)
 *   write_non_interlaced_frame :547-601   write_interlaced_frame :603-676   correlate_channels :678-718
 *   compute_coefficients :720-777 (orders 4 and 8, order 4 if bits4 < bits8 + 64)   window_signal :779-818
 *   autocorrelate :820-838   compute_lp_coefficients :840-883   quantize_coefficients :885-909
 *   calculate_residuals :936-1018 (adaptive coefficients)   encode_residuals :1034-1100   write_residual :1102-1122
 *   write_subframe_header :1125-1139   ALACEncoder_encode_alac :95-214 (mdat atom: size, "mdat", framesets)
 * PINNED: tests/test_alac_oracle.py compares whole outputs with the compiled, unmodified reference
 * (oracle/_ref/alacenc) and with tests/golden/alac_golden.json made from it.
 *
 * Compile with -ffp-contract=off: the reference's floating point runs without fused multiply-add.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAX_LPC_ORDER 8
#define INTERLACING_SHIFT 2

typedef struct {
    unsigned block_size, initial_history, history_multiplier, maximum_k;
    unsigned min_leftweight, max_leftweight;
    unsigned bits_per_sample;
} alac_options;

/* ---- MSB-first bit recorder ---- */
typedef struct {
    uint8_t *buf;
    size_t cap;
    uint64_t bits;
} bw;

static void bw_reset(bw *w) { w->bits = 0; }

static void bw_put(bw *w, unsigned count, uint64_t value)
{
    while (count) {
        const size_t byte = (size_t)(w->bits >> 3);
        const unsigned used = (unsigned)(w->bits & 7), room = 8 - used;
        const unsigned take = count < room ? count : room;
        if (byte >= w->cap) {
            const size_t ncap = w->cap ? w->cap * 2 : 4096;
            w->buf = (uint8_t *)realloc(w->buf, ncap);
            memset(w->buf + w->cap, 0, ncap - w->cap);
            w->cap = ncap;
        }
        if (used == 0) w->buf[byte] = 0;
        const unsigned chunk = (unsigned)((value >> (count - take)) & ((1u << take) - 1u));
        w->buf[byte] |= (uint8_t)(chunk << (room - take));
        w->bits += take;
        count -= take;
    }
}

static void bw_put_signed(bw *w, unsigned count, int value)
{
    bw_put(w, count, (uint64_t)(uint32_t)value & ((count >= 32) ? 0xFFFFFFFFu : ((1u << count) - 1u)));
}

/* write_unary(bs, 0, n): n one-bits, then a zero */
static void bw_unary0(bw *w, unsigned n)
{
    while (n > 0) {
        const unsigned c = n <= 30 ? n : 30;
        bw_put(w, c, (1u << c) - 1u);
        n -= c;
    }
    bw_put(w, 1, 0);
}

static void bw_copy(bw *dst, const bw *src)
{
    uint64_t i = 0;
    for (; i + 8 <= src->bits; i += 8) bw_put(dst, 8, src->buf[i >> 3]);
    if (i < src->bits) bw_put(dst, (unsigned)(src->bits - i), src->buf[i >> 3] >> (8 - (src->bits - i)));
}

/* ---- the model ---- */
static void window_signal(const int *s, unsigned N, double *out)   /* :779-818 */
{
    const double alpha = 0.5;
    const unsigned window1 = (unsigned)(alpha * (N - 1)) / 2;
    const unsigned window2 = (unsigned)((N - 1) * (1.0 - (alpha / 2.0)));
    for (unsigned n = 0; n < N; n++) {
        double w;
        if (n <= window1) w = 0.5 * (1.0 + cos(M_PI * (((2 * n) / (alpha * (N - 1))) - 1.0)));
        else if (n <= window2) w = 1.0;
        else w = 0.5 * (1.0 + cos(M_PI * (((2.0 * n) / (alpha * (N - 1))) - (2.0 / alpha) + 1.0)));
        out[n] = s[n] * w;
    }
}

static void autocorrelate(const double *w, unsigned N, double *r)     /* :820-838 */
{
    for (unsigned lag = 0; lag <= MAX_LPC_ORDER; lag++) {
        double acc = 0.0;
        for (unsigned i = 0; i < N - lag; i++) acc += w[i] * w[i + lag];
        r[lag] = acc;
    }
}

static void lp_coefficients(const double *r, double lp[MAX_LPC_ORDER][MAX_LPC_ORDER])   /* :840-883 */
{
    double err[MAX_LPC_ORDER];
    double k = r[1] / r[0];
    lp[0][0] = k;
    err[0] = r[0] * (1.0 - (k * k));
    for (unsigned i = 1; i < MAX_LPC_ORDER; i++) {
        double q = r[i + 1];
        for (unsigned j = 0; j < i; j++) q -= lp[i - 1][j] * r[i - j];
        k = q / err[i - 1];
        for (unsigned j = 0; j < i; j++) lp[i][j] = lp[i - 1][j] - (k * lp[i - 1][i - j - 1]);
        lp[i][i] = k;
        err[i] = err[i - 1] * (1.0 - (k * k));
    }
}

static void quantize(const double *lp, unsigned order, int *q)       /* :885-909 */
{
    const int qmax = (1 << 15) - 1, qmin = -(1 << 15);
    double error = 0.0;
    for (unsigned i = 0; i < order; i++) {
        error += lp[i] * (1 << 9);
        const int ei = (int)round(error);
        q[i] = ei < qmin ? qmin : ei > qmax ? qmax : ei;
        error -= (double)ei;
    }
}

static int sign_only(int v) { return v > 0 ? 1 : v < 0 ? -1 : 0; }

static int truncate_bits(int value, unsigned bits)
{
    const int t = value & ((1 << bits) - 1);
    return (t & (1 << (bits - 1))) ? t - (1 << bits) : t;
}

static void calculate_residuals(const int *s, unsigned n, unsigned sample_size, const int *q, unsigned cc, int *res)   /* :936-1018 */
{
    int coef[MAX_LPC_ORDER];
    unsigned i = 0;
    memcpy(coef, q, cc * sizeof(int));
    res[i] = s[i]; i++;
    for (; i < cc + 1; i++) res[i] = truncate_bits(s[i] - s[i - 1], sample_size);
    for (; i < n; i++) {
        const int base = s[i - cc - 1];
        int64_t sum = 1 << 8;
        for (unsigned j = 0; j < cc; j++) sum += (int64_t)coef[j] * (int64_t)(s[i - j - 1] - base);
        sum >>= 9;
        int error = truncate_bits(s[i] - base - (int)sum, sample_size);
        res[i] = error;
        if (error > 0) {
            for (unsigned j = 0; j < cc; j++) {
                const int diff = base - s[i - cc + j];
                const int sg = sign_only(diff);
                coef[cc - j - 1] -= sg;
                error -= ((diff * sg) >> 9) * (int)(j + 1);
                if (error <= 0) break;
            }
        } else if (error < 0) {
            for (unsigned j = 0; j < cc; j++) {
                const int diff = base - s[i - cc + j];
                const int sg = sign_only(diff);
                coef[cc - j - 1] += sg;
                error -= ((diff * -sg) >> 9) * (int)(j + 1);
                if (error >= 0) break;
            }
        }
    }
}

static unsigned log2u(unsigned v)       /* LOG2 :1020-1031 (0 -> UINT_MAX, as the NDEBUG build computes it) */
{
    unsigned bits = 0;
    while (v) { bits++; v >>= 1; }
    return bits - 1;
}

static void write_residual(bw *w, unsigned value, unsigned k, unsigned sample_size)   /* :1102-1122 */
{
    const unsigned msb = value / ((1u << k) - 1), lsb = value % ((1u << k) - 1);
    if (msb > 8) {
        bw_put(w, 9, 0x1FF);
        bw_put(w, sample_size, value);
    } else {
        bw_unary0(w, msb);
        if (k > 1) {
            if (lsb > 0) bw_put(w, k, lsb + 1);
            else bw_put(w, k - 1, 0);
        }
    }
}

/* returns 0, or 1 on residual overflow (the reference longjmps, :1059-1063) */
static int encode_residuals(const alac_options *o, unsigned sample_size, const int *res, unsigned n, bw *w)   /* :1034-1100 */
{
    int history = (int)o->initial_history;
    unsigned sign_modifier = 0, i = 0;
    const unsigned max_unsigned = 1u << sample_size;
    while (i < n) {
        unsigned u;
        if (res[i] >= 0) u = (unsigned)(res[i] << 1);
        else u = (unsigned)(-res[i] << 1) - 1;
        if (u >= max_unsigned) return 1;
        unsigned k = log2u((unsigned)((history >> 9) + 3));
        if (k > o->maximum_k) k = o->maximum_k;
        write_residual(w, u - sign_modifier, k, sample_size);
        sign_modifier = 0;
        if (u <= 0xFFFF) {
            history += ((int)(u * o->history_multiplier) - ((history * (int)o->history_multiplier) >> 9));
            i++;
            if ((history < 128) && (i < n)) {
                unsigned zeroes = 0;
                k = 7 - log2u((unsigned)history) + (unsigned)((history + 16) >> 6);
                if (k > o->maximum_k) k = o->maximum_k;
                while ((i < n) && (res[i] == 0)) { zeroes++; i++; }
                write_residual(w, zeroes, k, 16);
                if (zeroes < 0xFFFF) sign_modifier = 1;
                history = 0;
            }
        } else {
            i++;
            history = 0xFFFF;
        }
    }
    return 0;
}

typedef struct {
    double *windowed;
    int *res4, *res8;
    bw blk4, blk8;
} scratch;

/* compute_coefficients :720-777; returns 1 on residual overflow */
static int compute_coefficients(const alac_options *o, scratch *sc, const int *s, unsigned n, unsigned sample_size,
                                int *q, unsigned *order, bw *residual)
{
    double r[MAX_LPC_ORDER + 1];
    window_signal(s, n, sc->windowed);
    autocorrelate(sc->windowed, n, r);
    if (r[0] != 0.0) {
        double lp[MAX_LPC_ORDER][MAX_LPC_ORDER];
        int q4[4], q8[8];
        lp_coefficients(r, lp);
        quantize(lp[3], 4, q4);
        quantize(lp[7], 8, q8);
        calculate_residuals(s, n, sample_size, q4, 4, sc->res4);
        calculate_residuals(s, n, sample_size, q8, 8, sc->res8);
        bw_reset(&sc->blk4);
        if (encode_residuals(o, sample_size, sc->res4, n, &sc->blk4)) return 1;
        bw_reset(&sc->blk8);
        if (encode_residuals(o, sample_size, sc->res8, n, &sc->blk8)) return 1;
        if (sc->blk4.bits < sc->blk8.bits + 64) { memcpy(q, q4, sizeof(q4)); *order = 4; bw_copy(residual, &sc->blk4); }
        else { memcpy(q, q8, sizeof(q8)); *order = 8; bw_copy(residual, &sc->blk8); }
    } else {
        memset(q, 0, 4 * sizeof(int));
        *order = 4;
        calculate_residuals(s, n, sample_size, q, 4, sc->res4);
        if (encode_residuals(o, sample_size, sc->res4, n, residual)) return 1;
    }
    return 0;
}

static void write_subframe_header(bw *w, const int *q, unsigned order)    /* :1125-1139 */
{
    bw_put(w, 4, 0);
    bw_put(w, 4, 9);
    bw_put(w, 3, 4);
    bw_put(w, 5, order);
    for (unsigned i = 0; i < order; i++) bw_put_signed(w, 16, q[i]);
}

static void frame_head(bw *w, const alac_options *o, unsigned n, unsigned lsbs, unsigned not_compressed)
{
    bw_put(w, 16, 0);
    bw_put(w, 1, n == o->block_size ? 0 : 1);
    bw_put(w, 2, lsbs);
    bw_put(w, 1, not_compressed);
    if (n != o->block_size) bw_put(w, 32, n);
}

static void write_uncompressed_frame(bw *w, const alac_options *o, int **ch, unsigned nch, unsigned n)   /* :405-436 */
{
    frame_head(w, o, n, 0, 1);
    for (unsigned i = 0; i < n; i++)
        for (unsigned c = 0; c < nch; c++) bw_put_signed(w, o->bits_per_sample, ch[c][i]);
}

/* write_compressed_frame :438-545; returns 1 on residual overflow */
static int write_compressed_frame(bw *out, const alac_options *o, scratch *sc, int **ch, unsigned nch, unsigned n)
{
    const unsigned lsb_bytes = o->bits_per_sample <= 16 ? 0 : (o->bits_per_sample - 16) / 8;
    const unsigned lsb_bits = lsb_bytes * 8;
    int *msb[2] = {NULL, NULL}, *cor[2] = {NULL, NULL};
    int rc = 0;
    bw res0 = {0}, res1 = {0}, cand = {0}, best = {0};
    for (unsigned c = 0; c < nch; c++) {
        msb[c] = (int *)malloc((size_t)n * sizeof(int));
        cor[c] = (int *)malloc((size_t)n * sizeof(int));
        for (unsigned i = 0; i < n; i++) msb[c][i] = lsb_bits ? ch[c][i] >> lsb_bits : ch[c][i];
    }
    if (nch == 1) {
        int q[8];
        unsigned order;
        frame_head(out, o, n, lsb_bytes, 0);
        bw_put(out, 8, 0);
        bw_put(out, 8, 0);
        rc = compute_coefficients(o, sc, msb[0], n, o->bits_per_sample - lsb_bits, q, &order, &res0);
        if (!rc) {
            write_subframe_header(out, q, order);
            if (lsb_bits) for (unsigned i = 0; i < n; i++) bw_put(out, lsb_bits, (unsigned)ch[0][i] & ((1u << lsb_bits) - 1u));
            bw_copy(out, &res0);
        }
    } else {
        uint64_t best_bits = ~0ull;
        for (unsigned lw = o->min_leftweight; lw <= o->max_leftweight && !rc; lw++) {
            int q0[8], q1[8];
            unsigned o0, o1;
            bw_reset(&cand); bw_reset(&res0); bw_reset(&res1);
            frame_head(&cand, o, n, lsb_bytes, 0);
            bw_put(&cand, 8, INTERLACING_SHIFT);
            bw_put(&cand, 8, lw);
            /* correlate_channels :678-718 */
            if (lw > 0) {
                for (unsigned i = 0; i < n; i++) {
                    int64_t t = msb[0][i] - msb[1][i];
                    t *= lw;
                    t >>= INTERLACING_SHIFT;
                    cor[0][i] = msb[1][i] + (int)t;
                    cor[1][i] = msb[0][i] - msb[1][i];
                }
            } else {
                memcpy(cor[0], msb[0], (size_t)n * sizeof(int));
                memcpy(cor[1], msb[1], (size_t)n * sizeof(int));
            }
            const unsigned ss = o->bits_per_sample - lsb_bits + 1;
            rc = compute_coefficients(o, sc, cor[0], n, ss, q0, &o0, &res0);
            if (!rc) rc = compute_coefficients(o, sc, cor[1], n, ss, q1, &o1, &res1);
            if (rc) break;
            write_subframe_header(&cand, q0, o0);
            write_subframe_header(&cand, q1, o1);
            if (lsb_bits)
                for (unsigned i = 0; i < n; i++)
                    for (unsigned c = 0; c < 2; c++) bw_put(&cand, lsb_bits, (unsigned)ch[c][i] & ((1u << lsb_bits) - 1u));
            bw_copy(&cand, &res0);
            bw_copy(&cand, &res1);
            if (cand.bits < best_bits) {
                best_bits = cand.bits;
                bw t = cand; cand = best; best = t;
            }
        }
        if (!rc) bw_copy(out, &best);
    }
    for (unsigned c = 0; c < nch; c++) { free(msb[c]); free(cor[c]); }
    free(res0.buf); free(res1.buf); free(cand.buf); free(best.buf);
    return rc;
}

static void write_frame(bw *out, const alac_options *o, scratch *sc, int **ch, unsigned nch, unsigned n)   /* :375-403 */
{
    bw_put(out, 3, nch - 1);
    if (n >= 10) {
        bw tmp = {0};
        if (!write_compressed_frame(&tmp, o, sc, ch, nch, n)) bw_copy(out, &tmp);
        else write_uncompressed_frame(out, o, ch, nch, n);
        free(tmp.buf);
    } else {
        write_uncompressed_frame(out, o, ch, nch, n);
    }
}

/* write_frameset :288-373: which channels go together, in which order */
static unsigned frameset_layout(unsigned channels, unsigned groups[8][2], unsigned sizes[8])
{
    static const int L[9][5][2] = {
        {{0}}, {{0, -1}}, {{0, 1}},
        {{2, -1}, {0, 1}}, {{2, -1}, {0, 1}, {3, -1}}, {{2, -1}, {0, 1}, {3, 4}},
        {{2, -1}, {0, 1}, {4, 5}, {3, -1}}, {{2, -1}, {0, 1}, {4, 5}, {6, -1}, {3, -1}},
        {{2, -1}, {6, 7}, {0, 1}, {4, 5}, {3, -1}}};
    static const unsigned N[9] = {0, 1, 1, 2, 3, 3, 4, 5, 5};
    if (channels > 8) {
        for (unsigned c = 0; c < channels && c < 8; c++) { groups[c][0] = c; sizes[c] = 1; }
        return channels;
    }
    for (unsigned g = 0; g < N[channels]; g++) {
        groups[g][0] = (unsigned)L[channels][g][0];
        sizes[g] = L[channels][g][1] < 0 ? 1 : 2;
        if (sizes[g] == 2) groups[g][1] = (unsigned)L[channels][g][1];
    }
    return N[channels];
}

static int unpack(const uint8_t *p, unsigned bytes)
{
    if (bytes == 2) return (int16_t)(p[0] | (p[1] << 8));
    return ((int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24)) >> 8;
}

/* Framesets only, back to back (each byte aligned), and their sizes: what ALACEncoder_encode_alac writes after the
 * 8-byte mdat header, and what alac_log_output returns.  frame_lengths (optional): PCM frames per frameset.
 * *out is malloc'd.  Returns the number of framesets. */
unsigned alac_oracle_encode_framesets(const uint8_t *pcm, uint64_t n_pcm_frames, unsigned channels, unsigned bps,
                                      unsigned block_size, unsigned initial_history, unsigned history_multiplier,
                                      unsigned maximum_k, unsigned min_leftweight, unsigned max_leftweight,
                                      const uint32_t *frame_lengths, unsigned n_lengths,
                                      uint8_t **out, uint64_t *out_bytes, uint32_t *frame_sizes)
{
    alac_options o = {block_size, initial_history, history_multiplier, maximum_k, min_leftweight, max_leftweight, bps};
    const unsigned B = bps / 8;
    unsigned maxn = block_size;
    for (unsigned i = 0; i < n_lengths; i++) if (frame_lengths[i] > maxn) maxn = frame_lengths[i];
    scratch sc;
    memset(&sc, 0, sizeof(sc));
    sc.windowed = (double *)malloc((size_t)maxn * sizeof(double));
    sc.res4 = (int *)malloc((size_t)maxn * sizeof(int));
    sc.res8 = (int *)malloc((size_t)maxn * sizeof(int));
    int **ch = (int **)malloc(channels * sizeof(int *));
    for (unsigned c = 0; c < channels; c++) ch[c] = (int *)malloc((size_t)maxn * sizeof(int));
    bw w = {0};
    uint64_t pos = 0;
    unsigned nf = 0;
    unsigned groups[8][2], sizes[8];
    const unsigned ng = frameset_layout(channels, groups, sizes);
    while (pos < n_pcm_frames) {
        const unsigned n = frame_lengths ? frame_lengths[nf] : (unsigned)((n_pcm_frames - pos) < block_size ? (n_pcm_frames - pos) : block_size);
        const uint64_t start = w.bits;
        for (unsigned i = 0; i < n; i++)
            for (unsigned c = 0; c < channels; c++) ch[c][i] = unpack(pcm + ((pos + i) * channels + c) * B, B);
        for (unsigned g = 0; g < ng; g++) {
            int *pair[2];
            pair[0] = ch[groups[g][0]];
            if (sizes[g] == 2) pair[1] = ch[groups[g][1]];
            write_frame(&w, &o, &sc, pair, sizes[g], n);
        }
        bw_put(&w, 3, 7);
        if (w.bits & 7) bw_put(&w, 8 - (unsigned)(w.bits & 7), 0);
        if (frame_sizes) frame_sizes[nf] = (uint32_t)((w.bits - start) >> 3);
        nf++;
        pos += n;
    }
    for (unsigned c = 0; c < channels; c++) free(ch[c]);
    free(ch); free(sc.windowed); free(sc.res4); free(sc.res8); free(sc.blk4.buf); free(sc.blk8.buf);
    *out = w.buf ? w.buf : (uint8_t *)malloc(1);
    *out_bytes = w.bits >> 3;
    return nf;
}

/* the standalone driver's output (:95-214): 32-bit size, "mdat", framesets */
uint64_t alac_oracle_encode_mdat(const uint8_t *pcm, uint64_t n_pcm_frames, unsigned channels, unsigned bps,
                                 unsigned block_size, unsigned initial_history, unsigned history_multiplier,
                                 unsigned maximum_k, uint8_t **out)
{
    uint8_t *frames = NULL;
    uint64_t nbytes = 0;
    alac_oracle_encode_framesets(pcm, n_pcm_frames, channels, bps, block_size, initial_history, history_multiplier,
                                 maximum_k, 0, 4, NULL, 0, &frames, &nbytes, NULL);
    uint8_t *f = (uint8_t *)malloc((size_t)nbytes + 9);
    const uint32_t size = (uint32_t)(nbytes + 8);
    f[0] = (uint8_t)(size >> 24); f[1] = (uint8_t)(size >> 16); f[2] = (uint8_t)(size >> 8); f[3] = (uint8_t)size;
    memcpy(f + 4, "mdat", 4);
    if (nbytes) memcpy(f + 8, frames, (size_t)nbytes);
    free(frames);
    *out = f;
    return nbytes + 8;
}

void alac_oracle_free(void *p) { free(p); }
