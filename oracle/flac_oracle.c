/* flac_oracle.c -- plain-C restatement of the reference FLAC encoder
 * (src/encoders/flac.c of widgital/python-audio-tools), written as the parity
 * checker for the B200 engine.  TEST INFRASTRUCTURE ONLY -- see flac_oracle.h.
 *
 * The structure follows the reference function by function so that each
 * decision can be audited against the cited lines, but the code is our own:
 * a flat bit-vector writer replaces the reference's BitstreamWriter
 * recorder/accumulator objects and plain C arrays replace a_int/l_int.
 *
 * All paths cited are relative to /root/reference.
 */
#include "flac_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* bit writer: MSB-first, growable (src/bitstream.c:1905-1947 semantics) */
/* ------------------------------------------------------------------ */
typedef struct {
    uint8_t *buf;
    size_t cap;      /* bytes */
    uint64_t nbits;  /* bits written */
} bitvec;

static void bv_init(bitvec *b) { b->buf = NULL; b->cap = 0; b->nbits = 0; }
static void bv_free(bitvec *b) { free(b->buf); b->buf = NULL; b->cap = 0; b->nbits = 0; }
static void bv_reset(bitvec *b) { b->nbits = 0; }

static void bv_reserve(bitvec *b, uint64_t more_bits)
{
    size_t need = (size_t)((b->nbits + more_bits + 7) / 8) + 8;
    if (need > b->cap) {
        size_t ncap = b->cap ? b->cap : 256;
        while (ncap < need) ncap *= 2;
        b->buf = (uint8_t *)realloc(b->buf, ncap);
        b->cap = ncap;
    }
}

/* write the low `count` bits of value, most significant first (count <= 64) */
static void bv_put(bitvec *b, unsigned count, uint64_t value)
{
    bv_reserve(b, count);
    while (count > 0) {
        unsigned bitpos = (unsigned)(b->nbits & 7);
        unsigned room = 8 - bitpos;
        unsigned take = count < room ? count : room;
        unsigned chunk = (unsigned)((value >> (count - take)) & ((1u << take) - 1));
        uint8_t *p = b->buf + (b->nbits >> 3);
        if (bitpos == 0) *p = 0;
        *p |= (uint8_t)(chunk << (room - take));
        b->nbits += take;
        count -= take;
    }
}

/* src/bitstream.c:2020-2034 : sign bit then count-1 bits (two's complement) */
static void bv_put_signed(bitvec *b, unsigned count, int value)
{
    if (value >= 0) {
        bv_put(b, 1, 0);
        bv_put(b, count - 1, (uint64_t)(unsigned)value);
    } else {
        bv_put(b, 1, 1);
        bv_put(b, count - 1, (uint64_t)(unsigned)((1 << (count - 1)) + value));
    }
}

/* src/bitstream.c:2233-2251 : `value` copies of !stop_bit, then stop_bit */
static void bv_put_unary(bitvec *b, int stop_bit, uint64_t value)
{
    while (value > 0) {
        unsigned n = value > 30 ? 30 : (unsigned)value;
        bv_put(b, n, stop_bit ? 0 : ((1u << n) - 1));
        value -= n;
    }
    bv_put(b, 1, (uint64_t)stop_bit);
}

static void bv_align(bitvec *b) /* src/bitstream.c:2299-2305 */
{
    if (b->nbits & 7) bv_put(b, 8 - (unsigned)(b->nbits & 7), 0);
}

/* append all bits of src (bw_rec_copy, src/bitstream.c:2754-2823) */
static void bv_append(bitvec *dst, const bitvec *src)
{
    uint64_t full = src->nbits >> 3;
    unsigned rest = (unsigned)(src->nbits & 7);
    uint64_t i;
    bv_reserve(dst, src->nbits);
    if ((dst->nbits & 7) == 0) {
        memcpy(dst->buf + (dst->nbits >> 3), src->buf, (size_t)full);
        dst->nbits += full * 8;
    } else {
        for (i = 0; i < full; i++) bv_put(dst, 8, src->buf[i]);
    }
    if (rest) bv_put(dst, rest, (uint64_t)(src->buf[full] >> (8 - rest)));
}

/* ------------------------------------------------------------------ */
/* CRCs (src/common/flac_crc.c): MSB-first, init 0, no reflection      */
/* ------------------------------------------------------------------ */
uint8_t orc_crc8(const uint8_t *data, size_t len) /* poly x^8+x^2+x+1 */
{
    unsigned crc = 0;
    size_t i;
    int k;
    for (i = 0; i < len; i++) {
        crc ^= data[i];
        for (k = 0; k < 8; k++)
            crc = (crc & 0x80) ? ((crc << 1) ^ 0x07) & 0xFF : (crc << 1) & 0xFF;
    }
    return (uint8_t)crc;
}

uint16_t orc_crc16(const uint8_t *data, size_t len) /* poly x^16+x^15+x^2+1 */
{
    unsigned crc = 0;
    size_t i;
    int k;
    for (i = 0; i < len; i++) {
        crc ^= (unsigned)data[i] << 8;
        for (k = 0; k < 8; k++)
            crc = (crc & 0x8000) ? ((crc << 1) ^ 0x8005) & 0xFFFF : (crc << 1) & 0xFFFF;
    }
    return (uint16_t)crc;
}

/* ------------------------------------------------------------------ */
/* MD5 (RFC 1321; the reference vendors the same algorithm in           */
/* src/common/md5.c:177-273)                                            */
/* ------------------------------------------------------------------ */
typedef struct {
    uint32_t h[4];
    uint64_t len;
    uint8_t block[64];
    unsigned fill;
} orc_md5_ctx;

static const uint32_t md5_k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
static const uint8_t md5_s[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                                  5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                                  4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                                  6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};

static void md5_block(orc_md5_ctx *c, const uint8_t *p)
{
    uint32_t m[16], a = c->h[0], b = c->h[1], cc = c->h[2], d = c->h[3];
    int i;
    for (i = 0; i < 16; i++)
        m[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
               ((uint32_t)p[4 * i + 3] << 24);
    for (i = 0; i < 64; i++) {
        uint32_t f, t;
        int g;
        if (i < 16) { f = (b & cc) | (~b & d); g = i; }
        else if (i < 32) { f = (d & b) | (~d & cc); g = (5 * i + 1) & 15; }
        else if (i < 48) { f = b ^ cc ^ d; g = (3 * i + 5) & 15; }
        else { f = cc ^ (b | ~d); g = (7 * i) & 15; }
        t = a + f + md5_k[i] + m[g];
        a = d; d = cc; cc = b;
        b = b + ((t << md5_s[i]) | (t >> (32 - md5_s[i])));
    }
    c->h[0] += a; c->h[1] += b; c->h[2] += cc; c->h[3] += d;
}

static void md5_init(orc_md5_ctx *c)
{
    c->h[0] = 0x67452301; c->h[1] = 0xefcdab89; c->h[2] = 0x98badcfe; c->h[3] = 0x10325476;
    c->len = 0; c->fill = 0;
}

static void md5_update(orc_md5_ctx *c, const uint8_t *p, size_t n)
{
    c->len += n;
    while (n > 0) {
        size_t take = 64 - c->fill;
        if (take > n) take = n;
        memcpy(c->block + c->fill, p, take);
        c->fill += (unsigned)take; p += take; n -= take;
        if (c->fill == 64) { md5_block(c, c->block); c->fill = 0; }
    }
}

static void md5_final(orc_md5_ctx *c, uint8_t digest[16])
{
    uint64_t bits = c->len * 8;
    uint8_t pad[72];
    size_t padlen = (c->fill < 56) ? (56 - c->fill) : (120 - c->fill);
    int i;
    memset(pad, 0, sizeof(pad));
    pad[0] = 0x80;
    for (i = 0; i < 8; i++) pad[padlen + i] = (uint8_t)(bits >> (8 * i));
    md5_update(c, pad, padlen + 8);
    for (i = 0; i < 16; i++) digest[i] = (uint8_t)(c->h[i >> 2] >> (8 * (i & 3)));
}

void orc_md5(const uint8_t *data, size_t len, uint8_t digest[16])
{
    orc_md5_ctx c;
    md5_init(&c);
    md5_update(&c, data, len);
    md5_final(&c, digest);
}

/* ------------------------------------------------------------------ */
/* encoder context (src/encoders/flac.h:62-95)                          */
/* ------------------------------------------------------------------ */
struct orc_encoder {
    orc_options opt;
    unsigned sample_rate, channels, bits_per_sample;
    unsigned qlp_coeff_precision; /* flac.c:165-178 */
    unsigned max_rice_parameter;  /* flac.c:180-184 */
    /* cached Tukey window (flac.c:1139) */
    double *window;
    unsigned window_len;
    /* scratch */
    int32_t *scratch[8];
    unsigned scratch_len;
    double *windowed;
    bitvec cand_bits[4];
    bitvec fixed_bits, lpc_bits, frame_bits;
};

orc_encoder *orc_encoder_new(const orc_options *opt, unsigned sample_rate, unsigned channels,
                             unsigned bits_per_sample)
{
    orc_encoder *e = (orc_encoder *)calloc(1, sizeof(*e));
    unsigned bs = opt->block_size;
    int i;
    e->opt = *opt;
    e->sample_rate = sample_rate;
    e->channels = channels;
    e->bits_per_sample = bits_per_sample;
    /* flac.c:165-178 */
    if (bs <= 192) e->qlp_coeff_precision = 7;
    else if (bs <= 384) e->qlp_coeff_precision = 8;
    else if (bs <= 576) e->qlp_coeff_precision = 9;
    else if (bs <= 1152) e->qlp_coeff_precision = 10;
    else if (bs <= 2304) e->qlp_coeff_precision = 11;
    else if (bs <= 4608) e->qlp_coeff_precision = 12;
    else e->qlp_coeff_precision = 13;
    /* flac.c:180-184 */
    e->max_rice_parameter = (bits_per_sample <= 16) ? 0xE : 0x1E;
    for (i = 0; i < 4; i++) bv_init(&e->cand_bits[i]);
    bv_init(&e->fixed_bits);
    bv_init(&e->lpc_bits);
    bv_init(&e->frame_bits);
    return e;
}

void orc_encoder_free(orc_encoder *e)
{
    int i;
    if (!e) return;
    free(e->window);
    free(e->windowed);
    for (i = 0; i < 8; i++) free(e->scratch[i]);
    for (i = 0; i < 4; i++) bv_free(&e->cand_bits[i]);
    bv_free(&e->fixed_bits);
    bv_free(&e->lpc_bits);
    bv_free(&e->frame_bits);
    free(e);
}

static void ensure_scratch(orc_encoder *e, unsigned n)
{
    int i;
    if (n <= e->scratch_len) return;
    for (i = 0; i < 8; i++) e->scratch[i] = (int32_t *)realloc(e->scratch[i], sizeof(int32_t) * (n + 8));
    e->windowed = (double *)realloc(e->windowed, sizeof(double) * (n + 8));
    e->scratch_len = n;
}

static void decision_clear(orc_subframe_decision *d)
{
    if (d) { free(d->rice); memset(d, 0, sizeof(*d)); }
}

void orc_frame_decision_free(orc_frame_decision *d)
{
    unsigned i;
    if (!d) return;
    for (i = 0; i < ORC_MAX_CHANNELS; i++) decision_clear(&d->sub[i]);
    for (i = 0; i < 4; i++) decision_clear(&d->cand[i]);
    memset(d, 0, sizeof(*d));
}

static void decision_copy(orc_subframe_decision *dst, const orc_subframe_decision *src)
{
    decision_clear(dst);
    *dst = *src;
    if (src->rice && src->n_partitions) {
        dst->rice = (uint8_t *)malloc(src->n_partitions);
        memcpy(dst->rice, src->rice, src->n_partitions);
    } else {
        dst->rice = NULL;
    }
}

/* ------------------------------------------------------------------ */
/* small sample scans                                                   */
/* ------------------------------------------------------------------ */
static int all_identical(const int32_t *s, unsigned n) /* flac.c:1606-1620 */
{
    unsigned i;
    for (i = 1; i < n; i++)
        if (s[i] != s[0]) return 0;
    return 1;
}

unsigned orc_wasted_bits(const int32_t *s, unsigned n) /* flac.c:1578-1604 */
{
    unsigned best = INT_MAX, i;
    for (i = 0; i < n; i++) {
        int v = s[i];
        if (v != 0) {
            unsigned w = 0;
            while (((v & 1) == 0) && (v != 0)) { v >>= 1; w++; }
            if (w < best) best = w;
            if (best == 0) return 0;
        }
    }
    return (best == INT_MAX) ? 0 : best;
}

/* ------------------------------------------------------------------ */
/* residual coding                                                      */
/* ------------------------------------------------------------------ */
/* flac.c:1437-1505 flacenc_encode_residual_partitions.
 * part_len[p] receives the number of residuals actually placed in partition p
 * (the clamped l_int split of src/array.c:559-580). */
uint64_t orc_residual_partitions(const int32_t *residuals, unsigned n_residuals,
                                 unsigned block_size, unsigned predictor_order,
                                 unsigned partition_order, unsigned max_rice, uint8_t *rice,
                                 unsigned *part_len)
{
    uint64_t total = 0;
    unsigned p, pos = 0;
    for (p = 0; p < (1u << partition_order); p++) {
        unsigned plength, take, i, k;
        uint64_t sum = 0;
        /* flac.c:1461-1465: unsigned arithmetic, may wrap below zero */
        if (p == 0) plength = (block_size >> partition_order) - predictor_order;
        else plength = block_size >> partition_order;
        take = plength < (n_residuals - pos) ? plength : (n_residuals - pos);
        for (i = 0; i < take; i++) {
            int32_t r = residuals[pos + i];
            /* flac.c:1469-1474: uint64 += / -= of an int */
            if (r >= 0) sum += (uint64_t)(int64_t)r;
            else sum -= (uint64_t)(int64_t)r;
        }
        pos += take;
        /* flac.c:1477-1484: the shift is evaluated in 32 bits, then widened */
        k = 0;
        while ((uint64_t)(uint32_t)(plength << k) < sum) {
            if (k < max_rice) k++;
            else break;
        }
        /* flac.c:1487-1501 */
        if (k > 0)
            total += 4 + (sum >> (k - 1)) + (uint64_t)((1 + k) * plength) - (plength / 2);
        else
            total += 4 + (sum << 1) + plength - (plength / 2);
        rice[p] = (uint8_t)k;
        if (part_len) part_len[p] = take;
    }
    return total;
}
/* note on `(1 + k) * plength`: in the reference this product is `unsigned`
 * (both operands unsigned int) and wraps modulo 2^32 before it is added to the
 * uint64_t total; `(uint64_t)((1 + k) * plength)` above keeps that. */

/* flac.c:1326-1435 flacenc_encode_residuals */
static void encode_residuals(bitvec *bs, const orc_encoder *e, unsigned block_size,
                             unsigned predictor_order, const int32_t *residuals,
                             orc_subframe_decision *dec)
{
    unsigned n_res = block_size - predictor_order;
    unsigned po, best_po = 0, p, pos;
    uint64_t best_total = ULLONG_MAX;
    size_t maxparts = (size_t)1 << (e->opt.max_residual_partition_order > 15
                                        ? 15 : e->opt.max_residual_partition_order);
    uint8_t *rice = (uint8_t *)malloc(maxparts);
    uint8_t *best_rice = (uint8_t *)malloc(maxparts);
    unsigned *plen = (unsigned *)malloc(sizeof(unsigned) * maxparts);
    unsigned *best_plen = (unsigned *)malloc(sizeof(unsigned) * maxparts);
    unsigned coding_method, maxk = 0;

    for (po = 0; po <= e->opt.max_residual_partition_order && po <= 15; po++) {
        if ((block_size % (1u << po)) == 0) {
            uint64_t total = orc_residual_partitions(residuals, n_res, block_size, predictor_order,
                                                     po, e->max_rice_parameter, rice, plen);
            if (total < best_total) { /* strict: first minimum wins (flac.c:1378) */
                uint8_t *t = rice; unsigned *tl = plen;
                rice = best_rice; best_rice = t;
                plen = best_plen; best_plen = tl;
                best_po = po;
                best_total = total;
            }
        } else {
            break; /* flac.c:1389-1393 */
        }
    }

    for (p = 0; p < (1u << best_po); p++)
        if (best_rice[p] > maxk) maxk = best_rice[p];
    coding_method = (maxk > 14) ? 1 : 0; /* flac.c:1399-1402 */

    bv_put(bs, 2, coding_method);
    bv_put(bs, 4, best_po);
    pos = 0;
    for (p = 0; p < (1u << best_po); p++) {
        unsigned k = best_rice[p], i;
        bv_put(bs, coding_method ? 5 : 4, k);
        for (i = 0; i < best_plen[p]; i++) {
            int32_t r = residuals[pos + i];
            uint32_t u, msb, lsb;
            /* flac.c:1424-1432 */
            if (r >= 0) u = (uint32_t)r << 1;
            else u = (((uint32_t)(-(int64_t)r) - 1u) << 1) | 1u;
            msb = u >> k;
            lsb = u - (msb << k);
            bv_put_unary(bs, 1, msb);
            bv_put(bs, k, lsb);
        }
        pos += best_plen[p];
    }

    if (dec) {
        dec->coding_method = coding_method;
        dec->partition_order = best_po;
        dec->n_partitions = 1u << best_po;
        free(dec->rice);
        dec->rice = (uint8_t *)malloc(dec->n_partitions);
        memcpy(dec->rice, best_rice, dec->n_partitions);
    }
    free(rice); free(best_rice); free(plen); free(best_plen);
}

/* subframe header: flac.c:820-826 / 897-905 / 977-985 */
static void put_subframe_header(bitvec *bs, unsigned type_bits, unsigned wasted)
{
    bv_put(bs, 1, 0);
    bv_put(bs, 6, type_bits);
    if (wasted) {
        bv_put(bs, 1, 1);
        bv_put_unary(bs, 1, wasted - 1);
    } else {
        bv_put(bs, 1, 0);
    }
}

/* flac.c:872-895: errors of orders 0..4 over the common tail; first strict min */
unsigned orc_best_fixed_order(const int32_t *s, unsigned n, uint64_t err[5])
{
    unsigned best = 0, k, i;
    uint64_t sums[5] = {0, 0, 0, 0, 0};
    /* order 0, truncated by 4 (de_head clamps: empty when n <= 4) */
    for (i = 4; i < n; i++) sums[0] += (uint64_t)(int64_t)abs(s[i]);
    if (n > 4) {
        int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * n);
        unsigned len = n;
        memcpy(cur, s, sizeof(int32_t) * n);
        for (k = 1; k <= 4; k++) {
            /* flac.c:918-930 next order = successive differences (in place) */
            for (i = 0; i + 1 < len; i++) cur[i] = (int32_t)((uint32_t)cur[i + 1] - (uint32_t)cur[i]);
            len--;
            /* drop the first 4-k so every order is summed over the last n-4 */
            for (i = 4 - k; i < len; i++) sums[k] += (uint64_t)(int64_t)abs(cur[i]);
        }
        free(cur);
        {
            uint64_t best_sum = sums[0];
            for (k = 1; k <= 4; k++)
                if (sums[k] < best_sum) { best_sum = sums[k]; best = k; }
        }
    }
    if (err) memcpy(err, sums, sizeof(sums));
    return best;
}

/* flac.c:856-916 flacenc_write_fixed_subframe */
static void write_fixed_subframe(bitvec *bs, orc_encoder *e, unsigned bps, unsigned wasted,
                                 const int32_t *s, unsigned n, orc_subframe_decision *dec)
{
    unsigned order = orc_best_fixed_order(s, n, NULL);
    int32_t *res = e->scratch[5];
    unsigned i, k, len = n;
    memcpy(res, s, sizeof(int32_t) * n);
    for (k = 0; k < order; k++) {
        for (i = 0; i + 1 < len; i++) res[i] = (int32_t)((uint32_t)res[i + 1] - (uint32_t)res[i]);
        len--;
    }
    put_subframe_header(bs, 0x8 | order, wasted); /* 1 bit pad, 3 bits "001", 3 bits order */
    for (i = 0; i < order; i++) bv_put_signed(bs, bps - wasted, s[i]);
    if (dec) { dec->type = ORC_FIXED; dec->order = order; }
    encode_residuals(bs, e, n, order, res, dec);
}

/* flac.c:1139-1161 Tukey window, alpha = 0.5 */
void orc_tukey_window(unsigned N, double *w)
{
    const double alpha = 0.5;
    const unsigned window1 = (unsigned)(alpha * (N - 1)) / 2;
    const unsigned window2 = (unsigned)((N - 1) * (1.0 - (alpha / 2.0)));
    unsigned n;
    for (n = 0; n < N; n++) {
        if (n <= window1)
            w[n] = 0.5 * (1.0 + cos(M_PI * (((2 * n) / (alpha * (N - 1))) - 1.0)));
        else if (n <= window2)
            w[n] = 1.0;
        else
            w[n] = 0.5 * (1.0 + cos(M_PI * (((2.0 * n) / (alpha * (N - 1))) - (2.0 / alpha) + 1.0)));
    }
}

/* flac.c:1169-1188: sequential sums, one rounding per multiply and per add */
void orc_autocorrelate(unsigned max_lpc_order, const double *x, unsigned n, double *autoc)
{
    unsigned lag, i;
    for (lag = 0; lag <= max_lpc_order; lag++) {
        double acc = 0.0; /* built with -ffp-contract=off: mul and add round separately */
        for (i = 0; i < n - lag; i++) acc += x[i] * x[i + lag];
        autoc[lag] = acc;
    }
}

/* flac.c:1190-1231 Levinson-Durbin */
void orc_lp_coefficients(unsigned max_lpc_order, const double *r, double *lp, double *err)
{
    unsigned i, j;
    double k, q, t;
    if (max_lpc_order == 0) return;
    k = r[1] / r[0];
    lp[0] = k;
    t = k * k; t = 1.0 - t;
    err[0] = r[0] * t;
    for (i = 1; i < max_lpc_order; i++) {
        const double *prev = lp + (size_t)(i - 1) * ORC_MAX_LPC_ORDER;
        double *cur = lp + (size_t)i * ORC_MAX_LPC_ORDER;
        q = r[i + 1];
        for (j = 0; j < i; j++) { t = prev[j] * r[i - j]; q = q - t; }
        k = q / err[i - 1];
        for (j = 0; j < i; j++) { t = k * prev[i - j - 1]; cur[j] = prev[j] - t; }
        cur[i] = k;
        t = k * k; t = 1.0 - t;
        err[i] = err[i - 1] * t;
    }
}

/* flac.c:1233-1268 */
unsigned orc_estimate_best_lpc_order(unsigned bps, unsigned precision, unsigned max_lpc_order,
                                     unsigned block_size, const double *lp_error)
{
    const double error_scale = (M_LN2 * M_LN2) / ((double)block_size * 2.0);
    unsigned best_order = 0, i;
    double best_bits = DBL_MAX;
    for (i = 0; i < max_lpc_order; i++) {
        const unsigned order = i + 1;
        if (lp_error[i] > 0.0) {
            const unsigned header_bits = order * (bps + precision);
            double bpr = log(lp_error[i] * error_scale) / (M_LN2 * 2);
            double est;
            if (!(bpr > 0.0)) bpr = 0.0; /* MAX(x, 0.0): NaN -> 0.0 as `(x) > (y) ? x : y` */
            est = header_bits + bpr * (block_size - order);
            if (est < best_bits) { best_order = order; best_bits = est; }
        } else {
            return order; /* flac.c:1261-1262: also taken for NaN */
        }
    }
    return best_order;
}

/* (int)round(x) as x86-64 cvttsd2si does it: NaN and out-of-range give INT_MIN */
static int x86_double_to_int(double x)
{
    if (!(x > -2147483649.0 && x < 2147483648.0)) return INT_MIN;
    return (int)x;
}

/* flac.c:1270-1324 */
void orc_quantize_coefficients(const double *c, unsigned order, unsigned precision, int *qlp,
                               int *shift_out)
{
    double l = DBL_MIN;
    int log2cmax, shift, qmax, qmin;
    double error = 0.0, t;
    unsigned i;
    for (i = 0; i < order; i++) {
        double a = fabs(c[i]);
        l = (a > l) ? a : l; /* MAX(fabs(c), l): NaN leaves l unchanged */
    }
    frexp(l, &log2cmax);
    shift = (int)(precision - 1) - (log2cmax - 1) - 1;
    if (shift < -(1 << 4)) shift = -(1 << 4);
    if (shift > (1 << 4) - 1) shift = (1 << 4) - 1;
    qmax = (1 << (precision - 1)) - 1;
    qmin = -(1 << (precision - 1));
    if (shift >= 0) {
        for (i = 0; i < order; i++) {
            int ei;
            t = c[i] * (double)(1 << shift);
            error = error + t;
            ei = x86_double_to_int(round(error));
            qlp[i] = ei < qmin ? qmin : (ei > qmax ? qmax : ei);
            error = error - (double)ei;
        }
    } else {
        for (i = 0; i < order; i++) {
            int ei;
            t = c[i] / (double)(1 << -shift);
            error = error + t;
            ei = x86_double_to_int(round(error));
            qlp[i] = ei < qmin ? qmin : (ei > qmax ? qmax : ei);
            error = error - (double)ei;
        }
        shift = 0;
    }
    *shift_out = shift;
}

/* flac.c:961-1016 flacenc_encode_lpc_subframe */
static void encode_lpc_subframe(bitvec *bs, orc_encoder *e, unsigned bps, unsigned wasted,
                                unsigned precision, int shift, const int *qlp, unsigned order,
                                const int32_t *s, unsigned n, orc_subframe_decision *dec)
{
    int32_t *res = e->scratch[6];
    unsigned i, j;
    put_subframe_header(bs, 0x20 | (order - 1), wasted);
    for (i = 0; i < order; i++) bv_put_signed(bs, bps - wasted, s[i]);
    bv_put(bs, 4, precision - 1);
    bv_put_signed(bs, 5, shift);
    for (i = 0; i < order; i++) bv_put_signed(bs, precision, qlp[i]);
    for (i = 0; i < n - order; i++) {
        int64_t acc = 0;
        for (j = 0; j < order; j++) acc += (int64_t)qlp[j] * (int64_t)s[i + order - j - 1];
        acc >>= shift;
        /* flac.c:1007: truncating int64 -> int cast, then a wrapping int subtract */
        res[i] = (int32_t)((uint32_t)s[i + order] - (uint32_t)(int32_t)acc);
    }
    if (dec) {
        dec->type = ORC_LPC; dec->order = order; dec->precision = precision; dec->shift = shift;
        memset(dec->coeffs, 0, sizeof(dec->coeffs));
        memcpy(dec->coeffs, qlp, sizeof(int) * order);
    }
    encode_residuals(bs, e, n, order, res, dec);
}

/* flac.c:1018-1127 flacenc_best_lpc_coefficients + 932-959 write_lpc_subframe */
static void write_lpc_subframe(bitvec *bs, orc_encoder *e, unsigned bps, unsigned wasted,
                               const int32_t *s, unsigned n, orc_subframe_decision *dec)
{
    const unsigned max_order = e->opt.max_lpc_order;
    int qlp[ORC_MAX_LPC_ORDER];
    unsigned order, precision;
    int shift;

    if (n > max_order + 1) {
        double autoc[ORC_MAX_LPC_ORDER + 1];
        double *lp = (double *)calloc((size_t)ORC_MAX_LPC_ORDER * ORC_MAX_LPC_ORDER, sizeof(double));
        double lp_error[ORC_MAX_LPC_ORDER];
        unsigned i;
        if (e->window_len != n) { /* flac.c:1139 cached per length */
            e->window = (double *)realloc(e->window, sizeof(double) * n);
            orc_tukey_window(n, e->window);
            e->window_len = n;
        }
        for (i = 0; i < n; i++) e->windowed[i] = s[i] * e->window[i]; /* flac.c:1164-1166 */
        orc_autocorrelate(max_order, e->windowed, n, autoc);
        orc_lp_coefficients(max_order, autoc, lp, lp_error);
        precision = e->qlp_coeff_precision;
        if (!e->opt.exhaustive_model_search) {
            /* flac.c:1055-1069: note bps here is the subframe bps BEFORE removing wasted bits */
            order = orc_estimate_best_lpc_order(bps, precision, max_order, n, lp_error);
            orc_quantize_coefficients(lp + (size_t)(order - 1) * ORC_MAX_LPC_ORDER, order, precision,
                                      qlp, &shift);
        } else {
            /* flac.c:1070-1120: full encode of every order into a bit counter */
            uint64_t best_bits = UINT_MAX; /* unsigned best_bits = UINT_MAX */
            unsigned o, best_order = 0;
            int best_shift = 0, best_qlp[ORC_MAX_LPC_ORDER];
            bitvec tmp;
            bv_init(&tmp);
            for (o = 1; o <= max_order; o++) {
                int cq[ORC_MAX_LPC_ORDER], cs;
                unsigned cbits;
                orc_quantize_coefficients(lp + (size_t)(o - 1) * ORC_MAX_LPC_ORDER, o, precision, cq, &cs);
                bv_reset(&tmp);
                encode_lpc_subframe(&tmp, e, bps, wasted, precision, cs, cq, o, s, n, NULL);
                cbits = (unsigned)tmp.nbits; /* accumulator is `unsigned` */
                if (cbits < best_bits) {
                    best_bits = cbits; best_order = o; best_shift = cs;
                    memcpy(best_qlp, cq, sizeof(int) * o);
                }
            }
            bv_free(&tmp);
            order = best_order; shift = best_shift;
            memcpy(qlp, best_qlp, sizeof(int) * order);
        }
        free(lp);
    } else {
        /* flac.c:1121-1126 dummy coefficients */
        order = 1; qlp[0] = 1; precision = 2; shift = 0;
    }
    encode_lpc_subframe(bs, e, bps, wasted, precision, shift, qlp, order, s, n, dec);
}

/* flac.c:832-854 */
static void write_verbatim_subframe(bitvec *bs, unsigned bps, unsigned wasted, const int32_t *s,
                                    unsigned n)
{
    unsigned i;
    put_subframe_header(bs, 1, wasted);
    for (i = 0; i < n; i++) bv_put_signed(bs, bps - wasted, s[i]);
}

/* flac.c:673-811 flacenc_write_subframe */
static void write_subframe(bitvec *bs, orc_encoder *e, unsigned bps, const int32_t *samples,
                           unsigned n, orc_subframe_decision *dec)
{
    const int try_VERBATIM = !e->opt.no_verbatim_subframes;
    const int try_CONSTANT = !e->opt.no_constant_subframes;
    const int try_FIXED = !e->opt.no_fixed_subframes;
    const int try_LPC = !(e->opt.no_lpc_subframes || (e->opt.max_lpc_order == 0));
    uint64_t start = bs->nbits;
    orc_subframe_decision dfix, dlpc;
    memset(&dfix, 0, sizeof(dfix));
    memset(&dlpc, 0, sizeof(dlpc));
    if (dec) { decision_clear(dec); dec->bits_per_sample = bps; }

    if (try_CONSTANT && all_identical(samples, n)) {
        /* flac.c:691-693,813-830: written with wasted = 0 */
        put_subframe_header(bs, 0, 0);
        bv_put_signed(bs, bps, samples[0]);
        if (dec) dec->type = ORC_CONSTANT;
    } else {
        unsigned wasted = orc_wasted_bits(samples, n), i;
        int32_t *s = e->scratch[4];
        unsigned verbatim_bits = INT_MAX, fixed_bits = 0, lpc_bits = 0;
        int choice; /* ORC_FIXED / ORC_LPC / ORC_VERBATIM */
        for (i = 0; i < n; i++) s[i] = samples[i] >> wasted; /* flac.c:700-705 */

        if (try_FIXED) {
            bv_reset(&e->fixed_bits);
            write_fixed_subframe(&e->fixed_bits, e, bps, wasted, s, n, &dfix);
            fixed_bits = (unsigned)e->fixed_bits.nbits;
        }
        if (try_LPC) {
            bv_reset(&e->lpc_bits);
            write_lpc_subframe(&e->lpc_bits, e, bps, wasted, s, n, &dlpc);
            lpc_bits = (unsigned)e->lpc_bits.nbits;
        }
        if (try_VERBATIM) verbatim_bits = (bps - wasted) * n; /* flac.c:727-730: no header */

        /* flac.c:732-809 */
        if (try_FIXED && try_LPC && try_VERBATIM) {
            unsigned m = lpc_bits < verbatim_bits ? lpc_bits : verbatim_bits;
            if (fixed_bits < m) choice = ORC_FIXED;
            else if (lpc_bits < verbatim_bits) choice = ORC_LPC;
            else choice = ORC_VERBATIM;
        } else if (!try_FIXED && !try_LPC && !try_VERBATIM) choice = ORC_VERBATIM;
        else if (try_FIXED && !try_LPC && !try_VERBATIM) choice = ORC_FIXED;
        else if (!try_FIXED && try_LPC && !try_VERBATIM) choice = ORC_LPC;
        else if (try_FIXED && try_LPC && !try_VERBATIM) choice = (fixed_bits < lpc_bits) ? ORC_FIXED : ORC_LPC;
        else if (!try_FIXED && !try_LPC && try_VERBATIM) choice = ORC_VERBATIM;
        else if (try_FIXED && !try_LPC && try_VERBATIM) choice = (fixed_bits < verbatim_bits) ? ORC_FIXED : ORC_VERBATIM;
        else choice = (lpc_bits < verbatim_bits) ? ORC_LPC : ORC_VERBATIM;

        if (choice == ORC_FIXED) {
            bv_append(bs, &e->fixed_bits);
            if (dec) { decision_copy(dec, &dfix); }
        } else if (choice == ORC_LPC) {
            bv_append(bs, &e->lpc_bits);
            if (dec) { decision_copy(dec, &dlpc); }
        } else {
            write_verbatim_subframe(bs, bps, wasted, s, n);
            if (dec) dec->type = ORC_VERBATIM;
        }
        if (dec) { dec->wasted = wasted; dec->bits_per_sample = bps; }
    }
    if (dec) dec->bits = bs->nbits - start;
    decision_clear(&dfix);
    decision_clear(&dlpc);
}

/* flac.c:1531-1566 */
static void put_utf8(bitvec *bs, unsigned value)
{
    if (value <= 0x7F) {
        bv_put(bs, 8, value);
    } else {
        unsigned total_bytes = 0;
        int shift;
        if (value <= 0x7FF) total_bytes = 2;
        else if (value <= 0xFFFF) total_bytes = 3;
        else if (value <= 0x1FFFFF) total_bytes = 4;
        else if (value <= 0x3FFFFFF) total_bytes = 5;
        else if (value <= 0x7FFFFFFF) total_bytes = 6;
        shift = (int)(total_bytes - 1) * 6;
        bv_put_unary(bs, 0, total_bytes);
        bv_put(bs, 7 - total_bytes, value >> shift);
        for (shift -= 6; shift >= 0; shift -= 6) {
            bv_put_unary(bs, 0, 1);
            bv_put(bs, 6, (value >> shift) & 0x3F);
        }
    }
}

/* flac.c:412-518 */
static void write_frame_header(bitvec *bs, const orc_encoder *e, unsigned block_size,
                               unsigned channel_assignment, unsigned frame_number)
{
    unsigned bsb, srb, bpsb;
    size_t start = (size_t)(bs->nbits >> 3);
    const unsigned sr = e->sample_rate;
    switch (block_size) {
    case 192: bsb = 0x1; break;   case 576: bsb = 0x2; break;
    case 1152: bsb = 0x3; break;  case 2304: bsb = 0x4; break;
    case 4608: bsb = 0x5; break;  case 256: bsb = 0x8; break;
    case 512: bsb = 0x9; break;   case 1024: bsb = 0xA; break;
    case 2048: bsb = 0xB; break;  case 4096: bsb = 0xC; break;
    case 8192: bsb = 0xD; break;  case 16384: bsb = 0xE; break;
    case 32768: bsb = 0xF; break;
    default:
        if (block_size < (0xFF + 1)) bsb = 0x6;
        else if (block_size < (0xFFFF + 1)) bsb = 0x7;
        else bsb = 0x0;
    }
    switch (sr) {
    case 88200: srb = 0x1; break;  case 176400: srb = 0x2; break;
    case 192000: srb = 0x3; break; case 8000: srb = 0x4; break;
    case 16000: srb = 0x5; break;  case 22050: srb = 0x6; break;
    case 24000: srb = 0x7; break;  case 32000: srb = 0x8; break;
    case 44100: srb = 0x9; break;  case 48000: srb = 0xA; break;
    case 96000: srb = 0xB; break;
    default:
        if ((sr <= 255000) && ((sr % 1000) == 0)) srb = 0xC;
        else if ((sr <= 655350) && ((sr % 10) == 0)) srb = 0xE;
        else if (sr <= 0xFFFF) srb = 0xD;
        else srb = 0x0;
    }
    switch (e->bits_per_sample) {
    case 8: bpsb = 0x1; break;  case 12: bpsb = 0x2; break;
    case 16: bpsb = 0x4; break; case 20: bpsb = 0x5; break;
    case 24: bpsb = 0x6; break; default: bpsb = 0x0;
    }
    bv_put(bs, 14, 0x3FFE);
    bv_put(bs, 1, 0);
    bv_put(bs, 1, 0);
    bv_put(bs, 4, bsb);
    bv_put(bs, 4, srb);
    bv_put(bs, 4, channel_assignment);
    bv_put(bs, 3, bpsb);
    bv_put(bs, 1, 0);
    put_utf8(bs, frame_number);
    if (bsb == 0x6) bv_put(bs, 8, (block_size - 1) & 0xFF);
    else if (bsb == 0x7) bv_put(bs, 16, (block_size - 1) & 0xFFFF);
    if (srb == 0xC) bv_put(bs, 8, sr / 1000);
    else if (srb == 0xD) bv_put(bs, 16, sr);
    else if (srb == 0xE) bv_put(bs, 16, sr / 10);
    bv_put(bs, 8, orc_crc8(bs->buf + start, (size_t)(bs->nbits >> 3) - start));
}

size_t orc_frame_bound(const orc_encoder *e, unsigned n)
{
    /* generous: every subframe verbatim at bps+1 plus headers, LPC params and slack.
       With VERBATIM disabled a FIXED/LPC subframe can exceed this; the bitvec grows
       on demand, only the caller's buffer needs care (tests pass 4x this). */
    return 64 + (size_t)e->channels * (((size_t)n * (e->bits_per_sample + 1) + 2048) / 8 + 8);
}

/* flac.c:520-671 flacenc_write_frame */
size_t orc_encode_frame(orc_encoder *e, const int32_t *const *samples, unsigned n,
                        unsigned frame_number, uint8_t *out, orc_frame_decision *decision)
{
    bitvec *bs = &e->frame_bits;
    const unsigned bps = e->bits_per_sample;
    unsigned c;
    size_t nbytes;
    uint16_t crc;

    ensure_scratch(e, n);
    bv_reset(bs);
    if (decision) orc_frame_decision_free(decision);

    if ((e->channels == 2) && (e->opt.mid_side || e->opt.adaptive_mid_side)) {
        int32_t *avg = e->scratch[0], *dif = e->scratch[1];
        orc_subframe_decision d[4];
        unsigned lb, rb, ab, db, assignment, i;
        int first, second; /* candidate indices: 0 L, 1 R, 2 avg, 3 diff */
        memset(d, 0, sizeof(d));
        /* flac.c:1507-1529 */
        for (i = 0; i < n; i++) {
            avg[i] = (samples[0][i] + samples[1][i]) >> 1;
            dif[i] = samples[0][i] - samples[1][i];
        }
        for (i = 0; i < 4; i++) bv_reset(&e->cand_bits[i]);
        write_subframe(&e->cand_bits[0], e, bps, samples[0], n, &d[0]);
        write_subframe(&e->cand_bits[1], e, bps, samples[1], n, &d[1]);
        write_subframe(&e->cand_bits[2], e, bps, avg, n, &d[2]);
        write_subframe(&e->cand_bits[3], e, bps + 1, dif, n, &d[3]);
        lb = (unsigned)e->cand_bits[0].nbits; rb = (unsigned)e->cand_bits[1].nbits;
        ab = (unsigned)e->cand_bits[2].nbits; db = (unsigned)e->cand_bits[3].nbits;
#define UMIN(a, b) ((a) < (b) ? (a) : (b))
        if (e->opt.mid_side) { /* flac.c:581-629 */
            if ((lb + rb) < UMIN(UMIN(lb + db, db + rb), ab + db)) { assignment = 0x1; first = 0; second = 1; }
            else if (lb < UMIN(rb, ab)) { assignment = 0x8; first = 0; second = 3; }
            else if (rb < ab) { assignment = 0x9; first = 3; second = 1; }
            else { assignment = 0xA; first = 2; second = 3; }
        } else if ((lb + rb) < (ab + db)) { assignment = 0x1; first = 0; second = 1; } /* flac.c:630-641 */
        else { assignment = 0xA; first = 2; second = 3; }
#undef UMIN
        write_frame_header(bs, e, n, assignment, frame_number);
        bv_append(bs, &e->cand_bits[first]);
        bv_append(bs, &e->cand_bits[second]);
        if (decision) {
            decision->channel_assignment = assignment;
            decision->n_subframes = 2;
            decision->have_candidates = 1;
            for (i = 0; i < 4; i++) decision_copy(&decision->cand[i], &d[i]);
            decision_copy(&decision->sub[0], &d[first]);
            decision_copy(&decision->sub[1], &d[second]);
        }
        for (i = 0; i < 4; i++) decision_clear(&d[i]);
    } else {
        /* flac.c:653-666 */
        write_frame_header(bs, e, n, e->channels - 1, frame_number);
        for (c = 0; c < e->channels; c++)
            write_subframe(bs, e, bps, samples[c], n, decision ? &decision->sub[c] : NULL);
        if (decision) {
            decision->channel_assignment = e->channels - 1;
            decision->n_subframes = e->channels;
            decision->have_candidates = 0;
        }
    }
    bv_align(bs);
    nbytes = (size_t)(bs->nbits >> 3);
    crc = orc_crc16(bs->buf, nbytes);
    bv_put(bs, 16, crc);
    nbytes += 2;
    memcpy(out, bs->buf, nbytes);
    return nbytes;
}

/* flac.c:376-409 */
static void put_streaminfo(bitvec *bs, unsigned min_block, unsigned max_block, unsigned min_frame,
                           unsigned max_frame, unsigned sample_rate, unsigned channels,
                           unsigned bps, uint64_t total_samples, const uint8_t md5[16])
{
    int i;
#define CLAMPU(v, hi) ((v) > (hi) ? (hi) : (v))
    bv_put(bs, 16, CLAMPU(min_block, 0xFFFFu));
    bv_put(bs, 16, CLAMPU(max_block, 0xFFFFu));
    bv_put(bs, 24, CLAMPU(min_frame, 0xFFFFFFu));
    bv_put(bs, 24, CLAMPU(max_frame, 0xFFFFFFu));
    bv_put(bs, 20, CLAMPU(sample_rate, 0xFFFFFu));
    bv_put(bs, 3, CLAMPU(channels - 1, 7u));
    bv_put(bs, 5, CLAMPU(bps - 1, 31u));
#undef CLAMPU
    bv_put(bs, 36, total_samples);
    for (i = 0; i < 16; i++) bv_put(bs, 8, md5[i]);
}

static int32_t load_le_signed(const uint8_t *p, unsigned bytes)
{
    uint32_t v = 0;
    unsigned i;
    for (i = 0; i < bytes; i++) v |= (uint32_t)p[i] << (8 * i);
    if (bytes < 4 && (v & (1u << (8 * bytes - 1)))) v |= ~((1u << (8 * bytes)) - 1);
    return (int32_t)v;
}

/* flac.c:124-306 encoders_encode_flac (standalone signature) */
int orc_encode_stream(const orc_options *opt, unsigned sample_rate, unsigned channels,
                      unsigned bps, const uint8_t *pcm, size_t pcm_bytes, uint8_t **file_out,
                      size_t *file_len, uint64_t **frame_offsets, uint32_t **frame_lengths,
                      size_t *n_frames_out)
{
    orc_encoder *e = orc_encoder_new(opt, sample_rate, channels, bps);
    const unsigned bytes_ps = bps / 8;
    const size_t frame_bytes = (size_t)channels * bytes_ps;
    const uint64_t total_frames = pcm_bytes / frame_bytes; /* partial frames dropped, pcmconv.c:398-400 */
    const unsigned block = opt->block_size;
    const char *ver = opt->version ? opt->version : "2.22alpha1";
    char vendor[0x200];
    bitvec file;
    uint8_t md5[16] = {0};
    unsigned min_frame = 0xFFFFFF, max_frame = 0, frame_number = 0, c;
    uint64_t total_samples = 0, pos = 0, current_offset = 0;
    size_t nfr = 0, cap = 0;
    uint64_t *offs = NULL;
    uint32_t *lens = NULL;
    int32_t *chan[ORC_MAX_CHANNELS];
    const int32_t *chanp[ORC_MAX_CHANNELS];
    uint8_t *fbuf;
    size_t streaminfo_pos;
    orc_md5_ctx md5c;

    for (c = 0; c < channels; c++) chan[c] = (int32_t *)malloc(sizeof(int32_t) * (block + 1));
    fbuf = (uint8_t *)malloc(orc_frame_bound(e, block) * 4 + 65536);
    snprintf(vendor, sizeof(vendor), "Python Audio Tools %s", ver);
    bv_init(&file);
    md5_init(&md5c);

    /* flac.c:209-238 */
    bv_put(&file, 32, 0x664C6143);
    bv_put(&file, 1, 0); bv_put(&file, 7, 0); bv_put(&file, 24, 34);
    streaminfo_pos = (size_t)(file.nbits >> 3);
    put_streaminfo(&file, block, block, min_frame, max_frame, sample_rate, channels, bps, 0, md5);
    bv_put(&file, 1, 0); bv_put(&file, 7, 4);
    bv_put(&file, 24, (unsigned)(4 + strlen(vendor) + 4));
    {   /* little-endian fields */
        unsigned L = (unsigned)strlen(vendor), i;
        for (i = 0; i < 4; i++) bv_put(&file, 8, (L >> (8 * i)) & 0xFF);
        for (i = 0; i < L; i++) bv_put(&file, 8, (uint8_t)vendor[i]);
        bv_put(&file, 32, 0);
    }
    bv_put(&file, 1, 1); bv_put(&file, 7, 1); bv_put(&file, 24, opt->padding_size);
    {
        unsigned i;
        bv_reserve(&file, (uint64_t)opt->padding_size * 8);
        for (i = 0; i < opt->padding_size; i++) bv_put(&file, 8, 0);
    }

    /* flac.c:247-274 frame loop */
    while (pos < total_frames) {
        unsigned n = (unsigned)((total_frames - pos) < block ? (total_frames - pos) : block);
        unsigned i;
        size_t flen;
        for (i = 0; i < n; i++)
            for (c = 0; c < channels; c++)
                chan[c][i] = load_le_signed(pcm + ((pos + i) * channels + c) * bytes_ps, bytes_ps);
        for (c = 0; c < channels; c++) chanp[c] = chan[c];
        md5_update(&md5c, pcm + pos * frame_bytes, (size_t)n * frame_bytes);
        if (nfr == cap) {
            cap = cap ? cap * 2 : 1024;
            offs = (uint64_t *)realloc(offs, cap * sizeof(uint64_t));
            lens = (uint32_t *)realloc(lens, cap * sizeof(uint32_t));
        }
        offs[nfr] = current_offset; lens[nfr] = n; nfr++;
        flen = orc_encode_frame(e, chanp, n, frame_number++, fbuf, NULL);
        total_samples += n;
        if (flen < min_frame) min_frame = (unsigned)flen;
        if (flen > max_frame) max_frame = (unsigned)flen;
        current_offset += flen;
        bv_reserve(&file, (uint64_t)flen * 8);
        memcpy(file.buf + (file.nbits >> 3), fbuf, flen);
        file.nbits += (uint64_t)flen * 8;
        pos += n;
    }

    /* flac.c:277-279 rewrite STREAMINFO */
    md5_final(&md5c, md5);
    {
        bitvec si;
        bv_init(&si);
        put_streaminfo(&si, block, block, min_frame, max_frame, sample_rate, channels, bps,
                       total_samples, md5);
        memcpy(file.buf + streaminfo_pos, si.buf, 34);
        bv_free(&si);
    }

    *file_out = file.buf;
    *file_len = (size_t)(file.nbits >> 3);
    if (frame_offsets) *frame_offsets = offs; else free(offs);
    if (frame_lengths) *frame_lengths = lens; else free(lens);
    if (n_frames_out) *n_frames_out = nfr;
    for (c = 0; c < channels; c++) free(chan[c]);
    free(fbuf);
    orc_encoder_free(e);
    return 1;
}

/* frames of a contiguous PCM range only (no stream head): what one shard of a frame-range
 * sharded encode produces.  Frame numbers start at first_frame_number. */
int orc_encode_range(const orc_options *opt, unsigned sample_rate, unsigned channels, unsigned bps,
                     const uint8_t *pcm, size_t pcm_bytes, unsigned first_frame_number,
                     uint8_t **frames_out, size_t *frames_len, uint32_t **frame_bytes_out,
                     size_t *n_frames_out)
{
    orc_encoder *e = orc_encoder_new(opt, sample_rate, channels, bps);
    const unsigned bytes_ps = bps / 8;
    const size_t fb = (size_t)channels * bytes_ps;
    const uint64_t total = pcm_bytes / fb;
    const unsigned block = opt->block_size;
    int32_t *chan[ORC_MAX_CHANNELS];
    const int32_t *chanp[ORC_MAX_CHANNELS];
    uint8_t *fbuf = (uint8_t *)malloc(orc_frame_bound(e, block) * 4 + 65536);
    uint8_t *out = NULL;
    size_t out_len = 0, out_cap = 0, nfr = 0, cap = 0;
    uint32_t *sizes = NULL;
    uint64_t pos = 0;
    unsigned c, fn = first_frame_number;
    for (c = 0; c < channels; c++) chan[c] = (int32_t *)malloc(sizeof(int32_t) * (block + 1));
    while (pos < total) {
        unsigned n = (unsigned)((total - pos) < block ? (total - pos) : block), i;
        size_t flen;
        for (i = 0; i < n; i++)
            for (c = 0; c < channels; c++)
                chan[c][i] = load_le_signed(pcm + ((pos + i) * channels + c) * bytes_ps, bytes_ps);
        for (c = 0; c < channels; c++) chanp[c] = chan[c];
        flen = orc_encode_frame(e, chanp, n, fn++, fbuf, NULL);
        if (out_len + flen > out_cap) { out_cap = (out_cap + flen) * 2; out = (uint8_t *)realloc(out, out_cap); }
        memcpy(out + out_len, fbuf, flen);
        out_len += flen;
        if (nfr == cap) { cap = cap ? cap * 2 : 256; sizes = (uint32_t *)realloc(sizes, cap * sizeof(uint32_t)); }
        sizes[nfr++] = (uint32_t)flen;
        pos += n;
    }
    for (c = 0; c < channels; c++) free(chan[c]);
    free(fbuf);
    orc_encoder_free(e);
    *frames_out = out; *frames_len = out_len;
    if (frame_bytes_out) *frame_bytes_out = sizes; else free(sizes);
    if (n_frames_out) *n_frames_out = nfr;
    return 1;
}

void orc_free(void *p) { free(p); }

/* ------------------------------------------------------------------ */
/* deterministic integer-only synthetic PCM              */
/* The same formulas are implemented on the device in                   */
/* python-audio-tools_b200/csrc/synth.cuh; tests compare the two.       */
/* ------------------------------------------------------------------ */
static uint64_t orc_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

/* parabolic "sine": phase in [0,2^32) -> [-16384, 16384] */
static int32_t orc_psin(uint32_t phase)
{
    int32_t x = (int32_t)(phase >> 16) - 32768;
    int32_t ax = x < 0 ? -x : x;
    return (x * (32768 - ax)) >> 14;
}

static const int32_t synth_amp[4] = {12000, 8000, 5000, 3000};
static const int32_t synth_share[8] = {3, 3, 4, 1, 1, 1, 2, 2}; /* quarters of the shared voices */

static int32_t synth_sample(uint64_t seed, unsigned c, unsigned bps, uint64_t t)
{
    int64_t shared = 0, own = 0, tot;
    int32_t env, noise;
    int64_t v, lo, hi;
    unsigned vix;
    const unsigned lfe = (c == 3);
    for (vix = 0; vix < 4; vix++) {
        uint32_t inc = 5000000u + (uint32_t)(orc_mix64(seed * 16 + vix) % 400000000u);
        uint32_t ph0 = (uint32_t)orc_mix64(seed * 16 + 8 + vix);
        if (lfe) inc >>= 5;
        shared += (int64_t)synth_amp[vix] * orc_psin((uint32_t)(inc * (uint32_t)t + ph0));
    }
    for (vix = 0; vix < 2; vix++) {
        uint32_t inc = 4000000u + (uint32_t)(orc_mix64(seed * 16 + 64 + c * 4 + vix) % 300000000u);
        uint32_t ph0 = (uint32_t)orc_mix64(seed * 16 + 128 + c * 4 + vix);
        if (lfe) inc >>= 5;
        own += (int64_t)synth_amp[vix] * 2 * orc_psin((uint32_t)(inc * (uint32_t)t + ph0));
    }
    /* >> 14 brings amp*psin back to amplitude units (<= 28000 / 40000) */
    tot = (shared * synth_share[c & 7] + own * (4 - synth_share[c & 7])) >> 16;
    env = 16384 + orc_psin((uint32_t)(9000u * (uint32_t)t + (uint32_t)(seed * 2654435761u)));
    tot = (tot * env) >> 16; /* peak about 14000 at 16-bit scale */
    {
        uint64_t h = orc_mix64(seed ^ ((uint64_t)c << 56) ^ t);
        if (bps >= 16) {
            noise = (int32_t)(h & ((1u << (bps - 6)) - 1)) - (int32_t)(1u << (bps - 7));
            v = tot * ((int64_t)1 << (bps - 16)) + noise;
        } else {
            noise = (int32_t)(h & 0x3FF) - 512;
            v = (tot + noise) >> (16 - bps);
        }
    }
    lo = -((int64_t)1 << (bps - 1));
    hi = ((int64_t)1 << (bps - 1)) - 1;
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return (int32_t)v;
}

void orc_synth_pcm(uint64_t seed, unsigned channels, unsigned bps, uint64_t first_frame,
                   uint64_t n_frames, int32_t *out)
{
    uint64_t t;
    unsigned c;
    for (t = 0; t < n_frames; t++)
        for (c = 0; c < channels; c++)
            out[t * channels + c] = synth_sample(seed, c, bps, first_frame + t);
}
