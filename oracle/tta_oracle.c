/* tta_oracle.c -- CPU restatement of the reference's TTA encoder.
 *
 * TEST INFRASTRUCTURE ONLY (like oracle/flac_oracle.c): imported by tests/ and nothing else; the product
 * path (python-audio-tools_b200/) never links, loads or executes it.
 *
 * Follows /root/reference/src/encoders/tta.c function by function:
 *   correlate_channels   tta.c:264-293      fixed_prediction   tta.c:295-312
 *   hybrid_filter        tta.c:314-399      encode_frame       tta.c:144-262 (adaptive Rice coding, CRC-32)
 *   write_header         tta.c:562-580      write_seektable    tta.c:582-595      main (file layout) tta.c:412-560
 * and src/common/tta_crc.c (the reflected CRC-32, polynomial 0xEDB88320) and the little-endian BitstreamWriter
 * (src/bitstream.c: write = LSB first; write_unary(0, n) = n one-bits then a zero, :2234-2251).
 * PINNED: tests/test_tta_oracle.py compares whole files with the compiled, unmodified reference
 * (oracle/_ref/ttaenc) and with tests/golden/tta_golden.json made from it.
 *
 * Arithmetic notes the GPU kernels have to reproduce:
 *   - the hybrid filter's sum is formed in 32-bit int arithmetic (every operand is int32_t; the int64_t
 *     only receives the result), so it wraps modulo 2^32 (compile with -fwrapv to make that defined here);
 *   - correlate_channels divides by two with C's truncating division;
 *   - the first residual of a frame is the predicted value itself (round >> shift == 0).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static uint32_t crc_table[256];
static int crc_ready = 0;

static void crc_init(void)
{
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
        crc_table[i] = c;
    }
    crc_ready = 1;
}

/* tta_crc32(byte, &checksum), src/common/tta_crc.c */
static uint32_t crc_bytes(uint32_t crc, const uint8_t *p, size_t n)
{
    for (size_t i = 0; i < n; i++) crc = crc_table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
    return crc;
}

typedef struct {
    uint8_t *buf;
    size_t cap, len;   /* bytes */
    uint32_t acc;      /* pending bits, LSB first */
    unsigned fill;
} bw_le;

static void bw_byte(bw_le *w, uint8_t b)
{
    if (w->len == w->cap) {
        w->cap = w->cap ? w->cap * 2 : 4096;
        w->buf = (uint8_t *)realloc(w->buf, w->cap);
    }
    w->buf[w->len++] = b;
}

/* bw_write_bits_*_le: the low `count` bits of value, least significant first */
static void bw_write(bw_le *w, unsigned count, uint32_t value)
{
    while (count) {
        const unsigned take = (8 - w->fill) < count ? (8 - w->fill) : count;
        w->acc |= (value & ((1u << take) - 1u)) << w->fill;
        w->fill += take;
        value >>= take;
        count -= take;
        if (w->fill == 8) { bw_byte(w, (uint8_t)w->acc); w->acc = 0; w->fill = 0; }
    }
}

/* bw_write_unary_f_p_r with stop bit 0: `value` one-bits in chunks of at most 30, then a zero */
static void bw_unary0(bw_le *w, uint32_t value)
{
    while (value > 0) {
        const unsigned n = value <= 30 ? value : 30;
        bw_write(w, n, (1u << n) - 1u);
        value -= n;
    }
    bw_write(w, 1, 0);
}

static void bw_align(bw_le *w)
{
    if (w->fill) bw_write(w, 8 - w->fill, 0);
}

/* tta.c:295-312 */
static void fixed_prediction(const int *ch, unsigned n, unsigned bps, int *pred)
{
    const unsigned shift = (bps == 8) ? 4 : 5;
    pred[0] = ch[0];
    for (unsigned i = 1; i < n; i++) {
        const int64_t v = (((int64_t)ch[i - 1]) << shift) - ch[i - 1];
        pred[i] = ch[i] - (int)(v >> shift);
    }
}

/* tta.c:314-399 */
static void hybrid_filter(const int *pred, unsigned n, unsigned bps, int *res)
{
    const int32_t shift = (bps == 16) ? 9 : 10;
    const int32_t round = 1 << (shift - 1);
    int32_t qm[8] = {0}, dx[8] = {0}, dl[8] = {0};
    for (unsigned i = 0; i < n; i++) {
        int p, r;
        if (i == 0) {
            p = pred[0];
            r = p + (round >> shift);
        } else {
            int64_t sum;
            if (res[i - 1] < 0) { for (int j = 0; j < 8; j++) qm[j] -= dx[j]; }
            else if (res[i - 1] > 0) { for (int j = 0; j < 8; j++) qm[j] += dx[j]; }
            sum = round + (dl[0] * qm[0]) + (dl[1] * qm[1]) + (dl[2] * qm[2]) + (dl[3] * qm[3]) +
                  (dl[4] * qm[4]) + (dl[5] * qm[5]) + (dl[6] * qm[6]) + (dl[7] * qm[7]);   /* int arithmetic */
            p = pred[i];
            r = p - (int)(sum >> shift);
        }
        res[i] = r;
        dx[0] = dx[1]; dx[1] = dx[2]; dx[2] = dx[3]; dx[3] = dx[4];
        dx[4] = (dl[4] >= 0) ? 1 : -1;
        dx[5] = (dl[5] >= 0) ? 2 : -2;
        dx[6] = (dl[6] >= 0) ? 2 : -2;
        dx[7] = (dl[7] >= 0) ? 4 : -4;
        dl[0] = dl[1]; dl[1] = dl[2]; dl[2] = dl[3]; dl[3] = dl[4];
        dl[4] = -dl[5] + (-dl[6] + (p - dl[7]));
        dl[5] = -dl[6] + (p - dl[7]);
        dl[6] = p - dl[7];
        dl[7] = p;
    }
}

/* encode_frame, tta.c:144-262.  ch[c][i]: the frame's samples, planar.  Returns the frame's size in bytes. */
static unsigned encode_frame(bw_le *w, int **ch, unsigned channels, unsigned n, unsigned bps)
{
    const size_t start = w->len;
    int **res = (int **)malloc(channels * sizeof(int *));
    int *corr = (int *)malloc((size_t)n * sizeof(int));
    int *prev_corr = (int *)malloc((size_t)n * sizeof(int));
    int *pred = (int *)malloc((size_t)n * sizeof(int));
    for (unsigned c = 0; c < channels; c++) {
        res[c] = (int *)malloc((size_t)n * sizeof(int));
        const int *in = ch[c];
        if (channels > 1) {
            /* correlate_channels, tta.c:264-293 */
            if (c < channels - 1) for (unsigned i = 0; i < n; i++) corr[i] = ch[c + 1][i] - ch[c][i];
            else for (unsigned i = 0; i < n; i++) corr[i] = ch[c][i] - (prev_corr[i] / 2);
            in = corr;
        }
        fixed_prediction(in, n, bps, pred);
        hybrid_filter(pred, n, bps, res[c]);
        if (channels > 1) memcpy(prev_corr, corr, (size_t)n * sizeof(int));
    }
    int k0[16], sum0[16], k1[16], sum1[16];
    for (unsigned c = 0; c < channels; c++) { k0[c] = 10; sum0[c] = 1 << 14; k1[c] = 10; sum1[c] = 1 << 14; }
    for (unsigned i = 0; i < n; i++)
        for (unsigned c = 0; c < channels; c++) {
            const int r = res[c][i];
            unsigned u;
            if (r > 0) u = (r * 2) - 1; else u = (-r) * 2;
            if (u < (1u << k0[c])) {
                bw_unary0(w, 0);
                bw_write(w, k0[c], u);
            } else {
                const unsigned shifted = u - (1u << k0[c]);
                const unsigned MSB = 1 + (shifted >> k1[c]);
                const unsigned LSB = shifted - ((MSB - 1) << k1[c]);
                bw_unary0(w, MSB);
                bw_write(w, k1[c], LSB);
                sum1[c] += shifted - (sum1[c] >> 4);
                if ((k1[c] > 0) && (sum1[c] < (1 << (k1[c] + 4)))) k1[c] -= 1;
                else if (sum1[c] > (1 << (k1[c] + 5))) k1[c] += 1;
            }
            sum0[c] += u - (sum0[c] >> 4);
            if ((k0[c] > 0) && (sum0[c] < (1 << (k0[c] + 4)))) k0[c] -= 1;
            else if (sum0[c] > (1 << (k0[c] + 5))) k0[c] += 1;
        }
    bw_align(w);
    const uint32_t crc = crc_bytes(0xFFFFFFFFu, w->buf + start, w->len - start) ^ 0xFFFFFFFFu;
    bw_write(w, 32, crc);
    for (unsigned c = 0; c < channels; c++) free(res[c]);
    free(res); free(corr); free(prev_corr); free(pred);
    return (unsigned)(w->len - start);
}

static int unpack(const uint8_t *p, unsigned bytes)
{
    if (bytes == 1) return (int8_t)p[0];
    if (bytes == 2) return (int16_t)(p[0] | (p[1] << 8));
    return ((int32_t)((uint32_t)p[0] << 8 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 24)) >> 8;
}

/* Frames only: what encoders.encode_tta writes to its file object (tta.c:66-83), and the list it returns.
 * pcm: interleaved signed little-endian.  frame_lengths (optional): PCM frames of every TTA frame (the reader's
 * read sizes); NULL cuts every (sample_rate * 256) / 245 frames.  *out is malloc'd.  Returns the frame count. */
unsigned tta_oracle_encode_frames(const uint8_t *pcm, uint64_t n_pcm_frames, unsigned sample_rate, unsigned channels,
                                  unsigned bps, const uint32_t *frame_lengths, unsigned n_lengths,
                                  uint8_t **out, uint64_t *out_bytes, uint32_t *frame_sizes)
{
    if (!crc_ready) crc_init();
    const unsigned block = (unsigned)(((uint64_t)sample_rate * 256) / 245);
    const unsigned B = bps / 8;
    bw_le w;
    memset(&w, 0, sizeof(w));
    int **ch = (int **)malloc(channels * sizeof(int *));
    unsigned maxn = block;
    for (unsigned i = 0; i < n_lengths; i++) if (frame_lengths[i] > maxn) maxn = frame_lengths[i];
    for (unsigned c = 0; c < channels; c++) ch[c] = (int *)malloc((size_t)maxn * sizeof(int));
    uint64_t pos = 0;
    unsigned nf = 0;
    while (pos < n_pcm_frames) {
        unsigned n = frame_lengths ? frame_lengths[nf] : (unsigned)((n_pcm_frames - pos) < block ? (n_pcm_frames - pos) : block);
        for (unsigned i = 0; i < n; i++)
            for (unsigned c = 0; c < channels; c++) ch[c][i] = unpack(pcm + ((pos + i) * channels + c) * B, B);
        const unsigned sz = encode_frame(&w, ch, channels, n, bps);
        if (frame_sizes) frame_sizes[nf] = sz;
        nf++;
        pos += n;
    }
    for (unsigned c = 0; c < channels; c++) free(ch[c]);
    free(ch);
    *out = w.buf;
    *out_bytes = w.len;
    return nf;
}

static void put_le(uint8_t *p, uint32_t v, int bytes)
{
    for (int i = 0; i < bytes; i++) p[i] = (uint8_t)(v >> (8 * i));
}

/* The whole file as the standalone reference writes it (tta.c:412-560): header + CRC, seektable + CRC, frames. */
uint64_t tta_oracle_encode_file(const uint8_t *pcm, uint64_t n_pcm_frames, unsigned sample_rate, unsigned channels,
                                unsigned bps, uint8_t **out)
{
    if (!crc_ready) crc_init();
    const unsigned block = (unsigned)(((uint64_t)sample_rate * 256) / 245);
    const unsigned total_frames = (unsigned)((n_pcm_frames + block - 1) / block);
    uint32_t *sizes = (uint32_t *)malloc((total_frames + 1) * sizeof(uint32_t));
    uint8_t *frames = NULL;
    uint64_t frame_bytes = 0;
    tta_oracle_encode_frames(pcm, n_pcm_frames, sample_rate, channels, bps, NULL, 0, &frames, &frame_bytes, sizes);
    const size_t head = 22 + 4 * (size_t)total_frames + 4;
    uint8_t *f = (uint8_t *)malloc(head + frame_bytes + 1);
    memcpy(f, "TTA1", 4);
    put_le(f + 4, 1, 2); put_le(f + 6, channels, 2); put_le(f + 8, bps, 2);
    put_le(f + 10, sample_rate, 4); put_le(f + 14, (uint32_t)n_pcm_frames, 4);
    put_le(f + 18, crc_bytes(0xFFFFFFFFu, f, 18) ^ 0xFFFFFFFFu, 4);
    for (unsigned i = 0; i < total_frames; i++) put_le(f + 22 + 4 * i, sizes[i], 4);
    put_le(f + 22 + 4 * total_frames, crc_bytes(0xFFFFFFFFu, f + 22, 4 * (size_t)total_frames) ^ 0xFFFFFFFFu, 4);
    if (frame_bytes) memcpy(f + head, frames, frame_bytes);
    free(frames);
    free(sizes);
    *out = f;
    return head + frame_bytes;
}

void tta_oracle_free(void *p) { free(p); }
