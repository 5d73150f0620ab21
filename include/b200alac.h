/* b200alac.h -- C ABI of the B200-native ALAC encoder (the second sibling lossless encoder that
 * shares the FLAC engine's runtime).
 *
 * Replaces, for all framesets of a stream at once, the reference's per-frameset call
 *     write_frameset(output, &encoder, channels)                src/encoders/alac.c:288-373
 * and everything below it (write_frame :375-403, write_compressed_frame :438-545 with its search over the
 * interlacing leftweights, compute_coefficients :720-777 -- Tukey window, autocorrelation, Levinson-Durbin,
 * quantisation at orders 4 and 8 --, calculate_residuals :936-1018 with its adaptive coefficients,
 * encode_residuals :1034-1100, the big-endian BitstreamWriter), and in b200alac_encode_mdat the standalone
 * driver's output (ALACEncoder_encode_alac :95-214: the mdat atom).  audiotools.encoders.encode_alac(file,
 * pcmreader, block_size, initial_history, history_multiplier, maximum_k, ...) (alac.c:30-93) is a thin wrapper
 * over b200alac_encode_framesets.
 *
 * Plain pointers and sizes only.  PCM is interleaved, signed, little-endian, bits_per_sample/8 bytes per sample.
 * There is no CPU fallback: every entry point fails when no CUDA device is usable.
 */
#ifndef B200ALAC_H
#define B200ALAC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* struct alac_encoding_options (src/encoders/alac.h:29-36) plus the stream's shape */
typedef struct b200alac_params {
    uint32_t channels;                          /* 1..8 */
    uint32_t bits_per_sample;                   /* 16 or 24 (alac.c:76-80) */
    uint32_t block_size;                        /* 4096 in ALACAudio.from_pcm */
    uint32_t initial_history;                   /* 10 */
    uint32_t history_multiplier;                /* 40 */
    uint32_t maximum_k;                         /* 14 */
    uint32_t minimum_interlacing_leftweight;    /* 0 */
    uint32_t maximum_interlacing_leftweight;    /* 4 */
} b200alac_params;

const char *b200alac_last_error(void);   /* thread-local, never NULL */
void        b200alac_free(void *p);

/* Host PCM -> the framesets of the stream, back to back (each byte aligned: what encode_alac writes after its
 * 8-byte mdat header) and their sizes in bytes (alac_log_output's list, alac.c:1289-1325).  frame_lengths
 * (optional, summing to n_pcm_frames): PCM frames of every frameset -- the reference encodes whatever
 * pcmreader->read(block_size) returns (alac.c:163-183); NULL cuts every block_size frames.  *out and
 * *frame_sizes are malloc'd (b200alac_free).  kernel_ms[4] (optional): model (window, autocorrelation,
 * Levinson, quantisation), candidate sizing (adaptive residual + adaptive Golomb code of every leftweight,
 * channel and order), selection + offsets, emission.  Returns 0 on success. */
int b200alac_encode_framesets(const b200alac_params *params, const uint8_t *pcm, uint64_t n_pcm_frames,
                              const uint32_t *frame_lengths, uint32_t n_lengths, int device,
                              uint8_t **out, uint64_t *out_bytes, uint32_t **frame_sizes, uint32_t *n_frames,
                              float *kernel_ms);

/* The same with the PCM already in device memory and the framesets left in device memory (d_out 16-byte aligned;
 * b200alac_output_bound() bytes are always enough). */
uint64_t b200alac_output_bound(const b200alac_params *params, uint64_t n_pcm_frames, uint32_t n_framesets);
int b200alac_encode_device(const b200alac_params *params, const void *d_pcm, uint64_t n_pcm_frames, int device,
                           void *d_out, uint64_t out_capacity, uint64_t *out_bytes, uint32_t *frame_sizes,
                           uint32_t *n_frames, float *kernel_ms);

/* What the reference's standalone `alacenc` writes (alac.c:95-214): 32-bit atom size, "mdat", the framesets. */
int b200alac_encode_mdat(const char *filename, const b200alac_params *params, const uint8_t *pcm,
                         uint64_t n_pcm_frames, int device);

#ifdef __cplusplus
}
#endif
#endif
