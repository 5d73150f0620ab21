/* b200tta.h -- C ABI of the B200-native TTA encoder (the sibling lossless encoder that shares
 * the FLAC engine's runtime).
 *
 * Replaces, for all frames of a stream at once, the reference's per-frame call
 *     encode_frame(output, &cache, framelist, bits_per_sample)      src/encoders/tta.c:144-262
 * (correlate_channels :264-293, fixed_prediction :295-312, hybrid_filter :314-399, the adaptive Rice coder and
 * the frame CRC-32 :199-261) and, in b200tta_encode_file, the standalone driver's file layout (write_header
 * :562-580, write_seektable :582-595, main :412-560).  audiotools.encoders.encode_tta(file, pcmreader)
 * (tta.c:31-117) is a thin wrapper over b200tta_encode_frames.
 *
 * Plain pointers and sizes only.  PCM is interleaved, signed, little-endian, bits_per_sample/8 bytes per sample.
 * There is no CPU fallback: every entry point fails when no CUDA device is usable.
 */
#ifndef B200TTA_H
#define B200TTA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200tta_params {
    uint32_t sample_rate;
    uint32_t channels;          /* 1..8 */
    uint32_t bits_per_sample;   /* 8, 16 or 24 */
} b200tta_params;

const char *b200tta_last_error(void);   /* thread-local, never NULL */
void        b200tta_free(void *p);

/* PCM frames of one TTA frame: (sample_rate * 256) / 245, tta.c:64,488 */
uint32_t b200tta_block_size(uint32_t sample_rate);

/* Host PCM -> the TTA frames of the stream, back to back (what encode_tta writes to its file object) and the
 * list of their sizes in bytes (what it returns).  frame_lengths (optional, n_lengths entries summing to
 * n_pcm_frames): the PCM frames of every TTA frame -- the reference encodes whatever length pcmreader->read()
 * returns (tta.c:69-83); NULL cuts a frame every b200tta_block_size() PCM frames.  *out and *frame_sizes are
 * malloc'd (b200tta_free).  kernel_ms[3] (optional): residual/Rice-state kernel, size scan, bit packing + CRC-32.
 * Returns 0 on success. */
int b200tta_encode_frames(const b200tta_params *params, const uint8_t *pcm, uint64_t n_pcm_frames,
                          const uint32_t *frame_lengths, uint32_t n_lengths, int device,
                          uint8_t **out, uint64_t *out_bytes, uint32_t **frame_sizes, uint32_t *n_frames,
                          float *kernel_ms);

/* The same with the PCM already in device memory and the frames left in device memory (d_out: 16-byte
 * aligned, out_capacity bytes; b200tta_output_bound() is always enough).  frame_sizes (optional, host): room
 * for ceil(n_pcm_frames / block) entries. */
uint64_t b200tta_output_bound(const b200tta_params *params, uint64_t n_pcm_frames, uint32_t n_frames);
int b200tta_encode_device(const b200tta_params *params, const void *d_pcm, uint64_t n_pcm_frames, int device,
                          void *d_out, uint64_t out_capacity, uint64_t *out_bytes, uint32_t *frame_sizes,
                          uint32_t *n_frames, float *kernel_ms);

/* The whole file as the reference's standalone `ttaenc` writes it: "TTA1" header + CRC-32, seektable + CRC-32,
 * frames (tta.c:412-560).  Returns 0 on success. */
int b200tta_encode_file(const char *filename, const b200tta_params *params, const uint8_t *pcm,
                        uint64_t n_pcm_frames, int device);

#ifdef __cplusplus
}
#endif
#endif
