/* b200flac.h -- C ABI of the B200-native FLAC encoding engine.
 *
 * This is the drop-in boundary for the reference's FLAC-encode hot path
 * (widgital/python-audio-tools, src/encoders/flac.c).  Plain pointers and
 * sizes only; no CUDA, torch or Python types appear in any signature, so the
 * library can be bound from C, cgo, JNI, ctypes or a CPython extension alike.
 *
 * Two layers:
 *
 *   frame layer  (b200flac_encoder_*)  replaces the reference's per-frame seam
 *       flacenc_write_frame()            src/encoders/flac.h:128-131, flac.c:520-671
 *     and everything below it (flacenc_write_subframe, the FIXED/LPC model
 *     search, flacenc_encode_residuals, the BitstreamWriter bit packing and the
 *     CRC-8/CRC-16 callbacks), for a whole batch of independent frames at once.
 *
 *   stream layer (b200flac_stream_*)   replaces the C-signature entry point
 *       encoders_encode_flac()           src/encoders/flac.c:124-306
 *     i.e. "fLaC" + STREAMINFO/VORBIS_COMMENT/PADDING, the frame loop, the
 *     STREAMINFO MD5 (host thread, overlapped) and the final STREAMINFO rewrite.
 *     audiotools.encoders.encode_flac (the CPython extension in
 *     python-audio-tools_b200/audiotools/) is a thin wrapper over this layer.
 *
 * There is no CPU fallback: every entry point fails (returns NULL / non-zero
 * and sets b200flac_last_error()) when no CUDA device is usable.
 *
 * PCM everywhere in this header is what the reference's standalone encoder
 * reads and what its MD5 callback hashes (flac.c:188,1784-1790): interleaved,
 * signed, little-endian, bits_per_sample/8 bytes per sample.
 */
#ifndef B200FLAC_H
#define B200FLAC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FLAC_ABI_VERSION 1
#define B200FLAC_MAX_CHANNELS 8
#define B200FLAC_MAX_LPC_ORDER 32

/* Encoding options: struct flac_encoding_options (src/encoders/flac.h:30-47)
 * plus the three stream parameters the frame header needs
 * (struct flac_STREAMINFO, flac.h:49-59).  The derived options
 * qlp_coeff_precision / max_rice_parameter (flac.c:165-184) are computed by
 * the library exactly as the reference does. */
typedef struct b200flac_params {
    uint32_t block_size;
    uint32_t max_lpc_order;
    uint32_t min_residual_partition_order; /* accepted and ignored, as in flac.c:1362 */
    uint32_t max_residual_partition_order;
    int32_t  mid_side;
    int32_t  adaptive_mid_side;
    int32_t  exhaustive_model_search;
    int32_t  no_verbatim_subframes;  /* the reference's debug switches, flac.c:62-65 */
    int32_t  no_constant_subframes;
    int32_t  no_fixed_subframes;
    int32_t  no_lpc_subframes;
    uint32_t sample_rate;
    uint32_t channels;          /* 1..8 */
    uint32_t bits_per_sample;   /* 8, 16 or 24 */
} b200flac_params;

/* A run of PCM frames that is cut into blocks on its own: blocks of
 * params.block_size frames, the last one possibly shorter (flac.c:244-274).
 * One segment = (part of) one stream; a batch may carry many streams. */
typedef struct b200flac_segment {
    uint64_t pcm_frame_offset;   /* first PCM frame of the segment inside the pcm buffer */
    uint64_t n_pcm_frames;
    uint32_t first_frame_number; /* FLAC frame number of its first block (flac.c:592 total_flac_frames) */
    uint32_t reserved;
} b200flac_segment;

typedef struct b200flac_encoder b200flac_encoder;
typedef struct b200flac_stream b200flac_stream;

/* ---- library ---- */
int         b200flac_abi_version(void);
int         b200flac_device_count(void);            /* usable CUDA devices; 0 if none */
const char *b200flac_last_error(void);              /* thread-local, never NULL */

/* ---- frame layer ---- */

/* One encoder is bound to one CUDA device and owns `n_slots` in-flight batch
 * slots (pinned staging + device buffers + one CUDA stream each), each able to
 * hold max_pcm_frames_per_batch PCM frames.  Replaces flacenc_init_encoder
 * (flac.c:310-341). */
b200flac_encoder *b200flac_encoder_create(const b200flac_params *params, int device,
                                          uint64_t max_pcm_frames_per_batch, int n_slots);
void              b200flac_encoder_destroy(b200flac_encoder *enc); /* flacenc_free_encoder, flac.c:343-374 */

/* upper bound on the bytes submit/collect can return for a batch */
uint64_t b200flac_encoder_output_bound(const b200flac_encoder *enc, uint64_t n_pcm_frames,
                                       uint32_t n_segments);
/* pinned host staging buffer of a slot (write PCM here to skip one host copy) */
uint8_t *b200flac_encoder_slot_pcm(b200flac_encoder *enc, int slot);
/* pinned host buffer of a slot for collect()'s `out` (large enough for any batch of this encoder; *capacity
 * receives its size): the device->host copy of the frames is then a direct asynchronous DMA */
uint8_t *b200flac_encoder_slot_out(b200flac_encoder *enc, int slot, uint64_t *capacity);

/* Asynchronous: copy PCM host->device (pinned if pcm == slot buffer), run the
 * kernels, start the device->host copy of the per-frame sizes.  Returns 0 on
 * success.  `pcm` holds at least max(pcm_frame_offset + n_pcm_frames) frames. */
int b200flac_encoder_submit(b200flac_encoder *enc, int slot, const uint8_t *pcm,
                            const b200flac_segment *segments, uint32_t n_segments);

/* Wait for the slot, copy the frame bytes device->host.
 *   out          receives the frames of all segments back to back, in order
 *   frame_bytes  receives the size of each frame (n_frames entries)
 *   frame_pcm    receives the PCM-frame count of each frame (may be NULL)
 * Exactly what flacenc_write_frame appended to the recorder per block. */
int b200flac_encoder_collect(b200flac_encoder *enc, int slot, uint8_t *out, uint64_t out_capacity,
                             uint64_t *out_bytes, uint32_t *frame_bytes, uint32_t *frame_pcm,
                             uint32_t frame_capacity, uint32_t *n_frames);

/* submit + collect on slot 0 */
int b200flac_encoder_encode(b200flac_encoder *enc, const uint8_t *pcm,
                            const b200flac_segment *segments, uint32_t n_segments,
                            uint8_t *out, uint64_t out_capacity, uint64_t *out_bytes,
                            uint32_t *frame_bytes, uint32_t *frame_pcm, uint32_t frame_capacity,
                            uint32_t *n_frames);

/* Device-resident variant (inputs already in HBM, frames left in HBM): `d_pcm`
 * and `d_out` are device pointers on the encoder's device.  d_out must be
 * 16-byte aligned and hold b200flac_encoder_output_bound() bytes; d_pcm must be
 * 4-byte aligned, and the 16-byte lines holding its first and last PCM byte must
 * be readable (rows are staged with 16-byte aligned bulk copies; memory from
 * b200flac_device_alloc, which adds 64 bytes, or any interior pointer qualifies).  Returns the
 * total through *out_bytes (host) after synchronising the slot's stream.
 * elapsed_ms (optional) receives the CUDA-event time of the kernels alone. */
int b200flac_encoder_encode_device(b200flac_encoder *enc, int slot, const void *d_pcm,
                                   const b200flac_segment *segments, uint32_t n_segments,
                                   void *d_out, uint64_t out_capacity, uint64_t *out_bytes,
                                   uint32_t *n_frames, float *elapsed_ms);

/* The same in two halves, so that a caller with two slots keeps the device busy: submit_device does the host's
 * share (frame descriptors, task lists, their upload, the launches) and returns; collect_device waits for the
 * slot and returns the totals.  Same argument rules as b200flac_encoder_encode_device. */
int b200flac_encoder_submit_device(b200flac_encoder *enc, int slot, const void *d_pcm,
                                   const b200flac_segment *segments, uint32_t n_segments,
                                   void *d_out, uint64_t out_capacity);
int b200flac_encoder_collect_device(b200flac_encoder *enc, int slot, uint64_t *out_bytes,
                                    uint32_t *n_frames, float *elapsed_ms);

/* After collect_device (or collect) on `slot`: the size in bytes and the PCM-frame count of every FLAC frame of the
 * batch, in order -- host arrays owned by the slot, valid until its next submit. */
const uint32_t *b200flac_encoder_slot_frame_bytes(b200flac_encoder *enc, int slot);
const uint32_t *b200flac_encoder_slot_frame_pcm(b200flac_encoder *enc, int slot);

/* per-kernel CUDA-event times of the slot's last batch, in ms:
 * [0] lpc model (window/autocorrelation, then Levinson/quantise), [1] subframe analysis,
 * [2] frame select + offset scan (+ output clear on the k_pack_v2 path), [3] frame packing
 * (k_pack_v3 computes the CRC-16 here too), [4] the separate CRC-16 kernel of the k_pack_v2 path (else ~0).
 * Returns the number of entries written; 0 when the batch ran as a pipeline of chunks (kernels of different
 * chunks overlap, so there is no per-kernel time: see b200flac_encoder_set_chunking). */
int b200flac_encoder_last_kernel_ms(b200flac_encoder *enc, int slot, float *ms, int capacity);
/* How a batch is scheduled on the device.  A batch of more than ~1.5 x chunk_frames FLAC frames runs as a
 * software pipeline of chunks over three CUDA streams: the floating-point model kernels of the next
 * `lookahead` chunks (FP64 pipe) overlap the integer analysis/packing kernels of the current one, and a chunk's
 * PCM is re-read from L2 instead of HBM.  chunk_frames = 0: the whole batch is one chunk, kernels back to back
 * on one stream (what b200flac_encoder_last_kernel_ms needs; the default -- on B200 the pipeline measured no faster).  Output bytes do not depend
 * on it.  No reference counterpart (the reference encodes one frame at a time, flac.c:247-274). */
int b200flac_encoder_set_chunking(b200flac_encoder *enc, uint32_t chunk_frames, uint32_t lookahead);
/* number of kernel launches issued by this encoder so far */
uint64_t b200flac_encoder_launch_count(const b200flac_encoder *enc);
/* ... and by every encoder of this process together (the stream layer's pooled encoders included) */
uint64_t b200flac_launch_count_total(void);

/* page-locked host memory: PCM handed to submit() from such memory (or from
 * b200flac_encoder_slot_pcm) is copied host->device directly and asynchronously;
 * pageable memory is first staged through the slot's pinned buffer */
void *b200flac_host_alloc(uint64_t bytes);
void  b200flac_host_free(void *p);

/* raw device memory helpers so a host language without a CUDA binding can stage
 * device-resident input (bench.py uses them; torch is not required) */
void *b200flac_device_alloc(int device, uint64_t bytes);
void  b200flac_device_free(int device, void *ptr);
int   b200flac_device_upload(int device, void *dst, const void *src, uint64_t bytes);
int   b200flac_device_download(int device, void *dst, const void *src, uint64_t bytes);
/* fill device memory with the deterministic synthetic PCM of SURVEY.md 8(d) */
int   b200flac_device_synth_pcm(int device, void *d_pcm, uint64_t seed, uint32_t channels,
                                uint32_t bits_per_sample, uint64_t first_frame, uint64_t n_frames);

/* ---- debug / test hooks (forced-decision packing) ---- */

/* what the analysis kernels decided for one subframe candidate */
typedef struct b200flac_plan {
    uint8_t  type;            /* 0 CONSTANT, 1 VERBATIM, 2 FIXED, 3 LPC */
    uint8_t  order;
    uint8_t  wasted;
    uint8_t  precision;
    int8_t   shift;
    uint8_t  coding_method;
    uint8_t  partition_order;
    uint8_t  flags;           /* bit0: partition-length underflow level (flac.c:1462);
                                 bit1: hint, LPC sum provably fits 32 bits (packer may use int32) */
    uint32_t bits;            /* exact subframe size in bits */
    int16_t  coeffs[B200FLAC_MAX_LPC_ORDER];
} b200flac_plan;

/* After a collect/encode on `slot`: copy out the plans of every candidate
 * (n_frames * candidates_per_frame entries, candidate-major inside a frame:
 * channels 0..C-1, or L,R,M,S when stereo decorrelation ran), their Rice
 * parameters (rice_stride bytes per candidate) and each frame's channel
 * assignment. Any pointer may be NULL. */
int b200flac_encoder_get_plans(b200flac_encoder *enc, int slot, b200flac_plan *plans,
                               uint8_t *rice, uint32_t *rice_stride,
                               uint8_t *assignments, uint32_t *candidates_per_frame);

/* ---- stream layer ---- */

/* Opens `filename` for writing and emits the stream head exactly as
 * flac.c:209-238: "fLaC", STREAMINFO placeholder, VORBIS_COMMENT with vendor
 * string "Python Audio Tools " + version (NULL -> "2.22alpha1"), PADDING of
 * padding_size bytes.  devices/n_devices: the CUDA devices to shard frame
 * ranges over (NULL/0 -> device 0). */
b200flac_stream *b200flac_stream_open(const char *filename, const b200flac_params *params,
                                      uint32_t padding_size, const char *version,
                                      const int *devices, int n_devices);
/* Appends PCM (any number of PCM frames per call).  Blocks are cut every
 * block_size frames across calls, like BufferedPCMReader feeding
 * pcmreader->read(block_size) (flac.c:244,272). */
int b200flac_stream_write(b200flac_stream *s, const uint8_t *pcm, uint64_t n_pcm_frames);
/* Forces a frame boundary: PCM written since the last boundary that does not fill
 * a whole block becomes a short frame -- what the reference does when
 * pcmreader->read() returns fewer than block_size frames mid-stream
 * (flac.c:247,525; it encodes whatever length it is given). */
int b200flac_stream_end_block(b200flac_stream *s);
/* Flushes the tail block, waits for the MD5 thread, rewrites STREAMINFO at
 * byte 8 (flac.c:277-279) and closes the file.  The (offset, pcm_frames) list of
 * flac.c:249-253 is returned through malloc'd arrays the caller frees with
 * b200flac_free().  On abort != 0 the file is closed without finalising. */
int b200flac_stream_close(b200flac_stream *s, int abort_encode, uint64_t **frame_offsets,
                          uint32_t **frame_pcm_frames, uint64_t *n_frames);
void b200flac_free(void *p);
/* The stream layer keeps up to 16 (env B200FLAC_POOL) idle encoders (device buffers, pinned staging) for the next stream with
 * the same options and device: set-up costs ~100 ms, a short file a few ms.  This releases them, and the
 * device buffers the decode layer keeps between calls (candidates, sample rows, chain arrays). */
void b200flac_pool_clear(void);

/* Metadata finalisation of a finished file, host only: what FlacAudio.from_pcm does in Python after
 * encode_flac returns (audiotools/flac.py:1811-1832).  Builds the SEEKTABLE from the encoder's
 * (byte offset, PCM frames) list, one point every seekpoint_interval PCM frames (0 -> 10 s, flac.py:1847-1876),
 * inserts it in FlacMetaData.add_block's order (flac.py:53-75), adds
 * WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x%.4X to the VORBIS_COMMENT when channel_mask != 0 (flac.py:1827-1832)
 * and writes the metadata back by update_metadata's rule (flac.py:1369-1462): the PADDING blocks absorb
 * the growth when they can (frames do not move), else the file is rewritten.  Returns 0 on success. */
int b200flac_finalize_metadata(const char *filename, const uint64_t *frame_offsets,
                               const uint32_t *frame_pcm_frames, uint64_t n_frames,
                               uint32_t seekpoint_interval, uint32_t channel_mask);

/* One-call form of the above for PCM already in memory: the standalone
 * reference's `flacenc < pcm` (flac.c:1637-1804). */
int b200flac_encode_file(const char *filename, const b200flac_params *params,
                         uint32_t padding_size, const char *version,
                         const uint8_t *pcm, uint64_t n_pcm_frames,
                         const int *devices, int n_devices);

/* Many tracks -> many files in one call: BASELINE.json config 5 ("a batch of 10,000 three-minute tracks").  No
 * reference counterpart -- FlacAudio.from_pcm (audiotools/flac.py:1235-1330) calls encoders.encode_flac once per
 * file -- and no new format: file i is byte for byte what b200flac_encode_file(filenames[i], params, padding_size,
 * version, pcm[i], n_pcm_frames[i], ...) writes.  The difference is where the time goes: tracks are packed into
 * many-segment batches of the frame layer, and the STREAMINFO MD5 of every track (the per-stream serial step that
 * bounds the one-call-per-file path: ~50 ms of one core per three-minute track) is computed on the device, one thread
 * per track, from the PCM that is there for the encoder anyway -- whole batches from the front of the list as they
 * arrive -- while host threads with no file to write hash the end of the list (what the device could not finish in
 * time) out of pcm[i], sixteen tracks at a time in the lanes of one vector (a few long tracks: host only).
 * pcm[i] should be page-locked (b200flac_host_alloc) for full copy speed.  device < 0: B200FLAC_DEVICE or 0.
 * host_threads (<= 0: 8) write the files and share the hashing.
 * Returns 0 on success. */
int b200flac_encode_files(uint32_t n_tracks, const char *const *filenames, const b200flac_params *params,
                          uint32_t padding_size, const char *version,
                          const uint8_t *const *pcm, const uint64_t *n_pcm_frames,
                          int device, int host_threads);

/* ---- file-backed PCM feed ----
 * The reference feeds the encoder through a Python PCMReader: WaveReader / AiffReader read the file,
 * build an int FrameList per call (audiotools/wav.py:504-527, aiff.py:434-456) and pcmreader_read
 * converts it again (src/pcmconv.c:219-297).  These entry points restate the two container parsers in C
 * and move the PCM from the file straight into the stream's pinned staging, one batch per read. */
#define B200FLAC_PCM_BIG_ENDIAN 1u /* samples are big-endian in the file (AIFF, aiff.py:452-456) */
#define B200FLAC_PCM_UNSIGNED   2u /* samples are unsigned in the file (8-bit WAVE, wav.py:523-527) */

typedef struct b200flac_pcm_source {
    uint32_t sample_rate, channels, bits_per_sample, channel_mask;
    uint32_t flags;             /* B200FLAC_PCM_* */
    uint32_t reserved;
    uint64_t data_offset;       /* file offset of the first PCM byte */
    uint64_t total_pcm_frames;  /* what the container says: data size / frame size (WAVE), COMM (AIFF) */
} b200flac_pcm_source;

/* WaveReader.__init__ + parse_fmt (audiotools/wav.py:424-502, 288-354) / AiffReader.__init__ + parse_comm
 * (audiotools/aiff.py:353-432, 327-347), chunk walk, checks and defaults included (the channel mask a
 * plain WAVE gets from its channel count, the `fmt ` remainder the reference does not skip).
 * Returns 0, or: 1 = the reference raises ValueError, 2 = IOError (unreadable/truncated header);
 * b200flac_last_error() then holds the reference's message text (audiotools/text.py:530-544,621-634). */
int b200flac_wave_probe(const char *path, b200flac_pcm_source *src);
int b200flac_aiff_probe(const char *path, b200flac_pcm_source *src);

/* Appends n_pcm_frames PCM frames read from `path` at byte_offset, exactly as if the same frames had gone
 * through b200flac_stream_write() after conversion to signed little-endian (flags: B200FLAC_PCM_*), but
 * read directly into the pinned staging of the batch being filled.  Returns 0; 1 on an encoder or I/O
 * error; 2 when the file ends early -- the reference's "premature end of data chunk" IOError
 * (wav.py:516-518, aiff.py:445-447); the stream is then failed and only good for b200flac_stream_close(abort). */
int b200flac_stream_write_file(b200flac_stream *s, const char *path, uint64_t byte_offset,
                               uint64_t n_pcm_frames, uint32_t flags);

/* One call, file to file: FlacAudio.from_pcm(flac, WaveAudio(wave).to_pcm() / AiffAudio(aiff).to_pcm())
 * without its metadata tail -- probe, stream head, feed, STREAMINFO.  sample_rate / channels /
 * bits_per_sample in *params are ignored and taken from the file.  `src` (optional) receives the probe
 * result (the caller needs channel_mask for b200flac_finalize_metadata); the (offset, pcm_frames) list
 * comes back as in b200flac_stream_close.  Returns 0, or 1/2 as the probe, or 3 for an encode error. */
int b200flac_encode_wave(const char *flac_filename, const char *wave_filename, const b200flac_params *params,
                         uint32_t padding_size, const char *version, const int *devices, int n_devices,
                         b200flac_pcm_source *src, uint64_t **frame_offsets, uint32_t **frame_pcm_frames,
                         uint64_t *n_frames);
int b200flac_encode_aiff(const char *flac_filename, const char *aiff_filename, const b200flac_params *params,
                         uint32_t padding_size, const char *version, const int *devices, int n_devices,
                         b200flac_pcm_source *src, uint64_t **frame_offsets, uint32_t **frame_pcm_frames,
                         uint64_t *n_frames);

/* ---- decode / verify ----
 * Frame-parallel restatement of the reference decoder, src/decoders/flac.c:174-286 (FlacDecoder_read),
 * :569-1270 (metadata, frame header, subframes, residuals, channel decorrelation) and :1340-1510 (flacdec):
 * every position holding a header that is valid for the stream is decoded speculatively by its own GPU
 * thread, the chain of frame ends is resolved on the device (the host walks it, as the reference's loop would,
 * only for streams the device chain does not vouch for), and a last kernel
 * writes the interleaved PCM (signed little-endian, the bytes the STREAMINFO MD5 is taken over).
 * Return codes: 0; 1 = the reference raises ValueError (flacdec: "*** Error: <text>"), 2 = IOError
 * (EOF), 3 = engine error; b200flac_last_error() holds the reference's message text
 * (FlacDecoder_strerror, flac.c:1273-1311; "invalid checksum in frame"; "MD5 mismatch at end of stream"). */
typedef struct b200flac_stream_info {
    uint32_t min_block_size, max_block_size, min_frame_size, max_frame_size;
    uint32_t sample_rate, channels, bits_per_sample;
    uint32_t channel_mask;        /* from the channel count, or the VORBIS_COMMENT's WAVEFORMATEXTENSIBLE_CHANNEL_MASK (flac.c:509-566,626-658) */
    uint64_t total_pcm_frames;
    uint64_t first_frame_offset;  /* byte offset of the first frame in the file */
    uint8_t  md5[16];
} b200flac_stream_info;

/* flacdec_read_metadata (flac.c:569-708) as far as decoding needs it */
int b200flac_read_streaminfo(const uint8_t *flac, uint64_t n_bytes, b200flac_stream_info *info);
/* A whole FLAC file image in host memory -> PCM in host memory.  pcm == NULL only fills *info (size the
 * buffer as total_pcm_frames * channels * bits_per_sample/8).  check_md5: compare with STREAMINFO as
 * FlacDecoder_verify_okay does (flac.c:479-490).  The frames' byte offsets (relative to the first frame)
 * and PCM lengths come back like the encoder's list (free with b200flac_free); kernel_ms[3] (optional):
 * candidate scan, frame decode, chain walk + PCM emit. */
int b200flac_decode_memory(const uint8_t *flac, uint64_t n_bytes, int device, uint8_t *pcm,
                           uint64_t pcm_capacity, b200flac_stream_info *info, int check_md5,
                           uint64_t **frame_offsets, uint32_t **frame_pcm_frames, uint64_t *n_frames,
                           float *kernel_ms);
/* The same with the frames (first frame to end of file; 16-byte aligned, 64 readable bytes of padding
 * after n_bytes) already in device memory and the PCM left in device memory (4-byte aligned). */
int b200flac_decode_device(const b200flac_stream_info *info, const void *d_frames, uint64_t n_bytes,
                           int device, void *d_pcm, uint64_t pcm_capacity, uint64_t *n_frames,
                           float *kernel_ms);
/* flacdec without output: decode everything, check every CRC-16 and the STREAMINFO MD5 */
int b200flac_verify_file(const char *flac_filename, int device);

/* FLAC file -> RIFF WAVE file in one call: WaveAudio.from_pcm(wave, FlacAudio(flac).to_pcm())
 * (audiotools/wav.py:357-418 wave_header, :660-729 from_pcm), MD5 and every frame CRC checked on the way.
 * Return codes as for b200flac_decode_memory; "total size too large for wave file" is a ValueError (1). */
int b200flac_decode_to_wave(const char *flac_filename, const char *wave_filename, int device);

#ifdef __cplusplus
}
#endif
#endif /* B200FLAC_H */
