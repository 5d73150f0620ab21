/* flac_types.h -- structures shared by the host side of the engine and its
 * sm_100a kernels.  Layouts are plain-old-data so they can be copied between
 * host and device verbatim.
 *
 * Vocabulary (follows the reference, src/encoders/flac.c):
 *   frame      one FLAC frame = one block of PCM frames, all channels
 *   candidate  one channel signal the encoder may code as a subframe:
 *              channels 0..C-1, or for the stereo path the four signals
 *              left, right, average (mid) and difference (side) of flac.c:548-570
 *   unit       (frame, candidate): the independent work item of the analysis
 *   plan       what the analysis decided for a unit (b200flac_plan, b200flac.h)
 */
#ifndef B200FLAC_TYPES_H
#define B200FLAC_TYPES_H

#include <stdint.h>
#include "../../include/b200flac.h"

#define BF_MAX_ORDER 32
#define BF_MAX_PO 15

enum { BF_CONSTANT = 0, BF_VERBATIM = 1, BF_FIXED = 2, BF_LPC = 3 };

/* options as the kernels see them (derived fields precomputed on the host) */
typedef struct bf_dev_params {
    uint32_t block_size;
    uint32_t max_lpc_order;
    uint32_t po_lim;             /* min(max_residual_partition_order, 15) */
    uint32_t channels;
    uint32_t bps;
    uint32_t bytes_ps;
    uint32_t sample_rate;
    uint32_t precision;          /* qlp_coeff_precision, flac.c:165-178 */
    uint32_t max_rice;           /* flac.c:180-184 */
    uint32_t K;                  /* candidates per frame */
    uint32_t stereo;             /* 1: stereo decorrelation path (flac.c:532) */
    uint32_t mid_side;
    uint32_t exhaustive;
    uint32_t try_verbatim, try_constant, try_fixed, try_lpc;
    uint32_t rice_stride;        /* bytes of Rice parameters kept per unit: 1 << po_lim */
    uint32_t model_stride;       /* int16 coefficient slots per unit in the LPC model */
    uint32_t heap_entries;       /* 2 << po_lim: partition-sum heap entries per unit */
    uint32_t samples_in_smem;    /* 1: a block's samples + residuals fit in shared memory */
    uint32_t heap_in_smem;       /* 1: partition heap fits in shared memory */
    uint32_t samp_stride;        /* ints per unit in the global sample scratch (when not in smem) */
    /* frame-header codes, flac.c:427-486 (block-size code is per frame) */
    uint32_t sr_code;
    uint32_t bps_code;
} bf_dev_params;

/* one frame of a batch */
typedef struct bf_frame_desc {
    uint64_t pcm_off;        /* first PCM frame of the block inside the batch's pcm buffer */
    uint32_t nsamp;          /* PCM frames in the block */
    uint32_t frame_number;
    uint32_t window_off;     /* offset (in doubles) of this length's Tukey window */
    uint32_t pad;
} bf_frame_desc;

/* work item of the autocorrelation kernel: a run of consecutive frames of one length */
typedef struct bf_lpc_task {
    uint32_t first_frame;
    uint32_t n_frames;       /* <= 32 / K */
} bf_lpc_task;

/* LPC model of one unit, written by the model kernel:
 *   shift[o-1], coefficient set of order o at coef[o*(o-1)/2 .. +o)
 * Non-exhaustive search fills only the estimated best order. */
typedef struct bf_lpc_head {
    uint8_t best_order;      /* estimated order (non-exhaustive), 0 if LPC not available */
    uint8_t precision;       /* qlp precision to signal (2 for the dummy model, flac.c:1124) */
    uint8_t dummy;           /* 1: block shorter than max_lpc_order + 2 -> order 1, coeff [1] */
    uint8_t pad;
    int8_t  shift[BF_MAX_ORDER];
} bf_lpc_head;

/* frame-level choice */
typedef struct bf_frame_choice {
    uint32_t unit[B200FLAC_MAX_CHANNELS];    /* unit index of each subframe, bitstream order */
    uint32_t bitoff[B200FLAC_MAX_CHANNELS];  /* bit offset of each subframe from the frame start */
    uint32_t frame_bytes;                    /* whole frame including CRC-16 */
    uint32_t header_bytes;                   /* including CRC-8 */
    uint8_t  assignment;                     /* 4-bit channel assignment */
    uint8_t  n_sub;
    uint8_t  side_slot;                      /* slot coded at bps+1, 0xFF if none */
    uint8_t  pad;
    uint32_t header_words[4];                /* the frame header with its CRC-8 (header_bytes bytes, zero padded),
                                                as big-endian words: what k_pack_v3 ORs into the frame image */
} bf_frame_choice;

#endif
