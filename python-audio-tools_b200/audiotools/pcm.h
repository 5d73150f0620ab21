/* pcm.h -- C layout of audiotools.pcm.FrameList, identical to the reference's
 * (src/pcm.h:40-54) so C encoders can read `samples` directly. */
#ifndef B200_AUDIOTOOLS_PCM_H
#define B200_AUDIOTOOLS_PCM_H
#include <Python.h>
#include <stdint.h>

typedef struct {
    PyObject_HEAD
    unsigned int frames;          /* PCM frames = rows of the samples array */
    unsigned int channels;        /* columns */
    unsigned int bits_per_sample;
    int *samples;                 /* interleaved, 32-bit signed */
    unsigned samples_length;      /* frames * channels */
} pcm_FrameList;

#endif
