/* audiotools.encoders -- CPython 3 module exposing encode_flac() with the reference's
 * exact signature, keyword names and return value, backed by the B200 engine.
 *
 * Reference being replaced:
 *   encoders_encode_flac()           src/encoders/flac.c:43-307 (Python entry)
 *   kwlist / format "sO&IIII|iiiiiiiI"   flac.c:52-67, 90-109
 *   pcmreader_converter / pcmreader_read  src/pcmconv.c:127-297
 *   method table row                 src/encoders.h:65-67
 *
 * Host code stays in C here; the per-frame work happens behind the C ABI of
 * libb200flac.so (include/b200flac.h, stream layer).  There is no CPU fallback.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pcm.h"
#include "../../include/b200flac.h"
#include "../../include/b200tta.h"
#include "../../include/b200alac.h"

struct py_pcmreader {
    PyObject *obj;            /* the Python PCMReader */
    PyObject *framelist_type; /* audiotools.pcm.FrameList */
    unsigned sample_rate, channels, channel_mask, bits_per_sample;
};

static int get_uint_attr(PyObject *obj, const char *name, unsigned *out)
{
    PyObject *a = PyObject_GetAttrString(obj, name);
    if (!a) return 0;
    long v = PyLong_AsLong(a);
    Py_DECREF(a);
    if (v == -1 && PyErr_Occurred()) return 0;
    *out = (unsigned)v;
    return 1;
}

/* "O&" converter: src/pcmconv.c:127-217 */
static int pcmreader_converter(PyObject *obj, void **out)
{
    struct py_pcmreader *r = (struct py_pcmreader *)calloc(1, sizeof(*r));
    if (!r) { PyErr_NoMemory(); return 0; }
    if (!get_uint_attr(obj, "sample_rate", &r->sample_rate) || !get_uint_attr(obj, "bits_per_sample", &r->bits_per_sample) ||
        !get_uint_attr(obj, "channels", &r->channels) || !get_uint_attr(obj, "channel_mask", &r->channel_mask)) {
        free(r);
        return 0;
    }
    if (!PyObject_HasAttrString(obj, "read")) { free(r); PyErr_SetString(PyExc_AttributeError, "pcmreader has no read() method"); return 0; }
    if (!PyObject_HasAttrString(obj, "close")) { free(r); PyErr_SetString(PyExc_AttributeError, "pcmreader has no close() method"); return 0; }
    PyObject *pcm = PyImport_ImportModule("audiotools.pcm");
    if (!pcm) { free(r); return 0; }
    r->framelist_type = PyObject_GetAttrString(pcm, "FrameList");
    Py_DECREF(pcm);
    if (!r->framelist_type) { free(r); return 0; }
    Py_INCREF(obj);
    r->obj = obj;
    *out = r;
    return 1;
}

static void pcmreader_del(struct py_pcmreader *r)
{
    if (!r) return;
    Py_XDECREF(r->obj);
    Py_XDECREF(r->framelist_type);
    free(r);
}

/* signed little-endian packing: what the reference's MD5 callback is fed (flac.c:188) */
static void pack_le_signed(const pcm_FrameList *f, uint8_t *out)
{
    const unsigned bytes = f->bits_per_sample / 8;
    const int *s = f->samples;
    const unsigned n = f->samples_length;
    if (bytes == 2) {
        int16_t *o = (int16_t *)out;
        for (unsigned i = 0; i < n; i++) o[i] = (int16_t)s[i];
    } else if (bytes == 3) {
        for (unsigned i = 0; i < n; i++, out += 3) {
            const uint32_t v = (uint32_t)s[i];
            out[0] = (uint8_t)v; out[1] = (uint8_t)(v >> 8); out[2] = (uint8_t)(v >> 16);
        }
    } else {
        for (unsigned i = 0; i < n; i++) out[i] = (uint8_t)s[i];
    }
}

static PyObject *encoders_encode_flac(PyObject *dummy, PyObject *args, PyObject *keywds)
{
    static char *kwlist[] = {"filename", "pcmreader", "block_size", "max_lpc_order",
                             "min_residual_partition_order", "max_residual_partition_order",
                             "mid_side", "adaptive_mid_side", "exhaustive_model_search",
                             "disable_verbatim_subframes", "disable_constant_subframes",
                             "disable_fixed_subframes", "disable_lpc_subframes", "padding_size", NULL};
    char *filename;
    struct py_pcmreader *reader = NULL;
    b200flac_params p;
    unsigned padding_size = 4096; /* DEFAULT_PADDING_SIZE, flac.c:33 */
    b200flac_stream *stream = NULL;
    PyObject *offsets = NULL;
    uint8_t *packed = NULL;
    size_t packed_cap = 0;

    memset(&p, 0, sizeof(p));
    if (!PyArg_ParseTupleAndKeywords(args, keywds, "sO&IIII|iiiiiiiI", kwlist, &filename, pcmreader_converter, &reader,
                                     &p.block_size, &p.max_lpc_order, &p.min_residual_partition_order,
                                     &p.max_residual_partition_order, &p.mid_side, &p.adaptive_mid_side,
                                     &p.exhaustive_model_search, &p.no_verbatim_subframes, &p.no_constant_subframes,
                                     &p.no_fixed_subframes, &p.no_lpc_subframes, &padding_size))
        return NULL;
    p.sample_rate = reader->sample_rate;
    p.channels = reader->channels;
    p.bits_per_sample = reader->bits_per_sample;

    /* flac.c:114-117: a file that cannot be opened is an IOError carrying errno and the name */
    {
        FILE *probe = fopen(filename, "wb");
        if (!probe) { PyErr_SetFromErrnoWithFilename(PyExc_IOError, filename); pcmreader_del(reader); return NULL; }
        fclose(probe);
    }
    Py_BEGIN_ALLOW_THREADS
    stream = b200flac_stream_open(filename, &p, padding_size, NULL, NULL, 0);
    Py_END_ALLOW_THREADS
    if (!stream) {
        /* FlacAudio.from_pcm turns IOError/ValueError into EncodingError (flac.py:1833-1845) and callers of
           the reference catch that: a rejected parameter is a ValueError, everything else (no usable device,
           allocation, the output file) an IOError */
        const char *msg = b200flac_last_error();
        const int bad_param = strstr(msg, "must be") != NULL || strstr(msg, "unsupported") != NULL;
        PyErr_SetString(bad_param ? PyExc_ValueError : PyExc_IOError, msg);
        pcmreader_del(reader);
        return NULL;
    }

    const unsigned bytes_ps = p.bits_per_sample / 8;
    /* The reference asks for block_size frames per call (flac.c:244).  audiotools.BufferedPCMReader --
       what FlacAudio.from_pcm always passes -- returns exactly the count it is asked for until the
       stream ends (audiotools/__init__.py:2561-2603), so asking it for 64 blocks at a time yields the
       same blocks with 1/64 of the Python calls.  Any other reader keeps the per-block protocol, because
       a short read mid-stream must become a short frame (H12). */
    unsigned read_blocks = 1;
    {
        PyObject *mod = PyImport_ImportModule("audiotools");
        PyObject *cls = mod ? PyObject_GetAttrString(mod, "BufferedPCMReader") : NULL;
        if (cls && (PyObject *)Py_TYPE(reader->obj) == cls && p.block_size <= (1u << 20)) read_blocks = 64;
        Py_XDECREF(cls);
        Py_XDECREF(mod);
        PyErr_Clear();
    }
    /* File-backed readers: when from_pcm's BufferedPCMReader wraps a reader that knows
       where its PCM lies in a file (audiotools.wav.WaveReader, audiotools.aiff.AiffReader: b200_file_span)
       and nothing is buffered yet, the engine reads that span itself, batch by batch, straight into its
       pinned staging -- no FrameList per read().  The reader is left as if it had been read to the end, and
       the loop below then sees the empty FrameList that ends the stream.  B200FLAC_FILE_FEED=0 turns this off. */
    if (read_blocks == 64) {
        const char *knob = getenv("B200FLAC_FILE_FEED");
        PyObject *inner = (knob && knob[0] == '0') ? NULL : PyObject_GetAttrString(reader->obj, "pcmreader");
        PyObject *buffered = inner ? PyObject_GetAttrString(reader->obj, "buffer") : NULL;
        PyObject *span = NULL;
        if (inner && buffered && (PyObject *)Py_TYPE(buffered) == reader->framelist_type &&
            ((pcm_FrameList *)buffered)->frames == 0 && PyObject_HasAttrString(inner, "b200_file_span") &&
            PyObject_HasAttrString(inner, "b200_file_span_consumed"))
            span = PyObject_CallMethod(inner, "b200_file_span", NULL);
        Py_XDECREF(buffered);
        if (span) {
            PyObject *path_obj = NULL, *path_bytes = NULL;
            unsigned long long offset = 0, frames = 0;
            unsigned flags = 0;
            const char *truncated = "premature end of data chunk";
            int ok = PyArg_ParseTuple(span, "OKKI|s", &path_obj, &offset, &frames, &flags, &truncated) &&
                     PyUnicode_FSConverter(path_obj, &path_bytes);
            if (ok && frames) {
                int rc;
                const char *path = PyBytes_AS_STRING(path_bytes);
                Py_BEGIN_ALLOW_THREADS
                rc = b200flac_stream_write_file(stream, path, offset, frames, flags);
                Py_END_ALLOW_THREADS
                if (rc) {
                    /* rc 2: the file ended early -- the IOError WaveReader.read() raises (wav.py:516-518) */
                    PyErr_SetString(PyExc_IOError, rc == 2 ? truncated : b200flac_last_error());
                    Py_DECREF(path_bytes); Py_DECREF(span); Py_DECREF(inner);
                    goto error;
                }
                PyObject *r = PyObject_CallMethod(inner, "b200_file_span_consumed", "K", frames);
                if (!r) { Py_DECREF(path_bytes); Py_DECREF(span); Py_DECREF(inner); goto error; }
                Py_DECREF(r);
            }
            Py_XDECREF(path_bytes);
            Py_DECREF(span);
        }
        Py_XDECREF(inner);
        PyErr_Clear(); /* a reader without a usable span simply takes the FrameList path */
    }
    const unsigned read_frames = p.block_size * read_blocks;
    for (;;) {
        /* pcmreader->read(block_size): src/pcmconv.c:236, with the exact-type check of :244 */
        PyObject *fl_obj = PyObject_CallMethod(reader->obj, "read", "i", (int)read_frames);
        if (!fl_obj) goto error;
        if ((PyObject *)Py_TYPE(fl_obj) != reader->framelist_type) {
            Py_DECREF(fl_obj);
            PyErr_SetString(PyExc_TypeError, "results from pcmreader.read() must be FrameLists");
            goto error;
        }
        pcm_FrameList *fl = (pcm_FrameList *)fl_obj;
        const unsigned frames = fl->frames;
        if (frames == 0) { Py_DECREF(fl_obj); break; } /* flac.c:247 */
        if (fl->channels != p.channels || fl->bits_per_sample != p.bits_per_sample) {
            Py_DECREF(fl_obj);
            PyErr_SetString(PyExc_ValueError, "FrameList does not match the pcmreader's channels / bits_per_sample");
            goto error;
        }
        const size_t need = (size_t)fl->samples_length * bytes_ps;
        if (need > packed_cap) {
            uint8_t *np_ = (uint8_t *)realloc(packed, need);
            if (!np_) { Py_DECREF(fl_obj); PyErr_NoMemory(); goto error; }
            packed = np_; packed_cap = need;
        }
        int rc;
        Py_BEGIN_ALLOW_THREADS /* flac.c:255-269 drops the GIL around the encode as well */
        pack_le_signed(fl, packed);
        rc = b200flac_stream_write(stream, packed, frames);
        /* the reference makes one frame of whatever read() returned (H12): a short read mid-stream
           becomes a short frame, so close the block here */
        if (!rc && frames != read_frames) rc = b200flac_stream_end_block(stream);
        Py_END_ALLOW_THREADS
        Py_DECREF(fl_obj);
        if (rc) { PyErr_SetString(PyExc_IOError, b200flac_last_error()); goto error; }
    }

    {
        uint64_t *offs = NULL, n = 0;
        uint32_t *lens = NULL;
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = b200flac_stream_close(stream, 0, &offs, &lens, &n);
        Py_END_ALLOW_THREADS
        stream = NULL;
        if (rc) { PyErr_SetString(PyExc_IOError, b200flac_last_error()); goto error; }
        /* flac.c:249-253: list of (byte offset from the first frame, PCM frames) */
        offsets = PyList_New((Py_ssize_t)n);
        for (uint64_t i = 0; offsets && i < n; i++) {
            PyObject *t = Py_BuildValue("(K, I)", (unsigned long long)offs[i], (unsigned)lens[i]);
            if (!t) { Py_CLEAR(offsets); break; }
            PyList_SET_ITEM(offsets, (Py_ssize_t)i, t);
        }
        b200flac_free(offs);
        b200flac_free(lens);
        if (!offsets) goto error;
    }
    free(packed);
    /* success: close() then del, flac.c:282-283 */
    {
        PyObject *r = PyObject_CallMethod(reader->obj, "close", NULL);
        if (!r) { Py_DECREF(offsets); pcmreader_del(reader); return NULL; }
        Py_DECREF(r);
    }
    pcmreader_del(reader);
    return offsets;

error:
    /* flac.c:288-296: no close() on the reader, the partial file is left for from_pcm to unlink */
    if (stream) {
        PyObject *et, *ev, *tb;
        PyErr_Fetch(&et, &ev, &tb);
        Py_BEGIN_ALLOW_THREADS
        b200flac_stream_close(stream, 1, NULL, NULL, NULL);
        Py_END_ALLOW_THREADS
        PyErr_Restore(et, ev, tb);
    }
    free(packed);
    pcmreader_del(reader);
    return NULL;
}

/* encode_flac_files(filenames, pcmreaders, block_size, max_lpc_order, min_residual_partition_order,
 *                   max_residual_partition_order, mid_side=0, ..., padding_size=4096, host_threads=0) -> None
 * Many tracks in one call (b200flac_encode_files; no reference counterpart -- the reference
 * encodes one file per encode_flac call).  Same option names and defaults as encode_flac.  Every reader is read to
 * its end and cut into block_size frames -- what encode_flac does with the BufferedPCMReader FlacAudio.from_pcm
 * always passes -- and every file is byte for byte what encode_flac writes for that reader.  All readers must share
 * sample rate, channels and bits per sample.  Readers are closed on success, like encode_flac's. */
static PyObject *encoders_encode_flac_files(PyObject *dummy, PyObject *args, PyObject *keywds)
{
    static char *kwlist[] = {"filenames", "pcmreaders", "block_size", "max_lpc_order",
                             "min_residual_partition_order", "max_residual_partition_order",
                             "mid_side", "adaptive_mid_side", "exhaustive_model_search",
                             "disable_verbatim_subframes", "disable_constant_subframes",
                             "disable_fixed_subframes", "disable_lpc_subframes", "padding_size", "host_threads", NULL};
    PyObject *names_obj, *readers_obj, *names = NULL, *readers = NULL, *result = NULL;
    b200flac_params p;
    unsigned padding_size = 4096;
    int host_threads = 0;
    Py_ssize_t n = 0;
    struct py_pcmreader **rd = NULL;
    PyObject **name_bytes = NULL;
    const char **c_names = NULL;
    uint8_t **bufs = NULL;
    size_t *caps = NULL, *used = NULL;
    uint64_t *frames = NULL;
    uint8_t *arena = NULL;
    const uint8_t **ptrs = NULL;

    memset(&p, 0, sizeof(p));
    if (!PyArg_ParseTupleAndKeywords(args, keywds, "OOIIII|iiiiiiiIi", kwlist, &names_obj, &readers_obj,
                                     &p.block_size, &p.max_lpc_order, &p.min_residual_partition_order,
                                     &p.max_residual_partition_order, &p.mid_side, &p.adaptive_mid_side,
                                     &p.exhaustive_model_search, &p.no_verbatim_subframes, &p.no_constant_subframes,
                                     &p.no_fixed_subframes, &p.no_lpc_subframes, &padding_size, &host_threads))
        return NULL;
    names = PySequence_Fast(names_obj, "filenames must be a sequence");
    readers = names ? PySequence_Fast(readers_obj, "pcmreaders must be a sequence") : NULL;
    if (!names || !readers) goto done;
    n = PySequence_Fast_GET_SIZE(names);
    if (n != PySequence_Fast_GET_SIZE(readers)) { PyErr_SetString(PyExc_ValueError, "one pcmreader per filename"); goto done; }
    if (n == 0) { result = Py_None; Py_INCREF(result); goto done; }
    if (p.block_size == 0 || p.block_size > (1u << 20)) { PyErr_SetString(PyExc_ValueError, "block_size must be 1..1048576"); goto done; }
    rd = (struct py_pcmreader **)calloc((size_t)n, sizeof(*rd));
    name_bytes = (PyObject **)calloc((size_t)n, sizeof(*name_bytes));
    c_names = (const char **)calloc((size_t)n, sizeof(*c_names));
    bufs = (uint8_t **)calloc((size_t)n, sizeof(*bufs));
    caps = (size_t *)calloc((size_t)n, sizeof(*caps));
    used = (size_t *)calloc((size_t)n, sizeof(*used));
    frames = (uint64_t *)calloc((size_t)n, sizeof(*frames));
    ptrs = (const uint8_t **)calloc((size_t)n, sizeof(*ptrs));
    if (!rd || !name_bytes || !c_names || !bufs || !caps || !used || !frames || !ptrs) { PyErr_NoMemory(); goto done; }
    for (Py_ssize_t i = 0; i < n; i++) {
        if (!PyUnicode_FSConverter(PySequence_Fast_GET_ITEM(names, i), &name_bytes[i])) goto done;
        c_names[i] = PyBytes_AS_STRING(name_bytes[i]);
        void *r = NULL;
        if (!pcmreader_converter(PySequence_Fast_GET_ITEM(readers, i), &r)) goto done;
        rd[i] = (struct py_pcmreader *)r;
        if (i == 0) {
            p.sample_rate = rd[0]->sample_rate; p.channels = rd[0]->channels; p.bits_per_sample = rd[0]->bits_per_sample;
        } else if (rd[i]->sample_rate != p.sample_rate || rd[i]->channels != p.channels || rd[i]->bits_per_sample != p.bits_per_sample) {
            PyErr_SetString(PyExc_ValueError, "all pcmreaders of one call must share sample rate, channels and bits per sample");
            goto done;
        }
        /* (encode_flac's IOError for a file that cannot be opened, flac.c:114-117) */
        FILE *probe = fopen(c_names[i], "wb");
        if (!probe) { PyErr_SetFromErrnoWithFilename(PyExc_IOError, c_names[i]); goto done; }
        fclose(probe);
    }
    {
        const unsigned bytes_ps = p.bits_per_sample / 8;
        if (bytes_ps < 1 || bytes_ps > 3) { PyErr_SetString(PyExc_ValueError, "unsupported bits_per_sample"); goto done; }
        const unsigned read_frames = p.block_size * 64;
        size_t total = 0;
        for (Py_ssize_t i = 0; i < n; i++) {
            for (;;) {
                PyObject *fl_obj = PyObject_CallMethod(rd[i]->obj, "read", "i", (int)read_frames);
                if (!fl_obj) goto done;
                if ((PyObject *)Py_TYPE(fl_obj) != rd[i]->framelist_type) {
                    Py_DECREF(fl_obj);
                    PyErr_SetString(PyExc_TypeError, "results from pcmreader.read() must be FrameLists");
                    goto done;
                }
                pcm_FrameList *fl = (pcm_FrameList *)fl_obj;
                if (fl->frames == 0) { Py_DECREF(fl_obj); break; }
                if (fl->channels != p.channels || fl->bits_per_sample != p.bits_per_sample) {
                    Py_DECREF(fl_obj);
                    PyErr_SetString(PyExc_ValueError, "FrameList does not match the pcmreader's channels / bits_per_sample");
                    goto done;
                }
                const size_t need = used[i] + (size_t)fl->samples_length * bytes_ps;
                if (need > caps[i]) {
                    size_t ncap = caps[i] ? caps[i] * 2 : (1u << 22);
                    while (ncap < need) ncap *= 2;
                    uint8_t *np_ = (uint8_t *)realloc(bufs[i], ncap);
                    if (!np_) { Py_DECREF(fl_obj); PyErr_NoMemory(); goto done; }
                    bufs[i] = np_; caps[i] = ncap;
                }
                pack_le_signed(fl, bufs[i] + used[i]);
                used[i] = need;
                frames[i] += fl->frames;
                Py_DECREF(fl_obj);
            }
            total += (used[i] + 63) & ~(size_t)63;
        }
        /* one page-locked arena for all tracks: the engine copies to the device straight from it */
        arena = (uint8_t *)b200flac_host_alloc(total ? total : 64);
        if (!arena) { PyErr_SetString(PyExc_IOError, b200flac_last_error()); goto done; }
        size_t off = 0;
        for (Py_ssize_t i = 0; i < n; i++) {
            if (used[i]) memcpy(arena + off, bufs[i], used[i]);
            ptrs[i] = arena + off;
            off += (used[i] + 63) & ~(size_t)63;
            free(bufs[i]); bufs[i] = NULL;
        }
        int rc;
        Py_BEGIN_ALLOW_THREADS
        rc = b200flac_encode_files((uint32_t)n, c_names, &p, padding_size, NULL, ptrs, frames, -1, host_threads);
        Py_END_ALLOW_THREADS
        if (rc) {
            const char *msg = b200flac_last_error();
            const int bad_param = strstr(msg, "must be") != NULL || strstr(msg, "unsupported") != NULL;
            PyErr_SetString(bad_param ? PyExc_ValueError : PyExc_IOError, msg);
            goto done;
        }
    }
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *r = PyObject_CallMethod(rd[i]->obj, "close", NULL);
        if (!r) goto done;
        Py_DECREF(r);
    }
    result = Py_None;
    Py_INCREF(result);
done:
    if (arena) b200flac_host_free(arena);
    for (Py_ssize_t i = 0; i < n; i++) {
        if (rd && rd[i]) pcmreader_del(rd[i]);
        if (name_bytes) Py_XDECREF(name_bytes[i]);
        if (bufs) free(bufs[i]);
    }
    free(rd); free(name_bytes); free((void *)c_names); free(bufs); free(caps); free(used); free(frames); free((void *)ptrs);
    Py_XDECREF(names);
    Py_XDECREF(readers);
    return result;
}

/* encode_tta(file, pcmreader) -> [frame size, ...]      src/encoders/tta.c:31-117
 * Same keyword names ("file", "pcmreader", tta.c:44-46), same protocol: block_size = sample_rate * 256 / 245
 * PCM frames are asked of the reader per call and WHATEVER comes back is one TTA frame (tta.c:69-83); the frames
 * are written to the file object (any object with write(); the reference needs a concrete py2 file) and the
 * list of their sizes in bytes is returned -- what TrueAudio.from_pcm builds the seektable from.  The reader is
 * not closed (the reference only drops its reference).  All frames are encoded in one GPU batch at the end. */
static PyObject *encoders_encode_tta(PyObject *dummy, PyObject *args, PyObject *keywds)
{
    static char *kwlist[] = {"file", "pcmreader", NULL};
    PyObject *file_obj;
    struct py_pcmreader *reader = NULL;
    uint8_t *pcm = NULL, *out = NULL;
    size_t cap = 0, used = 0;
    uint32_t *lengths = NULL, *sizes = NULL;
    size_t n_len = 0, len_cap = 0;
    uint64_t total_frames = 0, out_bytes = 0;
    uint32_t n_frames = 0;
    PyObject *result = NULL;
    b200tta_params p;

    if (!PyArg_ParseTupleAndKeywords(args, keywds, "OO&", kwlist, &file_obj, pcmreader_converter, &reader)) return NULL;
    p.sample_rate = reader->sample_rate;
    p.channels = reader->channels;
    p.bits_per_sample = reader->bits_per_sample;
    const unsigned bytes_ps = p.bits_per_sample / 8;
    const unsigned block_size = b200tta_block_size(p.sample_rate);
    if (block_size == 0 || (bytes_ps != 1 && bytes_ps != 2 && bytes_ps != 3)) {
        PyErr_SetString(PyExc_ValueError, "unsupported sample rate or bits_per_sample");
        goto done;
    }
    for (;;) {
        PyObject *fl_obj = PyObject_CallMethod(reader->obj, "read", "i", (int)block_size);
        if (!fl_obj) goto done;
        if ((PyObject *)Py_TYPE(fl_obj) != reader->framelist_type) {
            Py_DECREF(fl_obj);
            PyErr_SetString(PyExc_TypeError, "results from pcmreader.read() must be FrameLists");
            goto done;
        }
        pcm_FrameList *fl = (pcm_FrameList *)fl_obj;
        if (fl->frames == 0) { Py_DECREF(fl_obj); break; }   /* tta.c:71 */
        if (fl->channels != p.channels || fl->bits_per_sample != p.bits_per_sample) {
            Py_DECREF(fl_obj);
            PyErr_SetString(PyExc_ValueError, "FrameList does not match the pcmreader's channels / bits_per_sample");
            goto done;
        }
        const size_t need = used + (size_t)fl->samples_length * bytes_ps;
        if (need > cap) {
            size_t ncap = cap ? cap * 2 : (1u << 22);
            while (ncap < need) ncap *= 2;
            uint8_t *np_ = (uint8_t *)realloc(pcm, ncap);
            if (!np_) { Py_DECREF(fl_obj); PyErr_NoMemory(); goto done; }
            pcm = np_; cap = ncap;
        }
        if (n_len == len_cap) {
            len_cap = len_cap ? len_cap * 2 : 1024;
            uint32_t *nl = (uint32_t *)realloc(lengths, len_cap * sizeof(uint32_t));
            if (!nl) { Py_DECREF(fl_obj); PyErr_NoMemory(); goto done; }
            lengths = nl;
        }
        pack_le_signed(fl, pcm + used);
        used = need;
        lengths[n_len++] = fl->frames;
        total_frames += fl->frames;
        Py_DECREF(fl_obj);
    }
    {
        int rc;
        Py_BEGIN_ALLOW_THREADS     /* tta.c:73-79 drops the GIL around encode_frame as well */
        rc = b200tta_encode_frames(&p, pcm, total_frames, lengths, (uint32_t)n_len, 0, &out, &out_bytes, &sizes, &n_frames, NULL);
        Py_END_ALLOW_THREADS
        if (rc) { PyErr_SetString(PyExc_IOError, b200tta_last_error()); goto done; }
    }
    if (out_bytes) {
        PyObject *r = PyObject_CallMethod(file_obj, "write", "y#", (const char *)out, (Py_ssize_t)out_bytes);
        if (!r) goto done;
        Py_DECREF(r);
    }
    result = PyList_New((Py_ssize_t)n_frames);
    for (uint32_t i = 0; result && i < n_frames; i++) {
        PyObject *v = PyLong_FromUnsignedLong(sizes[i]);
        if (!v) { Py_CLEAR(result); break; }
        PyList_SET_ITEM(result, (Py_ssize_t)i, v);
    }
done:
    free(pcm);
    free(lengths);
    b200tta_free(out);
    b200tta_free(sizes);
    pcmreader_del(reader);
    return result;
}

/* encode_alac(file, pcmreader, block_size, initial_history, history_multiplier, maximum_k,
 *             minimum_interlacing_leftweight=0, maximum_interlacing_leftweight=4) -> ([frameset byte sizes], total PCM frames)
 * src/encoders/alac.c:30-214: same keyword names (:34-42) and format "OO&iiii|ii", same protocol: an 8-byte mdat
 * header (size placeholder, "mdat"), one frameset per pcmreader.read(block_size) -- whatever length comes back --,
 * then the header's size rewritten (the reference uses fgetpos/fsetpos on the FILE*; here the file object's
 * tell()/seek()), and alac_log_output's tuple returned (:1289-1325).  The reader is not closed. */
static PyObject *encoders_encode_alac(PyObject *dummy, PyObject *args, PyObject *keywds)
{
    static char *kwlist[] = {"file", "pcmreader", "block_size", "initial_history", "history_multiplier", "maximum_k",
                             "minimum_interlacing_leftweight", "maximum_interlacing_leftweight", NULL};
    PyObject *file_obj;
    struct py_pcmreader *reader = NULL;
    int block_size, initial_history, history_multiplier, maximum_k, min_lw = 0, max_lw = 4;
    uint8_t *pcm = NULL, *out = NULL;
    size_t cap = 0, used = 0;
    uint32_t *lengths = NULL, *sizes = NULL;
    size_t n_len = 0, len_cap = 0;
    uint64_t total_frames = 0, out_bytes = 0;
    uint32_t n_frames = 0;
    PyObject *result = NULL, *list = NULL, *start = NULL, *r = NULL;
    b200alac_params p;

    if (!PyArg_ParseTupleAndKeywords(args, keywds, "OO&iiii|ii", kwlist, &file_obj, pcmreader_converter, &reader,
                                     &block_size, &initial_history, &history_multiplier, &maximum_k, &min_lw, &max_lw))
        return NULL;
    if (reader->bits_per_sample != 16 && reader->bits_per_sample != 24) {       /* alac.c:76-80 */
        PyErr_SetString(PyExc_ValueError, "bits per sample must be 16 or 24");
        goto done;
    }
    if (block_size <= 0) { PyErr_SetString(PyExc_ValueError, "block_size must be positive"); goto done; }
    p.channels = reader->channels;
    p.bits_per_sample = reader->bits_per_sample;
    p.block_size = (uint32_t)block_size;
    p.initial_history = (uint32_t)initial_history;
    p.history_multiplier = (uint32_t)history_multiplier;
    p.maximum_k = (uint32_t)maximum_k;
    p.minimum_interlacing_leftweight = (uint32_t)min_lw;
    p.maximum_interlacing_leftweight = (uint32_t)max_lw;
    const unsigned bytes_ps = p.bits_per_sample / 8;
    for (;;) {
        PyObject *fl_obj = PyObject_CallMethod(reader->obj, "read", "i", block_size);
        if (!fl_obj) goto done;
        if ((PyObject *)Py_TYPE(fl_obj) != reader->framelist_type) {
            Py_DECREF(fl_obj);
            PyErr_SetString(PyExc_TypeError, "results from pcmreader.read() must be FrameLists");
            goto done;
        }
        pcm_FrameList *fl = (pcm_FrameList *)fl_obj;
        if (fl->frames == 0) { Py_DECREF(fl_obj); break; }   /* alac.c:165 */
        if (fl->channels != p.channels || fl->bits_per_sample != p.bits_per_sample) {
            Py_DECREF(fl_obj);
            PyErr_SetString(PyExc_ValueError, "FrameList does not match the pcmreader's channels / bits_per_sample");
            goto done;
        }
        const size_t need = used + (size_t)fl->samples_length * bytes_ps;
        if (need > cap) {
            size_t ncap = cap ? cap * 2 : (1u << 22);
            while (ncap < need) ncap *= 2;
            uint8_t *np_ = (uint8_t *)realloc(pcm, ncap);
            if (!np_) { Py_DECREF(fl_obj); PyErr_NoMemory(); goto done; }
            pcm = np_; cap = ncap;
        }
        if (n_len == len_cap) {
            len_cap = len_cap ? len_cap * 2 : 1024;
            uint32_t *nl = (uint32_t *)realloc(lengths, len_cap * sizeof(uint32_t));
            if (!nl) { Py_DECREF(fl_obj); PyErr_NoMemory(); goto done; }
            lengths = nl;
        }
        pack_le_signed(fl, pcm + used);
        used = need;
        lengths[n_len++] = fl->frames;
        total_frames += fl->frames;
        Py_DECREF(fl_obj);
    }
    {
        int rc;
        Py_BEGIN_ALLOW_THREADS     /* alac.c:166-181 drops the GIL around write_frameset as well */
        rc = b200alac_encode_framesets(&p, pcm, total_frames, lengths, (uint32_t)n_len, 0, &out, &out_bytes, &sizes, &n_frames, NULL);
        Py_END_ALLOW_THREADS
        if (rc) {
            const char *msg = b200alac_last_error();
            PyErr_SetString(strstr(msg, "must be") || strstr(msg, "unsupported") ? PyExc_ValueError : PyExc_IOError, msg);
            goto done;
        }
    }
    {
        /* placeholder header, framesets, then the size at the header's position (alac.c:153-157, 185-189) */
        const uint32_t size = (uint32_t)(out_bytes + 8);
        const char head[8] = {(char)(size >> 24), (char)(size >> 16), (char)(size >> 8), (char)size, 'm', 'd', 'a', 't'};
        if ((start = PyObject_CallMethod(file_obj, "tell", NULL)) == NULL) goto done;
        if ((r = PyObject_CallMethod(file_obj, "write", "y#", head, (Py_ssize_t)8)) == NULL) goto done;
        Py_CLEAR(r);
        if (out_bytes) {
            if ((r = PyObject_CallMethod(file_obj, "write", "y#", (const char *)out, (Py_ssize_t)out_bytes)) == NULL) goto done;
            Py_CLEAR(r);
        }
        /* (the reference leaves the file position just after the rewritten size, alac.c:187-189) */
        if ((r = PyObject_CallMethod(file_obj, "seek", "O", start)) == NULL) goto done;
        Py_CLEAR(r);
        if ((r = PyObject_CallMethod(file_obj, "write", "y#", head, (Py_ssize_t)4)) == NULL) goto done;
        Py_CLEAR(r);
    }
    list = PyList_New((Py_ssize_t)n_frames);
    for (uint32_t i = 0; list && i < n_frames; i++) {
        PyObject *v = PyLong_FromUnsignedLong(sizes[i]);
        if (!v) { Py_CLEAR(list); break; }
        PyList_SET_ITEM(list, (Py_ssize_t)i, v);
    }
    if (list) result = Py_BuildValue("(O,K)", list, (unsigned long long)total_frames);   /* alac.c:1318-1320 */
done:
    Py_XDECREF(list);
    Py_XDECREF(start);
    Py_XDECREF(r);
    free(pcm);
    free(lengths);
    b200alac_free(out);
    b200alac_free(sizes);
    pcmreader_del(reader);
    return result;
}

static PyObject *encoders_device_count(PyObject *dummy, PyObject *args)
{
    return PyLong_FromLong(b200flac_device_count());
}

/* finalize_flac_metadata(filename, offsets, seekpoint_interval=0, channel_mask=0)
 * the tail of FlacAudio.from_pcm (audiotools/flac.py:1811-1832) in C: see b200flac_finalize_metadata */
static PyObject *encoders_finalize_flac_metadata(PyObject *dummy, PyObject *args, PyObject *keywds)
{
    static char *kwlist[] = {"filename", "offsets", "seekpoint_interval", "channel_mask", NULL};
    const char *filename;
    PyObject *offsets, *seq;
    unsigned interval = 0, mask = 0;
    Py_ssize_t n, i;
    uint64_t *offs;
    uint32_t *lens;
    int rc;
    if (!PyArg_ParseTupleAndKeywords(args, keywds, "sO|II", kwlist, &filename, &offsets, &interval, &mask)) return NULL;
    if ((seq = PySequence_Fast(offsets, "offsets must be a sequence of (byte_offset, pcm_frames)")) == NULL) return NULL;
    n = PySequence_Fast_GET_SIZE(seq);
    offs = malloc(sizeof(uint64_t) * (size_t)(n ? n : 1));
    lens = malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
    if (!offs || !lens) { free(offs); free(lens); Py_DECREF(seq); return PyErr_NoMemory(); }
    for (i = 0; i < n; i++) {
        unsigned long long o;
        unsigned int f;
        if (!PyArg_ParseTuple(PySequence_Fast_GET_ITEM(seq, i), "KI", &o, &f)) {
            free(offs); free(lens); Py_DECREF(seq);
            return NULL;
        }
        offs[i] = o;
        lens[i] = f;
    }
    Py_DECREF(seq);
    Py_BEGIN_ALLOW_THREADS
    rc = b200flac_finalize_metadata(filename, offs, lens, (uint64_t)n, interval, mask);
    Py_END_ALLOW_THREADS
    free(offs);
    free(lens);
    if (rc) { PyErr_SetString(PyExc_IOError, b200flac_last_error()); return NULL; }
    Py_RETURN_NONE;
}

static PyMethodDef module_methods[] = {
    {"encode_flac", (PyCFunction)encoders_encode_flac, METH_VARARGS | METH_KEYWORDS,
     "encode_flac(filename, pcmreader, block_size, max_lpc_order, min_residual_partition_order, "
     "max_residual_partition_order, mid_side=0, adaptive_mid_side=0, exhaustive_model_search=0, "
     "disable_verbatim_subframes=0, disable_constant_subframes=0, disable_fixed_subframes=0, "
     "disable_lpc_subframes=0, padding_size=4096) -> [(byte_offset, pcm_frames), ...]"},
    {"encode_flac_files", (PyCFunction)encoders_encode_flac_files, METH_VARARGS | METH_KEYWORDS,
     "encode_flac_files(filenames, pcmreaders, block_size, max_lpc_order, min_residual_partition_order, "
     "max_residual_partition_order, ...) -- many tracks in one call; every file equals encode_flac's"},
    {"encode_tta", (PyCFunction)encoders_encode_tta, METH_VARARGS | METH_KEYWORDS,
     "encode_tta(file, pcmreader) -> [frame_size, ...]: TTA frames written to file (src/encoders/tta.c:31-117)"},
    {"encode_alac", (PyCFunction)encoders_encode_alac, METH_VARARGS | METH_KEYWORDS,
     "encode_alac(file, pcmreader, block_size, initial_history, history_multiplier, maximum_k, "
     "minimum_interlacing_leftweight=0, maximum_interlacing_leftweight=4) -> ([frameset sizes], total_pcm_frames): "
     "the mdat atom written to file (src/encoders/alac.c:30-214)"},
    {"finalize_flac_metadata", (PyCFunction)encoders_finalize_flac_metadata, METH_VARARGS | METH_KEYWORDS,
     "finalize_flac_metadata(filename, offsets, seekpoint_interval=0, channel_mask=0): SEEKTABLE from the "
     "encoder's offsets, channel-mask tag, PADDING adjustment -- FlacAudio.from_pcm's tail in C"},
    {"b200_device_count", (PyCFunction)encoders_device_count, METH_NOARGS, "usable CUDA devices"},
    {NULL}};

static struct PyModuleDef encoders_module = {PyModuleDef_HEAD_INIT, "audiotools.encoders",
                                             "Low-level audio format encoders (B200 FLAC engine)", -1, module_methods};

PyMODINIT_FUNC PyInit_encoders(void) { return PyModule_Create(&encoders_module); }
