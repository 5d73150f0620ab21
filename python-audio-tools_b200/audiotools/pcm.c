/* audiotools.pcm -- CPython 3 boundary type for the B200 FLAC engine.
 *
 * The reference's C encoders take PCM from Python as `audiotools.pcm.FrameList`
 * objects (src/pcm.h:40-54): interleaved `int` samples plus frames / channels /
 * bits_per_sample, with an exact type-identity check on the way in
 * (src/pcmconv.c:244).  The reference's module is Python-2 only
 * (Py_InitModule3, PyString_*), so this file provides the same type -- same C
 * layout, same constructor and the methods the FLAC-encode path and
 * BufferedPCMReader use (from_list, split, +, to_bytes, frame, channel) -- for
 * CPython 3.  It is boundary scaffolding, not a port of the reference's generic
 * PCM container (FloatFrameList, converters etc. are out of scope).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>
#include "pcm.h"

static PyTypeObject pcm_FrameListType;

static pcm_FrameList *framelist_alloc(unsigned frames, unsigned channels, unsigned bps)
{
    pcm_FrameList *f = (pcm_FrameList *)pcm_FrameListType.tp_alloc(&pcm_FrameListType, 0);
    if (!f) return NULL;
    f->frames = frames;
    f->channels = channels;
    f->bits_per_sample = bps;
    f->samples_length = frames * channels;
    f->samples = (int *)malloc(sizeof(int) * (f->samples_length ? f->samples_length : 1));
    if (!f->samples) { Py_DECREF(f); PyErr_NoMemory(); return NULL; }
    return f;
}

static void FrameList_dealloc(pcm_FrameList *self)
{
    free(self->samples);
    Py_TYPE(self)->tp_free((PyObject *)self);
}

static PyObject *FrameList_new(PyTypeObject *type, PyObject *args, PyObject *kwds)
{
    pcm_FrameList *self = (pcm_FrameList *)type->tp_alloc(type, 0);
    if (self) { self->samples = NULL; self->frames = self->channels = self->bits_per_sample = 0; self->samples_length = 0; }
    return (PyObject *)self;
}

/* FrameList(data, channels, bits_per_sample, is_big_endian, is_signed), src/pcm.c FrameList_init */
static int FrameList_init(pcm_FrameList *self, PyObject *args, PyObject *kwds)
{
    Py_buffer data;
    int channels, bps, big_endian, is_signed;
    if (!PyArg_ParseTuple(args, "y*iiii", &data, &channels, &bps, &big_endian, &is_signed)) return -1;
    if (channels < 1) { PyBuffer_Release(&data); PyErr_SetString(PyExc_ValueError, "number of channels must be > 0"); return -1; }
    if (bps != 8 && bps != 16 && bps != 24) {
        PyBuffer_Release(&data); PyErr_SetString(PyExc_ValueError, "bits_per_sample must be 8, 16 or 24"); return -1;
    }
    const unsigned bytes = (unsigned)bps / 8;
    if (data.len % (Py_ssize_t)(bytes * (unsigned)channels)) {
        PyBuffer_Release(&data);
        PyErr_SetString(PyExc_ValueError, "number of samples must be divisible by bits-per-sample and number of channels");
        return -1;
    }
    const size_t n = (size_t)data.len / bytes;
    free(self->samples);
    self->samples = (int *)malloc(sizeof(int) * (n ? n : 1));
    if (!self->samples) { PyBuffer_Release(&data); PyErr_NoMemory(); return -1; }
    const uint8_t *p = (const uint8_t *)data.buf;
    const size_t n_all = n;
    size_t first = 0;
    if (bytes == 2 && !big_endian && is_signed) {
        /* the common case (16-bit WAVE data), in a form the compiler vectorises */
        int *restrict dst = self->samples;
        for (size_t i = 0; i < n; i++) { int16_t v; memcpy(&v, p + 2 * i, 2); dst[i] = v; }
        first = n;
    }
    p += first * bytes;
    for (size_t i = first; i < n; i++, p += bytes) {
        uint32_t v = 0;
        for (unsigned b = 0; b < bytes; b++) v |= (uint32_t)p[big_endian ? (bytes - 1 - b) : b] << (8 * b);
        if (is_signed) {
            if (v & (1u << (bps - 1))) v |= ~((1u << bps) - 1u);
            self->samples[i] = (int)v;
        } else {
            self->samples[i] = (int)v - (1 << (bps - 1));
        }
    }
    self->samples_length = (unsigned)n_all;
    self->channels = (unsigned)channels;
    self->bits_per_sample = (unsigned)bps;
    self->frames = (unsigned)(n_all / (unsigned)channels);
    PyBuffer_Release(&data);
    return 0;
}

static Py_ssize_t FrameList_len(pcm_FrameList *o) { return (Py_ssize_t)o->samples_length; }

static PyObject *FrameList_item(pcm_FrameList *o, Py_ssize_t i)
{
    if (i < 0 || i >= (Py_ssize_t)o->samples_length) { PyErr_SetString(PyExc_IndexError, "index out of range"); return NULL; }
    return PyLong_FromLong(o->samples[i]);
}

static int same_shape(pcm_FrameList *a, PyObject *b)
{
    if (Py_TYPE(b) != &pcm_FrameListType) { PyErr_SetString(PyExc_TypeError, "can only concatenate FrameList with other FrameLists"); return 0; }
    pcm_FrameList *o = (pcm_FrameList *)b;
    if (a->channels != o->channels) { PyErr_SetString(PyExc_ValueError, "both FrameLists must have the same number of channels"); return 0; }
    if (a->bits_per_sample != o->bits_per_sample) { PyErr_SetString(PyExc_ValueError, "both FrameLists must have the same number of bits per sample"); return 0; }
    return 1;
}

static PyObject *FrameList_concat(pcm_FrameList *a, PyObject *bb)
{
    if (!same_shape(a, bb)) return NULL;
    pcm_FrameList *b = (pcm_FrameList *)bb;
    /* FrameLists never change after construction, so an empty operand needs no copy (BufferedPCMReader's
       `buffer += frame` with an empty buffer, audiotools/__init__.py:2585-2590) */
    if (a->frames == 0) { Py_INCREF(bb); return bb; }
    if (b->frames == 0) { Py_INCREF(a); return (PyObject *)a; }
    pcm_FrameList *r = framelist_alloc(a->frames + b->frames, a->channels, a->bits_per_sample);
    if (!r) return NULL;
    memcpy(r->samples, a->samples, sizeof(int) * a->samples_length);
    memcpy(r->samples + a->samples_length, b->samples, sizeof(int) * b->samples_length);
    return (PyObject *)r;
}

static PyObject *FrameList_split(pcm_FrameList *self, PyObject *args)
{
    int n;
    if (!PyArg_ParseTuple(args, "i", &n)) return NULL;
    if (n < 0) { PyErr_SetString(PyExc_IndexError, "split point must be positive"); return NULL; }
    unsigned head = (unsigned)n < self->frames ? (unsigned)n : self->frames;
    if (head == self->frames || head == 0) {
        /* one side is the whole list (immutable: shared, not copied), the other is empty */
        pcm_FrameList *e = framelist_alloc(0, self->channels, self->bits_per_sample);
        if (!e) return NULL;
        Py_INCREF(self);
        return head ? Py_BuildValue("(NN)", self, e) : Py_BuildValue("(NN)", e, self);
    }
    pcm_FrameList *h = framelist_alloc(head, self->channels, self->bits_per_sample);
    if (!h) return NULL;
    pcm_FrameList *t = framelist_alloc(self->frames - head, self->channels, self->bits_per_sample);
    if (!t) { Py_DECREF(h); return NULL; }
    memcpy(h->samples, self->samples, sizeof(int) * h->samples_length);
    memcpy(t->samples, self->samples + h->samples_length, sizeof(int) * t->samples_length);
    return Py_BuildValue("(NN)", h, t);
}

/* packs the samples as bits_per_sample/8-byte integers (to_bytes, src/pcm.c) */
static void pcm_framelist_pack(const pcm_FrameList *f, int big_endian, int is_signed, uint8_t *out)
{
    const unsigned bytes = f->bits_per_sample / 8;
    for (unsigned i = 0; i < f->samples_length; i++, out += bytes) {
        uint32_t v = (uint32_t)f->samples[i];
        if (!is_signed) v += 1u << (f->bits_per_sample - 1);
        for (unsigned b = 0; b < bytes; b++) out[big_endian ? (bytes - 1 - b) : b] = (uint8_t)(v >> (8 * b));
    }
}

static PyObject *FrameList_to_bytes(pcm_FrameList *self, PyObject *args)
{
    int big_endian, is_signed;
    if (!PyArg_ParseTuple(args, "ii", &big_endian, &is_signed)) return NULL;
    const Py_ssize_t n = (Py_ssize_t)self->samples_length * (self->bits_per_sample / 8);
    PyObject *b = PyBytes_FromStringAndSize(NULL, n);
    if (!b) return NULL;
    pcm_framelist_pack(self, big_endian, is_signed, (uint8_t *)PyBytes_AS_STRING(b));
    return b;
}

static PyObject *FrameList_frame(pcm_FrameList *self, PyObject *args)
{
    int i;
    if (!PyArg_ParseTuple(args, "i", &i)) return NULL;
    if (i < 0 || (unsigned)i >= self->frames) { PyErr_SetString(PyExc_IndexError, "frame number out of range"); return NULL; }
    pcm_FrameList *r = framelist_alloc(1, self->channels, self->bits_per_sample);
    if (!r) return NULL;
    memcpy(r->samples, self->samples + (size_t)i * self->channels, sizeof(int) * self->channels);
    return (PyObject *)r;
}

static PyObject *FrameList_channel(pcm_FrameList *self, PyObject *args)
{
    int c;
    if (!PyArg_ParseTuple(args, "i", &c)) return NULL;
    if (c < 0 || (unsigned)c >= self->channels) { PyErr_SetString(PyExc_IndexError, "channel number out of range"); return NULL; }
    pcm_FrameList *r = framelist_alloc(self->frames, 1, self->bits_per_sample);
    if (!r) return NULL;
    for (unsigned i = 0; i < self->frames; i++) r->samples[i] = self->samples[(size_t)i * self->channels + c];
    return (PyObject *)r;
}

static PyObject *FrameList_richcompare(PyObject *a, PyObject *b, int op)
{
    if ((op != Py_EQ && op != Py_NE) || Py_TYPE(a) != &pcm_FrameListType || Py_TYPE(b) != &pcm_FrameListType)
        Py_RETURN_NOTIMPLEMENTED;
    pcm_FrameList *x = (pcm_FrameList *)a, *y = (pcm_FrameList *)b;
    int eq = x->frames == y->frames && x->channels == y->channels && x->bits_per_sample == y->bits_per_sample &&
             memcmp(x->samples, y->samples, sizeof(int) * x->samples_length) == 0;
    if ((op == Py_EQ) == (eq != 0)) Py_RETURN_TRUE;
    Py_RETURN_FALSE;
}

static PyObject *get_frames(pcm_FrameList *s, void *c) { return PyLong_FromUnsignedLong(s->frames); }
static PyObject *get_channels(pcm_FrameList *s, void *c) { return PyLong_FromUnsignedLong(s->channels); }
static PyObject *get_bps(pcm_FrameList *s, void *c) { return PyLong_FromUnsignedLong(s->bits_per_sample); }

static PyGetSetDef FrameList_getset[] = {
    {"frames", (getter)get_frames, NULL, "frame count", NULL},
    {"channels", (getter)get_channels, NULL, "channel count", NULL},
    {"bits_per_sample", (getter)get_bps, NULL, "bits per sample", NULL},
    {NULL}};

static PyMethodDef FrameList_methods[] = {
    {"split", (PyCFunction)FrameList_split, METH_VARARGS, "split(pcm_frames) -> (head, tail)"},
    {"to_bytes", (PyCFunction)FrameList_to_bytes, METH_VARARGS, "to_bytes(is_big_endian, is_signed) -> bytes"},
    {"frame", (PyCFunction)FrameList_frame, METH_VARARGS, "frame(i) -> FrameList"},
    {"channel", (PyCFunction)FrameList_channel, METH_VARARGS, "channel(i) -> FrameList"},
    {NULL}};

static PySequenceMethods FrameList_as_sequence = {
    (lenfunc)FrameList_len, (binaryfunc)FrameList_concat, 0, (ssizeargfunc)FrameList_item};

static PyNumberMethods FrameList_as_number = {.nb_add = (binaryfunc)FrameList_concat};

static PyTypeObject pcm_FrameListType = {
    PyVarObject_HEAD_INIT(NULL, 0).tp_name = "pcm.FrameList",
    .tp_basicsize = sizeof(pcm_FrameList),
    .tp_dealloc = (destructor)FrameList_dealloc,
    .tp_as_number = &FrameList_as_number,
    .tp_as_sequence = &FrameList_as_sequence,
    .tp_flags = Py_TPFLAGS_DEFAULT,
    .tp_doc = "FrameList(bytes, channels, bits_per_sample, is_big_endian, is_signed)",
    .tp_richcompare = FrameList_richcompare,
    .tp_methods = FrameList_methods,
    .tp_getset = FrameList_getset,
    .tp_init = (initproc)FrameList_init,
    .tp_new = FrameList_new,
};

/* pcm.from_list(list, channels, bits_per_sample, is_signed), src/pcm.c FrameList_from_list */
static PyObject *pcm_from_list(PyObject *mod, PyObject *args)
{
    PyObject *list;
    int channels, bps, is_signed;
    if (!PyArg_ParseTuple(args, "Oiii", &list, &channels, &bps, &is_signed)) return NULL;
    PyObject *seq = PySequence_Fast(list, "from_list requires a sequence of integers");
    if (!seq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    if (channels < 1 || n % channels) {
        Py_DECREF(seq);
        PyErr_SetString(PyExc_ValueError, "number of samples must be divisible by number of channels");
        return NULL;
    }
    if (bps != 8 && bps != 16 && bps != 24) { Py_DECREF(seq); PyErr_SetString(PyExc_ValueError, "unsupported number of bits per sample"); return NULL; }
    pcm_FrameList *f = framelist_alloc((unsigned)(n / channels), (unsigned)channels, (unsigned)bps);
    if (!f) { Py_DECREF(seq); return NULL; }
    const int adjust = is_signed ? 0 : (1 << (bps - 1));
    for (Py_ssize_t i = 0; i < n; i++) {
        long v = PyLong_AsLong(PySequence_Fast_GET_ITEM(seq, i));
        if (v == -1 && PyErr_Occurred()) { Py_DECREF(seq); Py_DECREF(f); return NULL; }
        f->samples[i] = (int)v - adjust;
    }
    Py_DECREF(seq);
    return (PyObject *)f;
}

static PyMethodDef module_methods[] = {
    {"from_list", (PyCFunction)pcm_from_list, METH_VARARGS, "from_list(int_list, channels, bits_per_sample, is_signed) -> FrameList"},
    {NULL}};

static struct PyModuleDef pcm_module = {PyModuleDef_HEAD_INIT, "audiotools.pcm",
                                        "PCM FrameList boundary type (B200 FLAC engine)", -1, module_methods};

PyMODINIT_FUNC PyInit_pcm(void)
{
    if (PyType_Ready(&pcm_FrameListType) < 0) return NULL;
    PyObject *m = PyModule_Create(&pcm_module);
    if (!m) return NULL;
    Py_INCREF(&pcm_FrameListType);
    PyModule_AddObject(m, "FrameList", (PyObject *)&pcm_FrameListType);
    return m;
}
