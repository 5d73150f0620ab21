"""ctypes binding of the ALAC entry points of libb200flac.so (include/b200alac.h)."""
import ctypes as C
import os

import b200flac


class AlacParams(C.Structure):
    _fields_ = [("channels", C.c_uint32), ("bits_per_sample", C.c_uint32), ("block_size", C.c_uint32),
                ("initial_history", C.c_uint32), ("history_multiplier", C.c_uint32), ("maximum_k", C.c_uint32),
                ("minimum_interlacing_leftweight", C.c_uint32), ("maximum_interlacing_leftweight", C.c_uint32)]


class B200AlacError(RuntimeError):
    pass


_ready = False


def lib():
    global _ready
    L = b200flac.lib()
    if not _ready:
        vp, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.b200alac_last_error.restype = C.c_char_p
        L.b200alac_free.argtypes = [vp]
        L.b200alac_output_bound.restype = C.c_uint64
        L.b200alac_output_bound.argtypes = [C.POINTER(AlacParams), C.c_uint64, C.c_uint32]
        L.b200alac_encode_framesets.argtypes = [C.POINTER(AlacParams), vp, C.c_uint64, u32p, C.c_uint32, C.c_int,
                                                C.POINTER(vp), u64p, C.POINTER(vp), u32p, C.POINTER(C.c_float)]
        L.b200alac_encode_device.argtypes = [C.POINTER(AlacParams), vp, C.c_uint64, C.c_int, vp, C.c_uint64, u64p, vp,
                                             u32p, C.POINTER(C.c_float)]
        L.b200alac_encode_mdat.argtypes = [C.c_char_p, C.POINTER(AlacParams), vp, C.c_uint64, C.c_int]
        _ready = True
    return L


def _err():
    return B200AlacError(lib().b200alac_last_error().decode("utf-8", "replace"))


def make_params(channels=2, bits_per_sample=16, block_size=4096, initial_history=10, history_multiplier=40, maximum_k=14,
                minimum_interlacing_leftweight=0, maximum_interlacing_leftweight=4):
    return AlacParams(channels, bits_per_sample, block_size, initial_history, history_multiplier, maximum_k,
                      minimum_interlacing_leftweight, maximum_interlacing_leftweight)


def encode_framesets(pcm, n_pcm_frames, params, frame_lengths=None, device=0):
    """packed PCM bytes -> (frameset bytes, [frameset sizes], [kernel ms]): b200alac_encode_framesets"""
    L = lib()
    out, sizes = C.c_void_p(), C.c_void_p()
    nbytes, nfr = C.c_uint64(0), C.c_uint32(0)
    ms = (C.c_float * 4)()
    lens, nl = None, 0
    if frame_lengths is not None:
        lens = (C.c_uint32 * len(frame_lengths))(*frame_lengths)
        nl = len(frame_lengths)
    if L.b200alac_encode_framesets(C.byref(params), b200flac._buf_ptr(pcm), n_pcm_frames, lens, nl, device, C.byref(out),
                                   C.byref(nbytes), C.byref(sizes), C.byref(nfr), ms):
        raise _err()
    data = C.string_at(out, nbytes.value)
    sz = list((C.c_uint32 * nfr.value).from_address(sizes.value)) if nfr.value else []
    L.b200alac_free(out)
    L.b200alac_free(sizes)
    return data, sz, list(ms)


def encode_mdat(filename, pcm, n_pcm_frames, params, device=0):
    """the standalone reference's `alacenc`: the mdat atom (b200alac_encode_mdat)"""
    if lib().b200alac_encode_mdat(os.fsencode(filename), C.byref(params), b200flac._buf_ptr(pcm), n_pcm_frames, device):
        raise _err()
