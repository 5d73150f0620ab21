"""ctypes binding of the TTA entry points of libb200flac.so (include/b200tta.h)."""
import ctypes as C
import os

import b200flac


class TtaParams(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32)]


class B200TtaError(RuntimeError):
    pass


_ready = False


def lib():
    global _ready
    L = b200flac.lib()
    if not _ready:
        vp, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.b200tta_last_error.restype = C.c_char_p
        L.b200tta_free.argtypes = [vp]
        L.b200tta_block_size.restype = C.c_uint32
        L.b200tta_block_size.argtypes = [C.c_uint32]
        L.b200tta_output_bound.restype = C.c_uint64
        L.b200tta_output_bound.argtypes = [C.POINTER(TtaParams), C.c_uint64, C.c_uint32]
        L.b200tta_encode_frames.argtypes = [C.POINTER(TtaParams), vp, C.c_uint64, u32p, C.c_uint32, C.c_int,
                                            C.POINTER(vp), u64p, C.POINTER(vp), u32p, C.POINTER(C.c_float)]
        L.b200tta_encode_device.argtypes = [C.POINTER(TtaParams), vp, C.c_uint64, C.c_int, vp, C.c_uint64, u64p, vp,
                                            u32p, C.POINTER(C.c_float)]
        L.b200tta_encode_file.argtypes = [C.c_char_p, C.POINTER(TtaParams), vp, C.c_uint64, C.c_int]
        _ready = True
    return L


def _err():
    return B200TtaError(lib().b200tta_last_error().decode("utf-8", "replace"))


def block_size(sample_rate):
    return lib().b200tta_block_size(sample_rate)


def encode_frames(pcm, n_pcm_frames, sample_rate, channels, bits_per_sample, frame_lengths=None, device=0):
    """packed PCM bytes -> (frame bytes, [frame sizes], [kernel ms]): b200tta_encode_frames"""
    L = lib()
    p = TtaParams(sample_rate, channels, bits_per_sample)
    out, sizes = C.c_void_p(), C.c_void_p()
    nbytes, nfr = C.c_uint64(0), C.c_uint32(0)
    ms = (C.c_float * 3)()
    lens, nl = None, 0
    if frame_lengths is not None:
        lens = (C.c_uint32 * len(frame_lengths))(*frame_lengths)
        nl = len(frame_lengths)
    if L.b200tta_encode_frames(C.byref(p), b200flac._buf_ptr(pcm), n_pcm_frames, lens, nl, device, C.byref(out),
                               C.byref(nbytes), C.byref(sizes), C.byref(nfr), ms):
        raise _err()
    data = C.string_at(out, nbytes.value)
    sz = list((C.c_uint32 * nfr.value).from_address(sizes.value)) if nfr.value else []
    L.b200tta_free(out)
    L.b200tta_free(sizes)
    return data, sz, list(ms)


def encode_file(filename, pcm, n_pcm_frames, sample_rate, channels, bits_per_sample, device=0):
    """the standalone reference's `ttaenc`: header, seektable, frames (b200tta_encode_file)"""
    p = TtaParams(sample_rate, channels, bits_per_sample)
    if lib().b200tta_encode_file(os.fsencode(filename), C.byref(p), b200flac._buf_ptr(pcm), n_pcm_frames, device):
        raise _err()
