"""ctypes binding of the B200 FLAC engine's C ABI (include/b200flac.h).

This is the thinnest possible host-language stub over ``libb200flac.so``; it is
what bench.py, the GPU parity tests and ``__graft_entry__.smoke()`` call.  The
CPython extension ``audiotools.encoders`` binds the same library from C.

There is no CPU fallback: if the shared library is missing, or no CUDA device
is usable, every entry point raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200FLAC_LIB") or os.path.join(_HERE, "libb200flac.so")   # (the override is for A/B builds of tools/)

MAX_LPC_ORDER = 32


class Params(C.Structure):
    """struct b200flac_params (mirrors flac_encoding_options, src/encoders/flac.h:30-47)"""
    _fields_ = [("block_size", C.c_uint32),
                ("max_lpc_order", C.c_uint32),
                ("min_residual_partition_order", C.c_uint32),
                ("max_residual_partition_order", C.c_uint32),
                ("mid_side", C.c_int32),
                ("adaptive_mid_side", C.c_int32),
                ("exhaustive_model_search", C.c_int32),
                ("no_verbatim_subframes", C.c_int32),
                ("no_constant_subframes", C.c_int32),
                ("no_fixed_subframes", C.c_int32),
                ("no_lpc_subframes", C.c_int32),
                ("sample_rate", C.c_uint32),
                ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32)]


class Segment(C.Structure):
    _fields_ = [("pcm_frame_offset", C.c_uint64),
                ("n_pcm_frames", C.c_uint64),
                ("first_frame_number", C.c_uint32),
                ("reserved", C.c_uint32)]


class Plan(C.Structure):
    _fields_ = [("type", C.c_uint8), ("order", C.c_uint8), ("wasted", C.c_uint8),
                ("precision", C.c_uint8), ("shift", C.c_int8), ("coding_method", C.c_uint8),
                ("partition_order", C.c_uint8), ("flags", C.c_uint8), ("bits", C.c_uint32),
                ("coeffs", C.c_int16 * MAX_LPC_ORDER)]


class PcmSource(C.Structure):
    """struct b200flac_pcm_source: what WaveReader / AiffReader learn from the container
    (audiotools/wav.py:424-502, aiff.py:353-432)"""
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("channel_mask", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("data_offset", C.c_uint64), ("total_pcm_frames", C.c_uint64)]


PCM_BIG_ENDIAN, PCM_UNSIGNED = 1, 2


class StreamInfo(C.Structure):
    """struct b200flac_stream_info: STREAMINFO as flacdec_read_metadata reads it (src/decoders/flac.c:569-708)"""
    _fields_ = [("min_block_size", C.c_uint32), ("max_block_size", C.c_uint32), ("min_frame_size", C.c_uint32),
                ("max_frame_size", C.c_uint32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("channel_mask", C.c_uint32), ("total_pcm_frames", C.c_uint64),
                ("first_frame_offset", C.c_uint64), ("md5", C.c_uint8 * 16)]

_lib = None


def lib():
    """loads libb200flac.so (built by python-audio-tools_b200/Makefile or __graft_entry__.build)"""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the engine has no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u8p, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    L.b200flac_abi_version.restype = C.c_int
    L.b200flac_device_count.restype = C.c_int
    L.b200flac_last_error.restype = C.c_char_p
    L.b200flac_encoder_create.restype = vp
    L.b200flac_encoder_create.argtypes = [C.POINTER(Params), C.c_int, C.c_uint64, C.c_int]
    L.b200flac_encoder_destroy.argtypes = [vp]
    L.b200flac_encoder_output_bound.restype = C.c_uint64
    L.b200flac_encoder_output_bound.argtypes = [vp, C.c_uint64, C.c_uint32]
    L.b200flac_encoder_slot_pcm.restype = vp
    L.b200flac_encoder_slot_pcm.argtypes = [vp, C.c_int]
    L.b200flac_encoder_slot_out.restype = vp
    L.b200flac_encoder_slot_out.argtypes = [vp, C.c_int, u64p]
    L.b200flac_encoder_submit.argtypes = [vp, C.c_int, vp, C.POINTER(Segment), C.c_uint32]
    L.b200flac_encoder_collect.argtypes = [vp, C.c_int, vp, C.c_uint64, u64p, vp, vp, C.c_uint32, u32p]
    L.b200flac_encoder_encode.argtypes = [vp, vp, C.POINTER(Segment), C.c_uint32, vp, C.c_uint64, u64p,
                                          vp, vp, C.c_uint32, u32p]
    L.b200flac_encoder_encode_device.argtypes = [vp, C.c_int, vp, C.POINTER(Segment), C.c_uint32, vp,
                                                 C.c_uint64, u64p, u32p, C.POINTER(C.c_float)]
    L.b200flac_encoder_submit_device.argtypes = [vp, C.c_int, vp, C.POINTER(Segment), C.c_uint32, vp, C.c_uint64]
    L.b200flac_encoder_collect_device.argtypes = [vp, C.c_int, u64p, u32p, C.POINTER(C.c_float)]
    L.b200flac_encoder_last_kernel_ms.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.c_int]
    L.b200flac_encoder_set_chunking.restype = C.c_int
    L.b200flac_encoder_set_chunking.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.b200flac_encoder_launch_count.restype = C.c_uint64
    L.b200flac_encoder_launch_count.argtypes = [vp]
    L.b200flac_launch_count_total.restype = C.c_uint64
    L.b200flac_host_alloc.restype = vp
    L.b200flac_host_alloc.argtypes = [C.c_uint64]
    L.b200flac_host_free.argtypes = [vp]
    L.b200flac_device_alloc.restype = vp
    L.b200flac_device_alloc.argtypes = [C.c_int, C.c_uint64]
    L.b200flac_device_free.argtypes = [C.c_int, vp]
    L.b200flac_device_upload.argtypes = [C.c_int, vp, vp, C.c_uint64]
    L.b200flac_device_download.argtypes = [C.c_int, vp, vp, C.c_uint64]
    L.b200flac_device_synth_pcm.argtypes = [C.c_int, vp, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64,
                                            C.c_uint64]
    L.b200flac_encoder_get_plans.argtypes = [vp, C.c_int, vp, vp, u32p, vp, u32p]
    L.b200flac_stream_open.restype = vp
    L.b200flac_stream_open.argtypes = [C.c_char_p, C.POINTER(Params), C.c_uint32, C.c_char_p,
                                       C.POINTER(C.c_int), C.c_int]
    L.b200flac_stream_write.argtypes = [vp, vp, C.c_uint64]
    L.b200flac_stream_end_block.argtypes = [vp]
    L.b200flac_stream_close.argtypes = [vp, C.c_int, C.POINTER(u64p), C.POINTER(u32p), u64p]
    L.b200flac_free.argtypes = [vp]
    L.b200flac_finalize_metadata.argtypes = [C.c_char_p, u64p, u32p, C.c_uint64, C.c_uint32, C.c_uint32]
    L.b200flac_encode_file.argtypes = [C.c_char_p, C.POINTER(Params), C.c_uint32, C.c_char_p, vp,
                                       C.c_uint64, C.POINTER(C.c_int), C.c_int]
    if hasattr(L, "b200flac_encode_files"):        # (absent from older A/B builds loaded through B200FLAC_LIB)
        L.b200flac_encode_files.argtypes = [C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(Params), C.c_uint32, C.c_char_p,
                                            C.POINTER(vp), u64p, C.c_int, C.c_int]
        L.b200flac_internal_device_md5.argtypes = [C.c_int, vp, C.c_uint64, vp]
    L.b200flac_wave_probe.argtypes = [C.c_char_p, C.POINTER(PcmSource)]
    L.b200flac_aiff_probe.argtypes = [C.c_char_p, C.POINTER(PcmSource)]
    L.b200flac_stream_write_file.argtypes = [vp, C.c_char_p, C.c_uint64, C.c_uint64, C.c_uint32]
    for fn in (L.b200flac_encode_wave, L.b200flac_encode_aiff):
        fn.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(Params), C.c_uint32, C.c_char_p, C.POINTER(C.c_int),
                       C.c_int, C.POINTER(PcmSource), C.POINTER(u64p), C.POINTER(u32p), u64p]
    L.b200flac_read_streaminfo.argtypes = [vp, C.c_uint64, C.POINTER(StreamInfo)]
    L.b200flac_decode_memory.argtypes = [vp, C.c_uint64, C.c_int, vp, C.c_uint64, C.POINTER(StreamInfo), C.c_int,
                                         C.POINTER(u64p), C.POINTER(u32p), u64p, C.POINTER(C.c_float)]
    L.b200flac_decode_device.argtypes = [C.POINTER(StreamInfo), vp, C.c_uint64, C.c_int, vp, C.c_uint64, u64p,
                                         C.POINTER(C.c_float)]
    L.b200flac_verify_file.argtypes = [C.c_char_p, C.c_int]
    L.b200flac_decode_to_wave.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    _lib = L
    return L


class B200FlacError(RuntimeError):
    pass


def _err():
    return B200FlacError(lib().b200flac_last_error().decode("utf-8", "replace"))


def make_params(sample_rate=44100, channels=2, bits_per_sample=16, block_size=4096, max_lpc_order=8,
                min_residual_partition_order=0, max_residual_partition_order=5, mid_side=False,
                adaptive_mid_side=False, exhaustive_model_search=False,
                disable_verbatim_subframes=False, disable_constant_subframes=False,
                disable_fixed_subframes=False, disable_lpc_subframes=False):
    return Params(block_size, max_lpc_order, min_residual_partition_order, max_residual_partition_order,
                  int(bool(mid_side)), int(bool(adaptive_mid_side)), int(bool(exhaustive_model_search)),
                  int(bool(disable_verbatim_subframes)), int(bool(disable_constant_subframes)),
                  int(bool(disable_fixed_subframes)), int(bool(disable_lpc_subframes)),
                  sample_rate, channels, bits_per_sample)


def _buf_ptr(b):
    """address of a bytes/bytearray/numpy buffer without copying"""
    if isinstance(b, int):
        return b
    if hasattr(b, "ctypes"):
        return b.ctypes.data
    if isinstance(b, bytes):
        return C.cast(C.c_char_p(b), C.c_void_p).value
    return C.addressof((C.c_char * len(b)).from_buffer(b))


class Encoder(object):
    """frame layer: batches of PCM blocks -> FLAC frames on one CUDA device"""

    def __init__(self, params, device=0, max_pcm_frames_per_batch=1 << 22, n_slots=2):
        self.params = params
        self.device = device
        self.h = lib().b200flac_encoder_create(C.byref(params), device, max_pcm_frames_per_batch, n_slots)
        if not self.h:
            raise _err()
        self.max_pcm_frames = max_pcm_frames_per_batch
        self.n_slots = n_slots

    def close(self):
        if self.h:
            lib().b200flac_encoder_destroy(self.h)
            self.h = None

    __del__ = close

    def output_bound(self, n_pcm_frames, n_segments=1):
        return lib().b200flac_encoder_output_bound(self.h, n_pcm_frames, n_segments)

    @staticmethod
    def _segments(segments):
        arr = (Segment * len(segments))()
        for i, (off, n, first) in enumerate(segments):
            arr[i] = Segment(off, n, first, 0)
        return arr

    def submit(self, slot, pcm, segments):
        arr = self._segments(segments)
        if lib().b200flac_encoder_submit(self.h, slot, _buf_ptr(pcm), arr, len(segments)):
            raise _err()

    def collect(self, slot, n_pcm_frames, n_segments=1):
        import numpy as np
        cap = self.output_bound(n_pcm_frames, n_segments)
        fcap = n_pcm_frames // self.params.block_size + n_segments + 2
        out = np.empty(cap, dtype=np.uint8)
        fbytes = np.empty(fcap, dtype=np.uint32)
        fpcm = np.empty(fcap, dtype=np.uint32)
        nbytes, nfr = C.c_uint64(0), C.c_uint32(0)
        if lib().b200flac_encoder_collect(self.h, slot, out.ctypes.data, cap, C.byref(nbytes),
                                          fbytes.ctypes.data, fpcm.ctypes.data, fcap, C.byref(nfr)):
            raise _err()
        return out[:nbytes.value], fbytes[:nfr.value], fpcm[:nfr.value]

    def encode(self, pcm, n_pcm_frames, first_frame_number=0, segments=None):
        """host PCM bytes -> (frame bytes, per-frame sizes, per-frame PCM counts)"""
        if segments is None:
            segments = [(0, n_pcm_frames, first_frame_number)]
        self.submit(0, pcm, segments)
        return self.collect(0, n_pcm_frames, len(segments))

    def encode_device(self, d_pcm, segments, d_out, out_capacity, slot=0):
        """device pointers in, frames left on the device; returns (bytes, frames, kernel ms)"""
        arr = self._segments(segments)
        nbytes, nfr, ms = C.c_uint64(0), C.c_uint32(0), C.c_float(0)
        if lib().b200flac_encoder_encode_device(self.h, slot, d_pcm, arr, len(segments), d_out, out_capacity,
                                                C.byref(nbytes), C.byref(nfr), C.byref(ms)):
            raise _err()
        return nbytes.value, nfr.value, ms.value

    def submit_device(self, d_pcm, segments, d_out, out_capacity, slot=0):
        """asynchronous half of encode_device: host work and launches, no wait"""
        arr = self._segments(segments)
        if lib().b200flac_encoder_submit_device(self.h, slot, d_pcm, arr, len(segments), d_out, out_capacity):
            raise _err()

    def collect_device(self, slot=0):
        """waits for the slot: (bytes, frames, kernel ms)"""
        nbytes, nfr, ms = C.c_uint64(0), C.c_uint32(0), C.c_float(0)
        if lib().b200flac_encoder_collect_device(self.h, slot, C.byref(nbytes), C.byref(nfr), C.byref(ms)):
            raise _err()
        return nbytes.value, nfr.value, ms.value

    def kernel_ms(self, slot=0):
        ms = (C.c_float * 8)()
        n = lib().b200flac_encoder_last_kernel_ms(self.h, slot, ms, 8)
        return [ms[i] for i in range(n)]

    def set_chunking(self, chunk_frames, lookahead=3):
        """frames per pipeline chunk (0: one chunk per batch, kernels back to back: per-kernel times exist)"""
        if lib().b200flac_encoder_set_chunking(self.h, chunk_frames, lookahead):
            raise _err()

    def launch_count(self):
        return lib().b200flac_encoder_launch_count(self.h)

    def plans(self, slot, n_frames):
        import numpy as np
        stride, k = C.c_uint32(0), C.c_uint32(0)
        if lib().b200flac_encoder_get_plans(self.h, slot, None, None, C.byref(stride), None, C.byref(k)):
            raise _err()
        units = n_frames * k.value
        plans = (Plan * units)()
        rice = np.zeros(units * stride.value, dtype=np.uint8)
        asg = np.zeros(n_frames, dtype=np.uint8)
        if lib().b200flac_encoder_get_plans(self.h, slot, plans, rice.ctypes.data, None, asg.ctypes.data, None):
            raise _err()
        return plans, rice.reshape(units, stride.value), asg, k.value


def device_count():
    return lib().b200flac_device_count()


def synth_pcm(seed, channels, bits_per_sample, n_pcm_frames, device=0, first_frame=0):
    """the bench's integer synthetic signal (k_synth.cuh), generated on the device and downloaded: packed PCM bytes"""
    import numpy as np
    L = lib()
    nbytes = n_pcm_frames * channels * (bits_per_sample // 8)
    d = L.b200flac_device_alloc(device, max(nbytes, 1))
    if not d:
        raise _err()
    try:
        if L.b200flac_device_synth_pcm(device, d, seed, channels, bits_per_sample, first_frame, n_pcm_frames):
            raise _err()
        out = np.empty(max(nbytes, 1), dtype=np.uint8)
        if L.b200flac_device_download(device, out.ctypes.data, d, nbytes):
            raise _err()
    finally:
        L.b200flac_device_free(device, d)
    return out[:nbytes].tobytes()


def encode_file(filename, params, pcm, n_pcm_frames, padding_size=4096, version=None, devices=None):
    """standalone-reference equivalent: packed PCM in memory -> FLAC file (flac.c:124-306)"""
    devs = None
    ndev = 0
    if devices:
        devs = (C.c_int * len(devices))(*devices)
        ndev = len(devices)
    v = version.encode() if version else None
    if lib().b200flac_encode_file(os.fsencode(filename), C.byref(params), padding_size, v, _buf_ptr(pcm),
                                  n_pcm_frames, devs, ndev):
        raise _err()


def encode_files(filenames, params, pcms, n_pcm_frames, padding_size=4096, version=None, device=-1, host_threads=0):
    """many tracks -> many files in one call (b200flac_encode_files): pcms[i] is a buffer or an address holding
    n_pcm_frames[i] frames; every file equals what encode_file writes for that track"""
    n = len(filenames)
    names = (C.c_char_p * n)(*[os.fsencode(f) for f in filenames])
    ptrs = (C.c_void_p * n)(*[_buf_ptr(b) for b in pcms])
    lens = (C.c_uint64 * n)(*[int(v) for v in n_pcm_frames])
    v = version.encode() if version else None
    if lib().b200flac_encode_files(n, names, C.byref(params), padding_size, v, ptrs, lens, device, host_threads):
        raise _err()


def host_md5_many(datas):
    """test hook: MD5 of up to sixteen byte strings hashed side by side on one host core (csrc/md5_lanes.cpp)"""
    n = len(datas)
    keep = [bytes(d) if len(d) else b"\0" for d in datas]
    ptr = (C.c_char_p * n)(*keep)
    lens = (C.c_uint64 * n)(*[len(d) for d in datas])
    out = (C.c_uint8 * (16 * n))()
    fn = lib().b200flac_internal_md5_many
    fn.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
    if fn(ptr, lens, n, out, 1 << 16, None, None):
        raise _err()
    return [bytes(out[16 * i:16 * i + 16]) for i in range(n)]


def device_md5(data, device=0):
    """test hook: MD5 of a byte string computed by the batch entry's device kernel"""
    out = (C.c_uint8 * 16)()
    if lib().b200flac_internal_device_md5(device, _buf_ptr(data), len(data), out):
        raise _err()
    return bytes(out)


class Stream(object):
    """stream layer: open / write PCM / close, returns the (offset, frames) list of flac.c:249-253"""

    def __init__(self, filename, params, padding_size=4096, version=None, devices=None):
        devs, ndev = None, 0
        if devices:
            devs = (C.c_int * len(devices))(*devices)
            ndev = len(devices)
        v = version.encode() if version else None
        self.h = lib().b200flac_stream_open(os.fsencode(filename), C.byref(params), padding_size, v, devs, ndev)
        if not self.h:
            raise _err()
        self.frame_bytes = params.channels * (params.bits_per_sample // 8)

    def write(self, pcm):
        n = len(pcm) // self.frame_bytes
        if lib().b200flac_stream_write(self.h, _buf_ptr(pcm), n):
            raise _err()

    def end_block(self):
        if lib().b200flac_stream_end_block(self.h):
            raise _err()

    def write_file(self, path, byte_offset, n_pcm_frames, flags=0):
        """PCM frames straight from a file into the pinned staging (b200flac_stream_write_file)"""
        rc = lib().b200flac_stream_write_file(self.h, os.fsencode(path), byte_offset, n_pcm_frames, flags)
        if rc == 2:
            raise IOError(lib().b200flac_last_error().decode())
        if rc:
            raise _err()

    def close(self, abort=False):
        if not self.h:
            return []
        offs, lens, n = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.c_uint64(0)
        h, self.h = self.h, None
        if lib().b200flac_stream_close(h, int(abort), C.byref(offs), C.byref(lens), C.byref(n)):
            raise _err()
        if abort:
            return []
        res = [(offs[i], lens[i]) for i in range(n.value)]
        lib().b200flac_free(offs)
        lib().b200flac_free(lens)
        return res


def finalize_metadata(filename, offsets, seekpoint_interval=0, channel_mask=0):
    """SEEKTABLE (+ channel-mask tag) into the finished file, natively: what FlacAudio.from_pcm does after
    encode_flac returns (audiotools/flac.py:1811-1832).  offsets: the encoder's [(byte offset, PCM frames)]."""
    n = len(offsets)
    offs = (C.c_uint64 * max(n, 1))(*[o for o, _ in offsets])
    lens = (C.c_uint32 * max(n, 1))(*[f for _, f in offsets])
    if lib().b200flac_finalize_metadata(os.fsencode(filename), offs, lens, n, seekpoint_interval, channel_mask):
        raise _err()


def _probe(fn, path):
    src = PcmSource()
    rc = fn(os.fsencode(path), C.byref(src))
    if rc == 1:
        raise ValueError(lib().b200flac_last_error().decode())
    if rc:
        raise IOError(lib().b200flac_last_error().decode())
    return src


def wave_probe(path):
    """WaveReader.__init__ in C (audiotools/wav.py:424-502): ValueError / IOError as the reference raises"""
    return _probe(lib().b200flac_wave_probe, path)


def aiff_probe(path):
    """AiffReader.__init__ in C (audiotools/aiff.py:353-432)"""
    return _probe(lib().b200flac_aiff_probe, path)


def encode_container(flac_filename, in_filename, params, kind="wave", padding_size=4096, version=None, devices=None):
    """file to file: b200flac_encode_wave / b200flac_encode_aiff; returns (PcmSource, [(offset, pcm_frames)])"""
    devs, ndev = None, 0
    if devices:
        devs = (C.c_int * len(devices))(*devices)
        ndev = len(devices)
    v = version.encode() if version else None
    fn = lib().b200flac_encode_wave if kind == "wave" else lib().b200flac_encode_aiff
    src = PcmSource()
    offs, lens, n = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.c_uint64(0)
    rc = fn(os.fsencode(flac_filename), os.fsencode(in_filename), C.byref(params), padding_size, v, devs, ndev,
            C.byref(src), C.byref(offs), C.byref(lens), C.byref(n))
    if rc == 1:
        raise ValueError(lib().b200flac_last_error().decode())
    if rc == 2:
        raise IOError(lib().b200flac_last_error().decode())
    if rc:
        raise _err()
    res = [(offs[i], lens[i]) for i in range(n.value)]
    lib().b200flac_free(offs)
    lib().b200flac_free(lens)
    return src, res


def _raise_decode(rc):
    msg = lib().b200flac_last_error().decode("utf-8", "replace")
    if rc == 1:
        raise ValueError(msg)        # what FlacDecoder.read() raises (src/decoders/flac.c:218-253)
    if rc == 2:
        raise IOError(msg)           # "EOF reading frame", flac.c:262
    raise B200FlacError(msg)


def read_streaminfo(flac):
    info = StreamInfo()
    rc = lib().b200flac_read_streaminfo(_buf_ptr(flac), len(flac), C.byref(info))
    if rc:
        _raise_decode(rc)
    return info


def decode(flac, device=0, check_md5=True, want_frames=False):
    """a FLAC file image (bytes) -> (StreamInfo, PCM bytes[, [(frame offset, pcm frames)], kernel ms]):
    b200flac_decode_memory; raises ValueError / IOError with the reference decoder's messages"""
    import numpy as np
    info = StreamInfo()
    # sizing call: STREAMINFO's total is checked against the file size before a buffer is sized from it
    rc = lib().b200flac_decode_memory(_buf_ptr(flac), len(flac), device, None, 0, C.byref(info), 0, None, None, None, None)
    if rc:
        _raise_decode(rc)
    nbytes = info.total_pcm_frames * info.channels * (info.bits_per_sample // 8)
    pcm = np.empty(max(nbytes, 1), dtype=np.uint8)
    offs, lens, n = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint32)(), C.c_uint64(0)
    ms = (C.c_float * 3)()
    rc = lib().b200flac_decode_memory(_buf_ptr(flac), len(flac), device, pcm.ctypes.data, nbytes, None, int(check_md5),
                                      C.byref(offs), C.byref(lens), C.byref(n), ms)
    if rc:
        _raise_decode(rc)
    out = pcm[:nbytes].tobytes()
    if not want_frames:
        return info, out
    frames = [(offs[i], lens[i]) for i in range(n.value)]
    if n.value:
        lib().b200flac_free(offs)
        lib().b200flac_free(lens)
    return info, out, frames, list(ms)


def verify_file(path, device=0):
    """flacdec without output (b200flac_verify_file): every frame CRC-16 and the STREAMINFO MD5"""
    rc = lib().b200flac_verify_file(os.fsencode(path), device)
    if rc:
        _raise_decode(rc)
    return True


def decode_to_wave(flac_path, wave_path, device=0):
    """FLAC file -> RIFF WAVE file (b200flac_decode_to_wave)"""
    rc = lib().b200flac_decode_to_wave(os.fsencode(flac_path), os.fsencode(wave_path), device)
    if rc:
        _raise_decode(rc)
