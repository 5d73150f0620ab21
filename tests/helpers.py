"""Test-side helpers: ctypes binding of the ORACLE (oracle/liboracle_flac.so, the CPU
restatement of the reference) and of the compiled reference itself (oracle/_ref), PCM
generators restating the reference's test streams, and a FLAC decision parser.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may use oracle/."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle_flac.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_FLACENC = os.path.join(REF_DIR, "flacenc")
REF_FLACDEC = os.path.join(REF_DIR, "flacdec")
GOLDEN = os.path.join(ROOT, "tests", "golden")


class OrcOptions(C.Structure):
    _fields_ = [("block_size", C.c_uint), ("max_lpc_order", C.c_uint),
                ("min_residual_partition_order", C.c_uint), ("max_residual_partition_order", C.c_uint),
                ("mid_side", C.c_int), ("adaptive_mid_side", C.c_int), ("exhaustive_model_search", C.c_int),
                ("no_verbatim_subframes", C.c_int), ("no_constant_subframes", C.c_int),
                ("no_fixed_subframes", C.c_int), ("no_lpc_subframes", C.c_int),
                ("padding_size", C.c_uint), ("version", C.c_char_p)]


_orc = None


def orc():
    global _orc
    if _orc is None:
        L = C.CDLL(ORACLE_LIB)
        L.orc_encode_stream.argtypes = [C.POINTER(OrcOptions), C.c_uint, C.c_uint, C.c_uint, C.c_void_p,
                                        C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_encode_range.argtypes = [C.POINTER(OrcOptions), C.c_uint, C.c_uint, C.c_uint, C.c_void_p,
                                       C.c_size_t, C.c_uint, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                       C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        L.orc_synth_pcm.argtypes = [C.c_uint64, C.c_uint, C.c_uint, C.c_uint64, C.c_uint64, C.c_void_p]
        L.orc_crc8.restype = C.c_uint8
        L.orc_crc8.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_crc16.restype = C.c_uint16
        L.orc_crc16.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_md5.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_tukey_window.argtypes = [C.c_uint, C.c_void_p]
        L.orc_autocorrelate.argtypes = [C.c_uint, C.c_void_p, C.c_uint, C.c_void_p]
        L.orc_lp_coefficients.argtypes = [C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_estimate_best_lpc_order.restype = C.c_uint
        L.orc_estimate_best_lpc_order.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_void_p]
        L.orc_quantize_coefficients.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.POINTER(C.c_int)]
        L.orc_residual_partitions.restype = C.c_uint64
        L.orc_residual_partitions.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint,
                                              C.c_void_p, C.c_void_p]
        L.orc_best_fixed_order.restype = C.c_uint
        L.orc_best_fixed_order.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
        L.orc_wasted_bits.restype = C.c_uint
        L.orc_wasted_bits.argtypes = [C.c_void_p, C.c_uint]
        _orc = L
    return _orc


def options(block_size=4096, max_lpc_order=8, max_residual_partition_order=5, mid_side=False,
            adaptive_mid_side=False, exhaustive_model_search=False, min_residual_partition_order=0,
            disable_verbatim_subframes=False, disable_constant_subframes=False,
            disable_fixed_subframes=False, disable_lpc_subframes=False, padding_size=4096):
    return dict(block_size=block_size, max_lpc_order=max_lpc_order,
                min_residual_partition_order=min_residual_partition_order,
                max_residual_partition_order=max_residual_partition_order, mid_side=mid_side,
                adaptive_mid_side=adaptive_mid_side, exhaustive_model_search=exhaustive_model_search,
                disable_verbatim_subframes=disable_verbatim_subframes,
                disable_constant_subframes=disable_constant_subframes,
                disable_fixed_subframes=disable_fixed_subframes,
                disable_lpc_subframes=disable_lpc_subframes, padding_size=padding_size)


def pack_pcm(samples, bits_per_sample):
    """interleaved int array -> signed little-endian packed bytes (what the reference's
    standalone encoder reads and its MD5 hashes)"""
    a = np.ascontiguousarray(samples, dtype=np.int32).reshape(-1)
    if bits_per_sample == 8:
        return a.astype(np.int8).tobytes()
    if bits_per_sample == 16:
        return a.astype("<i2").tobytes()
    if bits_per_sample == 24:
        b = a.astype("<i4").view(np.uint8).reshape(-1, 4)[:, :3]
        return np.ascontiguousarray(b).tobytes()
    raise ValueError(bits_per_sample)


def unpack_pcm(data, bits_per_sample):
    if bits_per_sample == 8:
        return np.frombuffer(data, dtype=np.int8).astype(np.int32)
    if bits_per_sample == 16:
        return np.frombuffer(data, dtype="<i2").astype(np.int32)
    b = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    return np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32)


def synth_pcm(seed, channels, bits_per_sample, n_frames, first_frame=0):
    """the deterministic integer synthetic signal (oracle/flac_oracle.c orc_synth_pcm), packed"""
    out = np.empty(n_frames * channels, dtype=np.int32)
    orc().orc_synth_pcm(seed, channels, bits_per_sample, first_frame, n_frames, out.ctypes.data)
    return pack_pcm(out, bits_per_sample)


def oracle_encode(pcm, sample_rate, channels, bits_per_sample, opts, want_offsets=False):
    """whole file image from the CPU oracle"""
    o = OrcOptions(opts["block_size"], opts["max_lpc_order"], opts["min_residual_partition_order"],
                   opts["max_residual_partition_order"], int(opts["mid_side"]), int(opts["adaptive_mid_side"]),
                   int(opts["exhaustive_model_search"]), int(opts["disable_verbatim_subframes"]),
                   int(opts["disable_constant_subframes"]), int(opts["disable_fixed_subframes"]),
                   int(opts["disable_lpc_subframes"]), opts["padding_size"], None)
    out, n = C.c_void_p(), C.c_size_t()
    offs, lens, nfr = C.c_void_p(), C.c_void_p(), C.c_size_t()
    buf = np.frombuffer(pcm, dtype=np.uint8)
    ptr = buf.ctypes.data if len(buf) else None
    orc().orc_encode_stream(C.byref(o), sample_rate, channels, bits_per_sample, ptr, len(buf),
                            C.byref(out), C.byref(n), C.byref(offs), C.byref(lens), C.byref(nfr))
    data = C.string_at(out.value, n.value)
    res_offs = []
    if nfr.value:
        oa = np.ctypeslib.as_array(C.cast(offs.value, C.POINTER(C.c_uint64)), (nfr.value,)).copy()
        la = np.ctypeslib.as_array(C.cast(lens.value, C.POINTER(C.c_uint32)), (nfr.value,)).copy()
        res_offs = list(zip(oa.tolist(), la.tolist()))
    orc().orc_free(out)
    orc().orc_free(offs)
    orc().orc_free(lens)
    return (data, res_offs) if want_offsets else data


def _orc_options(opts):
    return OrcOptions(opts["block_size"], opts["max_lpc_order"], opts["min_residual_partition_order"],
                      opts["max_residual_partition_order"], int(opts["mid_side"]), int(opts["adaptive_mid_side"]),
                      int(opts["exhaustive_model_search"]), int(opts["disable_verbatim_subframes"]),
                      int(opts["disable_constant_subframes"]), int(opts["disable_fixed_subframes"]),
                      int(opts["disable_lpc_subframes"]), opts["padding_size"], None)


def oracle_encode_range(pcm, sample_rate, channels, bits_per_sample, opts, first_frame_number=0):
    """frames only (no stream head) of a PCM range: (frame bytes, [size of each frame])"""
    o = _orc_options(opts)
    out, n, sizes, nfr = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
    buf = np.frombuffer(pcm, dtype=np.uint8)
    ptr = buf.ctypes.data if len(buf) else None
    orc().orc_encode_range(C.byref(o), sample_rate, channels, bits_per_sample, ptr, len(buf),
                           first_frame_number, C.byref(out), C.byref(n), C.byref(sizes), C.byref(nfr))
    data = C.string_at(out.value, n.value) if n.value else b""
    sz = []
    if nfr.value:
        sz = np.ctypeslib.as_array(C.cast(sizes.value, C.POINTER(C.c_uint32)), (nfr.value,)).tolist()
    orc().orc_free(out)
    orc().orc_free(sizes)
    return data, sz


def have_ref():
    return os.path.exists(REF_FLACENC) and os.path.exists(REF_FLACDEC)


def _ref_flags(sample_rate, channels, bits_per_sample, opts):
    f = ["-c", str(channels), "-r", str(sample_rate), "-b", str(bits_per_sample),
         "-B", str(opts["block_size"]), "-l", str(opts["max_lpc_order"]),
         "-P", str(opts["min_residual_partition_order"]), "-R", str(opts["max_residual_partition_order"])]
    if opts["mid_side"]:
        f.append("-m")
    if opts["adaptive_mid_side"]:
        f.append("-M")
    if opts["exhaustive_model_search"]:
        f.append("-e")
    return f


def ref_encode(pcm, sample_rate, channels, bits_per_sample, opts):
    """file image from the COMPILED REFERENCE (oracle/_ref/flacenc, built from /root/reference/src)"""
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.flac")
        subprocess.run([REF_FLACENC] + _ref_flags(sample_rate, channels, bits_per_sample, opts) + [out],
                       input=pcm, stdout=subprocess.DEVNULL, check=True)
        with open(out, "rb") as fh:
            return fh.read()


def ref_decode(flac_bytes):
    """PCM bytes from the reference's standalone decoder; raises if CRC-16 or MD5 fail
    (src/decoders/flac.c:1453-1461,1494-1501 exit non-zero)"""
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "i.flac")
        with open(p, "wb") as fh:
            fh.write(flac_bytes)
        r = subprocess.run([REF_FLACDEC, p], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if r.returncode != 0:
            raise RuntimeError("reference flacdec rejected the stream: %s" % r.stderr[-300:])
        return r.stdout


# ---- TTA: the compiled reference (oracle/_ref/ttaenc, ttadec) and the plain-C oracle ----
REF_TTAENC = os.path.join(ROOT, "oracle", "_ref", "ttaenc")
REF_TTADEC = os.path.join(ROOT, "oracle", "_ref", "ttadec")
_tta = None


def have_tta_ref():
    return os.path.exists(REF_TTAENC) and os.path.exists(REF_TTADEC)


def ref_tta_encode(pcm, sample_rate, channels, bits_per_sample):
    """file image from the COMPILED REFERENCE TTA encoder (src/encoders/tta.c, -DSTANDALONE)"""
    n = len(pcm) // (channels * (bits_per_sample // 8))
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.tta")
        subprocess.run([REF_TTAENC, "-c", str(channels), "-r", str(sample_rate), "-b", str(bits_per_sample),
                        "-T", str(n), out], input=pcm, stdout=subprocess.DEVNULL, check=True)
        with open(out, "rb") as fh:
            return fh.read()


def ref_tta_decode(tta_bytes):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "i.tta")
        with open(p, "wb") as fh:
            fh.write(tta_bytes)
        r = subprocess.run([REF_TTADEC, p], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if r.returncode != 0:
            raise RuntimeError("reference ttadec rejected the stream: %s" % r.stderr[-300:])
        return r.stdout


def tta_orc():
    global _tta
    if _tta is None:
        path = os.path.join(ROOT, "oracle", "liboracle_tta.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle_tta.so"], check=True)
        L = C.CDLL(path)
        L.tta_oracle_encode_file.restype = C.c_uint64
        L.tta_oracle_encode_file.argtypes = [C.c_char_p, C.c_uint64, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_void_p)]
        L.tta_oracle_encode_frames.restype = C.c_uint
        L.tta_oracle_encode_frames.argtypes = [C.c_char_p, C.c_uint64, C.c_uint, C.c_uint, C.c_uint, C.POINTER(C.c_uint32),
                                               C.c_uint, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.tta_oracle_free.argtypes = [C.c_void_p]
        _tta = L
    return _tta


def oracle_tta_file(pcm, sample_rate, channels, bits_per_sample):
    """whole .tta file from the CPU oracle (oracle/tta_oracle.c)"""
    n = len(pcm) // (channels * (bits_per_sample // 8))
    out = C.c_void_p()
    ln = tta_orc().tta_oracle_encode_file(bytes(pcm), n, sample_rate, channels, bits_per_sample, C.byref(out))
    data = C.string_at(out, ln)
    tta_orc().tta_oracle_free(out)
    return data


def oracle_tta_frames(pcm, sample_rate, channels, bits_per_sample, frame_lengths=None):
    """(frame bytes, [sizes]) from the CPU oracle; frame_lengths: the reader's read sizes"""
    n = len(pcm) // (channels * (bits_per_sample // 8))
    block = (sample_rate * 256) // 245
    cap = (len(frame_lengths) if frame_lengths else (n + block - 1) // block) + 1
    sizes = (C.c_uint32 * cap)()
    lens = (C.c_uint32 * len(frame_lengths))(*frame_lengths) if frame_lengths else None
    out, nb = C.c_void_p(), C.c_uint64(0)
    nf = tta_orc().tta_oracle_encode_frames(bytes(pcm), n, sample_rate, channels, bits_per_sample, lens,
                                            len(frame_lengths) if frame_lengths else 0, C.byref(out), C.byref(nb), sizes)
    data = C.string_at(out, nb.value) if nb.value else b""
    tta_orc().tta_oracle_free(out)
    return data, list(sizes[:nf])


# ---- ALAC: the compiled reference (oracle/_ref/alacenc) and the plain-C oracle ----
REF_ALACENC = os.path.join(ROOT, "oracle", "_ref", "alacenc")
_alac = None


def have_alac_ref():
    return os.path.exists(REF_ALACENC)


def ref_alac_encode(pcm, channels, bits_per_sample, block_size=4096):
    """the mdat atom from the COMPILED REFERENCE ALAC encoder (src/encoders/alac.c, -DSTANDALONE; history 10/40, k 14)"""
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.m4a")
        subprocess.run([REF_ALACENC, "-c", str(channels), "-r", "44100", "-b", str(bits_per_sample), "-B", str(block_size), out],
                       input=pcm, stdout=subprocess.DEVNULL, check=True)
        with open(out, "rb") as fh:
            return fh.read()


def alac_orc():
    global _alac
    if _alac is None:
        path = os.path.join(ROOT, "oracle", "liboracle_alac.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle_alac.so"], check=True)
        L = C.CDLL(path)
        L.alac_oracle_encode_mdat.restype = C.c_uint64
        L.alac_oracle_encode_mdat.argtypes = [C.c_char_p, C.c_uint64] + [C.c_uint] * 6 + [C.POINTER(C.c_void_p)]
        L.alac_oracle_encode_framesets.restype = C.c_uint
        L.alac_oracle_encode_framesets.argtypes = [C.c_char_p, C.c_uint64] + [C.c_uint] * 8 + [
            C.POINTER(C.c_uint32), C.c_uint, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.alac_oracle_free.argtypes = [C.c_void_p]
        _alac = L
    return _alac


def oracle_alac_mdat(pcm, channels, bits_per_sample, block_size=4096, initial_history=10, history_multiplier=40, maximum_k=14):
    n = len(pcm) // (channels * (bits_per_sample // 8))
    out = C.c_void_p()
    ln = alac_orc().alac_oracle_encode_mdat(bytes(pcm), n, channels, bits_per_sample, block_size, initial_history,
                                            history_multiplier, maximum_k, C.byref(out))
    data = C.string_at(out, ln)
    alac_orc().alac_oracle_free(out)
    return data


def oracle_alac_framesets(pcm, channels, bits_per_sample, block_size=4096, initial_history=10, history_multiplier=40,
                          maximum_k=14, min_leftweight=0, max_leftweight=4, frame_lengths=None):
    """(frameset bytes, [sizes]) from the CPU oracle; frame_lengths: the reader's read sizes"""
    n = len(pcm) // (channels * (bits_per_sample // 8))
    cap = (len(frame_lengths) if frame_lengths else (n + block_size - 1) // block_size) + 1
    sizes = (C.c_uint32 * cap)()
    lens = (C.c_uint32 * len(frame_lengths))(*frame_lengths) if frame_lengths else None
    out, nb = C.c_void_p(), C.c_uint64(0)
    nf = alac_orc().alac_oracle_encode_framesets(bytes(pcm), n, channels, bits_per_sample, block_size, initial_history,
                                                 history_multiplier, maximum_k, min_leftweight, max_leftweight, lens,
                                                 len(frame_lengths) if frame_lengths else 0, C.byref(out), C.byref(nb), sizes)
    data = C.string_at(out, nb.value) if nb.value else b""
    alac_orc().alac_oracle_free(out)
    return data, list(sizes[:nf])


# ---- generators restating the reference's test streams (test/test_streams.py) ----
def sine_pcm(bits_per_sample, channels, n_frames, sample_rate, freqs_amps):
    """integer sines in the spirit of test_streams.Sine16_Stereo etc. (src/decoders/sine.c):
    per channel a sum of two sines, a1*sin(2 pi f1 t) + a2*sin(2 pi f2 t), full scale = 2^(bps-1)-1"""
    t = np.arange(n_frames, dtype=np.float64)
    full = (1 << (bits_per_sample - 1)) - 1
    chans = []
    for c in range(channels):
        f1, a1, f2, a2 = freqs_amps[c % len(freqs_amps)]
        v = a1 * np.sin(2 * np.pi * f1 * t / sample_rate) + a2 * np.sin(2 * np.pi * f2 * t / sample_rate)
        chans.append(np.clip(np.round(v * full), -full - 1, full).astype(np.int32))
    return pack_pcm(np.stack(chans, axis=1), bits_per_sample)


def wasted_bps16(n_frames):
    """test_streams.WastedBPS16 (test/test_streams.py:343-370): L=(i%2000)<<2, R=(i%1000)<<3"""
    i = np.arange(n_frames, dtype=np.int64)
    left = ((i % 2000) << 2).astype(np.int32)
    right = ((i % 1000) << 3).astype(np.int32)
    return pack_pcm(np.stack([left, right], axis=1), 16)


def full_scale_patterns(bits_per_sample):
    """test_streams PATTERN01..07 (test/test_streams.py:423-448): +/- full-scale deflections"""
    hi, lo = (1 << (bits_per_sample - 1)) - 1, -(1 << (bits_per_sample - 1))
    return [[hi, lo], [hi, hi, lo], [hi, hi, lo, lo], [hi, lo, lo], [hi, lo, hi, hi, lo, lo],
            [hi, hi, hi, lo, lo], [hi, lo, lo, lo]]


def streaminfo(flac):
    """parse STREAMINFO of a file image"""
    assert flac[:4] == b"fLaC"
    b = flac[8:42]
    v = int.from_bytes(b[10:18], "big")
    return dict(min_block=int.from_bytes(b[0:2], "big"), max_block=int.from_bytes(b[2:4], "big"),
                min_frame=int.from_bytes(b[4:7], "big"), max_frame=int.from_bytes(b[7:10], "big"),
                sample_rate=v >> 44, channels=((v >> 41) & 7) + 1, bits_per_sample=((v >> 36) & 31) + 1,
                total_samples=v & ((1 << 36) - 1), md5=bytes(b[18:34]))


def first_frame_offset(flac):
    """byte offset of the first frame (after all metadata blocks)"""
    pos = 4
    while True:
        last = flac[pos] >> 7
        ln = int.from_bytes(flac[pos + 1:pos + 4], "big")
        pos += 4 + ln
        if last:
            return pos
