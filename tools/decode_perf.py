#!/usr/bin/env python
"""Device-resident throughput of the GPU FLAC decoder on the bench workload: the hour of
stereo is encoded on the device, the frames stay in HBM and b200flac_decode_device decodes them back; the
PCM is compared with the input on the device side by downloading both.  Beside it: the reference decoder
(oracle/_ref/flacdec) on one host core over a sample.   python tools/decode_perf.py [seconds]"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
    rate, ch, bps, bs = 44100, 2, 16, 4096
    n = int(seconds * rate)
    L = b200flac.lib()
    p = b200flac.make_params(rate, ch, bps, block_size=bs, max_lpc_order=12, max_residual_partition_order=6,
                             adaptive_mid_side=True)
    enc = b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=n, n_slots=1)
    nbytes = n * ch * (bps // 8)
    cap = enc.output_bound(n, 1)
    d_pcm = L.b200flac_device_alloc(0, nbytes)
    d_out = L.b200flac_device_alloc(0, cap)
    d_dec = L.b200flac_device_alloc(0, nbytes + 64)
    L.b200flac_device_synth_pcm(0, d_pcm, 1235, ch, bps, 0, n)
    out_bytes, nfr, ms = enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
    info = b200flac.StreamInfo()
    info.min_block_size = info.max_block_size = bs
    info.sample_rate, info.channels, info.bits_per_sample, info.total_pcm_frames = rate, ch, bps, n
    kms = (C.c_float * 3)()
    nf = C.c_uint64(0)
    best = None
    for _ in range(4):
        t0 = time.perf_counter()
        rc = L.b200flac_decode_device(C.byref(info), d_out, out_bytes, 0, d_dec, nbytes, C.byref(nf), kms)
        dt = time.perf_counter() - t0
        if rc:
            raise RuntimeError(L.b200flac_last_error().decode())
        best = dt if best is None else min(best, dt)
    a = np.empty(nbytes, dtype=np.uint8)
    b = np.empty(nbytes, dtype=np.uint8)
    L.b200flac_device_download(0, a.ctypes.data, d_pcm, nbytes)
    L.b200flac_device_download(0, b.ctypes.data, d_dec, nbytes)
    same = bool((a == b).all())
    print("GPU decode, frames and PCM resident: %d frames of %.0f s stereo, %.1f ms wall (scan %.2f ms, frames %.2f ms, "
          "chain+emit %.2f ms) = %.1f Msamples/s; PCM identical to the encoder's input: %s" % (
              nf.value, seconds, best * 1e3, kms[0], kms[1], kms[2], n * ch / best / 1e6, same))
    algo = nbytes + out_bytes
    print("  algorithmic bytes (frames in + PCM out) %.0f MB -> %.0f GB/s over the frame-decode kernel" % (
        algo / 1e6, algo / (kms[1] * 1e-3) / 1e9))
    # the reference decoder on one host core over 120 s of the same stream shape
    ref = os.path.join(ROOT, "oracle", "_ref", "flacdec")
    if os.path.exists(ref):
        m = 120 * rate
        pcm = b200flac.synth_pcm(1235, ch, bps, m)
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
            path = os.path.join(d, "s.flac")
            b200flac.encode_file(path, p, pcm, m)
            t0 = time.perf_counter()
            r = subprocess.run([ref, path], stdout=subprocess.PIPE)
            dt = time.perf_counter() - t0
            print("reference flacdec, one core, 120 s of the same signal: %.0f ms = %.1f Msamples/s (output identical: %s)" % (
                dt * 1e3, m * ch / dt / 1e6, r.stdout == pcm))
            flac = open(path, "rb").read()
            t0 = time.perf_counter()
            info2, got = b200flac.decode(flac)
            dt = time.perf_counter() - t0
            print("b200flac_decode_memory (host file image -> host PCM, MD5 checked), same file: %.0f ms = %.1f Msamples/s (%s)" % (
                dt * 1e3, m * ch / dt / 1e6, got == pcm))


if __name__ == "__main__":
    main()
