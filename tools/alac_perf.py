#!/usr/bin/env python
"""Developer tool: device-resident throughput of the ALAC encoder (b200alac_encode_device) on synthetic stereo,
next to the compiled reference's `alacenc` on one host core.   python tools/alac_perf.py [seconds]"""
import ctypes as C
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200alac  # noqa: E402
import b200flac  # noqa: E402


def measure(seconds, rate=44100, ch=2, bps=16, reps=3, device=0):
    L = b200alac.lib()
    n = int(seconds * rate)
    nbytes = n * ch * (bps // 8)
    p = b200alac.make_params(ch, bps)
    nf = (n + 4095) // 4096
    cap = L.b200alac_output_bound(C.byref(p), n, nf)
    d_pcm = L.b200flac_device_alloc(device, nbytes)
    d_out = L.b200flac_device_alloc(device, cap)
    L.b200flac_device_synth_pcm(device, d_pcm, 1235, ch, bps, 0, n)
    ms = (C.c_float * 4)()
    out_bytes, nfr = C.c_uint64(0), C.c_uint32(0)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        if L.b200alac_encode_device(C.byref(p), d_pcm, n, device, d_out, cap, C.byref(out_bytes), None, C.byref(nfr), ms):
            raise SystemExit(L.b200alac_last_error().decode())
        wall = time.perf_counter() - t0
        k = list(ms)
        if best is None or sum(k) < sum(best[0]):
            best = (k, wall)
    L.b200flac_device_free(device, d_pcm)
    L.b200flac_device_free(device, d_out)
    k, wall = best
    return {"seconds": seconds, "samples": n * ch, "framesets": nfr.value, "out_bytes": out_bytes.value,
            "compressed_ratio": out_bytes.value / nbytes,
            "kernel_ms": {"model": k[0], "sizing": k[1], "select_scan": k[2], "emit": k[3]},
            "kernels_ms": sum(k), "wall_ms": wall * 1e3, "value": n * ch / (sum(k) * 1e-3) / 1e6, "unit": "Msamples/s",
            "algorithmic_bytes": nbytes + out_bytes.value}


def reference_single_core(seconds=30.0, rate=44100, ch=2, bps=16):
    ref = os.path.join(ROOT, "oracle", "_ref", "alacenc")
    if not os.path.exists(ref):
        return None
    n = int(seconds * rate)
    pcm = b200flac.synth_pcm(1235, ch, bps, n)
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else None
    with tempfile.TemporaryDirectory(dir=shm) as d:
        path = os.path.join(d, "i.pcm")
        open(path, "wb").write(pcm)
        t0 = time.perf_counter()
        subprocess.run([ref, "-c", str(ch), "-r", str(rate), "-b", str(bps), os.path.join(d, "o.m4a")],
                       stdin=open(path, "rb"), stdout=subprocess.DEVNULL, check=True)
        dt = time.perf_counter() - t0
    return {"value": n * ch / dt / 1e6, "unit": "Msamples/s", "seconds": seconds, "cores": 1,
            "what": "oracle/_ref/alacenc (reference src/encoders/alac.c) on one host core, PCM and output on tmpfs"}


if __name__ == "__main__":
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
    r = measure(secs)
    print("ALAC %.0f s stereo: kernels %.3f ms (model %.3f, sizing %.3f, select %.3f, emit %.3f)  wall %.1f ms  %.1f Gsamples/s  ratio %.3f"
          % (secs, r["kernels_ms"], r["kernel_ms"]["model"], r["kernel_ms"]["sizing"], r["kernel_ms"]["select_scan"],
             r["kernel_ms"]["emit"], r["wall_ms"], r["value"] / 1e3, r["compressed_ratio"]))
    q = measure(180.0)
    print("ALAC 180 s stereo: kernels %.3f ms  %.1f Gsamples/s" % (q["kernels_ms"], q["value"] / 1e3))
    c = reference_single_core()
    if c:
        print("reference alacenc, one core: %.1f Msamples/s" % c["value"])
