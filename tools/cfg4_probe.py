#!/usr/bin/env python
"""Developer probe: BASELINE config 4 (96 kHz/24-bit 5.1, block 4608) device resident: wall clock per call against the
CUDA-event time of its kernels.   python tools/cfg4_probe.py [seconds ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402


def main():
    L = b200flac.lib()
    for seconds in [float(a) for a in sys.argv[1:]] or [120.0, 600.0]:
        rate, ch, bps = 96000, 6, 24
        n = int(seconds * rate) // 4608 * 4608
        p = b200flac.make_params(rate, ch, bps, block_size=4608, max_lpc_order=12, max_residual_partition_order=6)
        enc = b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=n, n_slots=1)
        cap = enc.output_bound(n, 1)
        d_pcm = L.b200flac_device_alloc(0, n * ch * 3)
        d_out = L.b200flac_device_alloc(0, cap)
        L.b200flac_device_synth_pcm(0, d_pcm, 1234, ch, bps, 0, n)
        for i in range(5):
            t0 = time.perf_counter()
            out_bytes, nfr, ms = enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
            wall = 1000 * (time.perf_counter() - t0)
            if i >= 2:
                print("%5.0f s: wall %.2f ms, events %.2f ms, kernels %s, out/in %.3f" % (
                    seconds, wall, ms, " ".join("%.3f" % v for v in enc.kernel_ms(0)), out_bytes / (n * ch * 3)))
        L.b200flac_device_free(0, d_pcm)
        L.b200flac_device_free(0, d_out)
        enc.close()


if __name__ == "__main__":
    main()
