#!/usr/bin/env python
"""Developer probe: per-kernel times of the bench hour (default options), device resident, ten repetitions.
   B200FLAC_LIB=<variant .so> python tools/ab_probe.py   -- run once per library on the SAME box to compare builds"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402


def main():
    L = b200flac.lib()
    n = 158760000
    p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)
    enc = b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=n, n_slots=1)
    cap = enc.output_bound(n, 1)
    d_pcm = L.b200flac_device_alloc(0, n * 4)
    d_out = L.b200flac_device_alloc(0, cap)
    L.b200flac_device_synth_pcm(0, d_pcm, 1235, 2, 16, 0, n)
    acc = None
    reps = 10
    for i in range(3 + reps):
        enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
        k = enc.kernel_ms(0)
        if i >= 3:
            acc = k if acc is None else [a + b for a, b in zip(acc, k)]
    print(os.environ.get("B200FLAC_LIB", "default"), " ".join("%.3f" % (v / reps) for v in acc), "sum %.3f" % (sum(acc) / reps))
    enc.close()


if __name__ == "__main__":
    main()
