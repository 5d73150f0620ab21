#!/usr/bin/env python
"""Device-resident throughput of the engine on the other BASELINE.json configs (developer tool;
bench.py measures the headline config only).  python tools/perf_configs.py [seconds]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402

CONFIGS = [
    ("cfg2 44.1k/16 stereo l12 R6 -M", 44100, 2, 16, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)),
    ("level8 44.1k/16 stereo l12 R6 -m -e", 44100, 2, 16, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True, exhaustive_model_search=True)),
    ("cfg3 96k/24 stereo l12 R8 -m -e", 96000, 2, 24, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, mid_side=True, exhaustive_model_search=True)),
    ("cfg4 96k/24 6ch B4608 l12 R6", 96000, 6, 24, dict(block_size=4608, max_lpc_order=12, max_residual_partition_order=6)),
    ("level0 44.1k/16 stereo B1152 l0 R3", 44100, 2, 16, dict(block_size=1152, max_lpc_order=0, max_residual_partition_order=3)),
    ("level5 44.1k/16 stereo l8 R5 -m", 44100, 2, 16, dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=5, mid_side=True)),
]


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    L = b200flac.lib()
    for name, rate, ch, bps, o in CONFIGS:
        n = int(seconds * rate)
        n -= n % o["block_size"]
        p = b200flac.make_params(rate, ch, bps, **o)
        enc = b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=n, n_slots=1)
        nbytes = n * ch * (bps // 8)
        cap = enc.output_bound(n, 1)
        d_pcm = L.b200flac_device_alloc(0, nbytes)
        d_out = L.b200flac_device_alloc(0, cap)
        L.b200flac_device_synth_pcm(0, d_pcm, 1234, ch, bps, 0, n)
        for _ in range(2):
            enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
        out_bytes, nfr, ms = enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
        k = enc.kernel_ms(0)
        print("%-40s %8.1f Msamples/s  ratio %.3f  kernels ms %s" % (
            name, n * ch / ms / 1e3, out_bytes / nbytes, " ".join("%.2f" % v for v in k)))
        L.b200flac_device_free(0, d_pcm)
        L.b200flac_device_free(0, d_out)
        enc.close()


if __name__ == "__main__":
    main()
