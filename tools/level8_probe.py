#!/usr/bin/env python
"""Developer probe: the FlacAudio.from_pcm default (level 8: -m -e, lpc 12, partition order 6) and BASELINE config 3
(96 kHz/24-bit, -m -e, partition order 8), device resident.   python tools/level8_probe.py [seconds] [which]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402

CONFIGS = {
    "level8": (44100, 2, 16, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True, exhaustive_model_search=True)),
    "cfg3": (96000, 2, 24, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, mid_side=True, exhaustive_model_search=True)),
    "cfg3r6": (96000, 2, 24, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True, exhaustive_model_search=True)),
    "l8_24": (96000, 2, 24, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True)),
}


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    which = sys.argv[2:] or ["level8", "cfg3"]
    L = b200flac.lib()
    for name in which:
        rate, ch, bps, o = CONFIGS[name]
        n = int(seconds * rate) // 4096 * 4096
        p = b200flac.make_params(rate, ch, bps, **o)
        enc = b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=n, n_slots=1)
        nbytes = n * ch * (bps // 8)
        cap = enc.output_bound(n, 1)
        d_pcm = L.b200flac_device_alloc(0, nbytes)
        d_out = L.b200flac_device_alloc(0, cap)
        L.b200flac_device_synth_pcm(0, d_pcm, 1234, ch, bps, 0, n)
        for _ in range(3):
            out_bytes, nfr, ms = enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
        print("%-8s %6.0f s: %8.1f Msamples/s  kernels ms %s" % (name, seconds, n * ch / ms / 1e3, " ".join("%.3f" % v for v in enc.kernel_ms(0))))
        L.b200flac_device_free(0, d_pcm)
        L.b200flac_device_free(0, d_out)
        enc.close()


if __name__ == "__main__":
    main()
