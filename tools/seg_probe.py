#!/usr/bin/env python
"""Developer probe: a level-8 batch of N three-minute tracks (N segments, each ending in a short block) against the
same PCM as one segment: what the per-track tails cost.   python tools/seg_probe.py [tracks]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402


def main():
    tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    tn = 7938000
    n = tn * tracks
    L = b200flac.lib()
    p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True,
                             exhaustive_model_search=True)
    enc = b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=n + 4096 * tracks, n_slots=1)
    cap = enc.output_bound(n, tracks)
    d_pcm = L.b200flac_device_alloc(0, n * 4)
    d_out = L.b200flac_device_alloc(0, cap)
    L.b200flac_device_synth_pcm(0, d_pcm, 1234, 2, 16, 0, n)
    for name, segs in (("one segment", [(0, n, 0)]), ("%d segments" % tracks, [(i * tn, tn, 0) for i in range(tracks)])):
        for _ in range(3):
            out_bytes, nfr, ms = enc.encode_device(d_pcm, segs, d_out, cap)
        print("%-14s %8.3f ms  kernels %s" % (name, ms, " ".join("%.3f" % v for v in enc.kernel_ms(0))))
    enc.close()


if __name__ == "__main__":
    main()

