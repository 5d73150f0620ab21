#!/usr/bin/env python
"""Small encodes through every kernel generation, for compute-sanitizer (developer tool):
    compute-sanitizer --tool racecheck python tools/sanitize_case.py
    compute-sanitizer --tool memcheck  python tools/sanitize_case.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import b200flac  # noqa: E402
import helpers  # noqa: E402

CASES = [
    (44100, 2, 16, 4096 * 3 + 100, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)),
    (96000, 2, 24, 4096 * 2 + 9, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True)),
    (48000, 2, 16, 4608 * 2 + 11, dict(block_size=4608, max_lpc_order=10, max_residual_partition_order=6, adaptive_mid_side=True)),
    (44100, 2, 16, 4096 * 2 + 7, dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True, exhaustive_model_search=True)),
    (96000, 6, 24, 4608 + 500, dict(block_size=4608, max_lpc_order=12, max_residual_partition_order=6)),
    (44100, 1, 16, 512 * 5 + 1, dict(block_size=512, max_lpc_order=4, max_residual_partition_order=2)),
]


def main():
    for i, (rate, ch, bps, n, o) in enumerate(CASES):
        opts = helpers.options(**o)
        kw = {k: v for k, v in opts.items() if k != "padding_size"}
        p = b200flac.make_params(rate, ch, bps, **kw)
        pcm = helpers.synth_pcm(10 + i, ch, bps, n)
        enc = b200flac.Encoder(p, max_pcm_frames_per_batch=n, n_slots=1)
        out, fbytes, fpcm = enc.encode(pcm, n)
        want, sizes = helpers.oracle_encode_range(pcm, rate, ch, bps, opts, 0)
        assert out.tobytes() == want, "case %d differs" % i
        enc.close()
        print("case", i, "ok", len(want), "bytes")


if __name__ == "__main__":
    main()
