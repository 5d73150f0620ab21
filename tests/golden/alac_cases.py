"""Deterministic inputs for the ALAC golden manifest: the shapes the reference's ALAC tests feed the encoder
(test/test_formats.py ALACFileTest: small files, full-scale deflection, sines, noise, silence, wasted bits,
fractional frames, every channel count 1..8, 16 and 24 bits, several block sizes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers  # noqa: E402

ALAC_CASES = []


def _add(name, gen, channels, bps, block_size=4096):
    ALAC_CASES.append(dict(name=name, gen=gen, channels=channels, bps=bps, block_size=block_size))


_add("synth16_stereo", ("synth", 1401, 20000), 2, 16)
_add("synth16_mono", ("synth", 1402, 9000), 1, 16)
_add("synth24_stereo", ("synth", 1403, 12000), 2, 24)
_add("synth24_mono", ("synth", 1404, 9000), 1, 24)
for ch in (3, 4, 5, 6, 7, 8):
    _add("synth16_%dch" % ch, ("synth", 1405 + ch, 6000), ch, 16)
_add("synth24_6ch", ("synth", 1420, 9000), 6, 24)
_add("synth16_stereo_bs1024", ("synth", 1421, 8200), 2, 16, 1024)
_add("synth16_stereo_bs1152", ("synth", 1422, 5000), 2, 16, 1152)
_add("tiny_5", ("list", [-25, 0, 25, 50, 100, 7, -7, 3, -3, 0]), 2, 16)            # 5 frames: uncompressed frame (< 10)
_add("tiny_9_then_10", ("synth", 1423, 4096 + 9), 2, 16)                          # a 9-frame tail: uncompressed
_add("exactly_10", ("synth", 1424, 10), 1, 16)
for bps in (16, 24):
    _add("noise%d_stereo" % bps, ("random", 31, 10000, bps), 2, bps)
    _add("fullscale%d" % bps, ("square", 10000, bps), 2, bps)
    _add("silence%d" % bps, ("list", [0] * 20000), 2, bps)
_add("sine16_stereo", ("sine", 20000, [(441.0, 0.50, 441.0, 0.49), (4410.0, 0.30, 8820.0, 0.10)]), 2, 16)
_add("wasted_bps16", ("wasted", 10000), 2, 16)
_add("sparse16", ("sparse", 12000), 2, 16)       # long runs of zeros between clicks: the zero-run codes


def alac_case_pcm(case):
    g = case["gen"]
    ch, bps = case["channels"], case["bps"]
    if g[0] == "synth":
        return helpers.synth_pcm(g[1], ch, bps, g[2])
    if g[0] == "list":
        return helpers.pack_pcm(np.array(g[1], dtype=np.int32), bps)
    if g[0] == "wasted":
        return helpers.wasted_bps16(g[1])
    if g[0] == "sine":
        return helpers.sine_pcm(bps, ch, g[1], 44100, g[2])
    if g[0] == "random":
        rng = np.random.RandomState(g[1])
        lo, hi = -(1 << (g[3] - 1)), (1 << (g[3] - 1))
        return helpers.pack_pcm(rng.randint(lo, hi, size=g[2] * ch).astype(np.int32), bps)
    if g[0] == "square":
        lo, hi = -(1 << (g[2] - 1)), (1 << (g[2] - 1)) - 1
        return helpers.pack_pcm(np.tile(np.array([hi, lo], dtype=np.int32), g[1] * ch // 2), bps)
    if g[0] == "sparse":
        rng = np.random.RandomState(77)
        x = np.zeros(g[1] * ch, dtype=np.int32)
        idx = rng.randint(0, len(x), size=60)
        x[idx] = rng.randint(-3000, 3000, size=60)
        x[4000 * ch:4300 * ch] = rng.randint(-200, 200, size=300 * ch)
        return helpers.pack_pcm(x, bps)
    raise ValueError(g)
