"""Deterministic inputs for the TTA golden manifest: the shapes the reference's own TTA tests feed the encoder
(test/test_formats.py TTAFileTest: small files, full-scale deflection, sines, noise, silence, fractional frames,
several channel counts and sample widths)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers  # noqa: E402

TTA_CASES = []


def _add(name, gen, rate, channels, bps):
    TTA_CASES.append(dict(name=name, gen=gen, rate=rate, channels=channels, bps=bps))


_add("synth16_stereo", ("synth", 1301, 100000), 44100, 2, 16)          # 2 full frames + a tail
_add("synth16_mono", ("synth", 1302, 50000), 44100, 1, 16)
_add("synth24_6ch", ("synth", 1303, 110000), 96000, 6, 24)             # one full frame (100,310) + a tail
_add("synth8_mono", ("synth", 1304, 9000), 8000, 1, 8)
_add("synth16_3ch", ("synth", 1305, 60000), 48000, 3, 16)
_add("synth24_stereo_2frames", ("synth", 1306, 2 * 100310), 96000, 2, 24)   # exactly two frames, no tail
_add("one_sample", ("list", [-32768, 32767]), 44100, 2, 16)
_add("small_5", ("list", [-25, 0, 25, 50, 100]), 44100, 1, 16)
for bps in (8, 16, 24):
    _add("noise%d_stereo" % bps, ("random", 21, 30000, bps), 44100, 2, bps)
    _add("fullscale%d" % bps, ("square", 30000, bps), 44100, 2, bps)
_add("silence16", ("list", [0] * 100000), 44100, 2, 16)
_add("sine16_stereo", ("sine", 60000, [(441.0, 0.50, 441.0, 0.49), (4410.0, 0.30, 8820.0, 0.10)]), 44100, 2, 16)
_add("wasted_bps16", ("wasted", 50000), 44100, 2, 16)
_add("synth16_8ch", ("synth", 1307, 20000), 22050, 8, 16)


def tta_case_pcm(case):
    g = case["gen"]
    ch, bps = case["channels"], case["bps"]
    if g[0] == "synth":
        return helpers.synth_pcm(g[1], ch, bps, g[2])
    if g[0] == "list":
        return helpers.pack_pcm(np.array(g[1], dtype=np.int32), bps)
    if g[0] == "wasted":
        return helpers.wasted_bps16(g[1])
    if g[0] == "sine":
        return helpers.sine_pcm(bps, ch, g[1], case["rate"], g[2])
    if g[0] == "random":
        rng = np.random.RandomState(g[1])
        lo, hi = -(1 << (g[3] - 1)), (1 << (g[3] - 1))
        return helpers.pack_pcm(rng.randint(lo, hi, size=g[2] * ch).astype(np.int32), bps)
    if g[0] == "square":
        lo, hi = -(1 << (g[2] - 1)), (1 << (g[2] - 1)) - 1
        return helpers.pack_pcm(np.tile(np.array([hi, lo], dtype=np.int32), g[1] * ch // 2), bps)
    raise ValueError(g)
