"""Golden cases: deterministic inputs + reference options.  Inputs restate the streams the
reference's own tests feed encode_flac (test/test_formats.py:3623-3877, test/test_streams.py)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import helpers  # noqa: E402

# the 9 compression levels of FlacAudio.from_pcm (audiotools/flac.py:1719-1764)
LEVELS = {
    "0": dict(block_size=1152, max_lpc_order=0, max_residual_partition_order=3),
    "1": dict(block_size=1152, max_lpc_order=0, adaptive_mid_side=True, max_residual_partition_order=3),
    "2": dict(block_size=1152, max_lpc_order=0, exhaustive_model_search=True, max_residual_partition_order=3),
    "3": dict(block_size=4096, max_lpc_order=6, max_residual_partition_order=4),
    "4": dict(block_size=4096, max_lpc_order=8, adaptive_mid_side=True, max_residual_partition_order=4),
    "5": dict(block_size=4096, max_lpc_order=8, mid_side=True, max_residual_partition_order=5),
    "6": dict(block_size=4096, max_lpc_order=8, mid_side=True, max_residual_partition_order=6),
    "7": dict(block_size=4096, max_lpc_order=8, mid_side=True, exhaustive_model_search=True,
              max_residual_partition_order=6),
    "8": dict(block_size=4096, max_lpc_order=12, mid_side=True, exhaustive_model_search=True,
              max_residual_partition_order=6),
}

CASES = []


def _add(name, gen, rate, channels, bps, options):
    CASES.append(dict(name=name, gen=gen, rate=rate, channels=channels, bps=bps, options=options))


# every compression level on the synthetic stereo signal (test_option_variations)
for lvl, o in LEVELS.items():
    _add("synth16_level%s" % lvl, ("synth", 1234, 20000), 44100, 2, 16, o)
# 24-bit, multichannel, 8-bit
_add("synth24_stereo_level8", ("synth", 1236, 12000), 96000, 2, 24, dict(LEVELS["8"], max_residual_partition_order=8))
_add("synth24_6ch_4608", ("synth", 1237, 12000), 96000, 6, 24,
     dict(block_size=4608, max_lpc_order=12, max_residual_partition_order=6))
_add("synth8_mono", ("synth", 1238, 9000), 8000, 1, 8, dict(block_size=256, max_lpc_order=6, max_residual_partition_order=4))
_add("synth16_8ch", ("synth", 1239, 6000), 48000, 8, 16, dict(block_size=1024, max_lpc_order=8, max_residual_partition_order=5))
# test_small_files (test_streams.Generate01..04)
_add("small_1", ("list", [-32768]), 44100, 1, 16, LEVELS["8"])
_add("small_2", ("list", [-32768, 32767]), 44100, 1, 16, LEVELS["8"])
_add("small_5", ("list", [-25, 0, 25, 50, 100]), 44100, 1, 16, LEVELS["8"])
_add("small_stereo10", ("list", [0, 0, 1, -1, 2, -2, 4, -4, 8, -8, 16, -16, 32, -32, 64, -64, 128, -128, 256, -256]),
     44100, 2, 16, LEVELS["8"])
# test_full_scale_deflection
for bps in (8, 16, 24):
    for i, pat in enumerate(helpers.full_scale_patterns(bps)):
        if i in (0, 3, 5):
            _add("fullscale%d_p%d" % (bps, i + 1), ("list", pat * 100), 44100, 1, bps, LEVELS["8"])
# test_wasted_bps
_add("wasted_bps16", ("wasted", 5000), 44100, 2, 16, LEVELS["8"])
# test_sines flavour
_add("sine16_stereo", ("sine", 20000, [(441.0, 0.50, 441.0, 0.49), (4410.0, 0.30, 8820.0, 0.10)]), 44100, 2, 16, LEVELS["6"])
_add("sine24_stereo", ("sine", 20000, [(441.0, 0.61, 661.5, 0.37), (882.0, 0.29, 1323.0, 0.17)]), 96000, 2, 24, LEVELS["8"])
# test_blocksizes: 32 random samples, tiny blocks, orders larger than the block
for bs, lpc in ((16, 32), (17, 8), (32, 31), (33, 16)):
    _add("tiny_bs%d_l%d" % (bs, lpc), ("random", 7, 32, 16), 44100, 1, 16,
         dict(block_size=bs, max_lpc_order=lpc, max_residual_partition_order=6, mid_side=True))
# test_frame_header_variations
_add("hdr_bs16", ("synth", 1240, 200), 44100, 2, 16, dict(block_size=16, max_lpc_order=4, max_residual_partition_order=2))
_add("hdr_bs65535", ("synth", 1241, 70000), 44100, 1, 16, dict(block_size=65535, max_lpc_order=8, max_residual_partition_order=5))
for rate in (9, 90, 90000):
    _add("hdr_rate%d" % rate, ("synth", 1242, 3000), rate, 2, 16, dict(block_size=1152, max_lpc_order=8, max_residual_partition_order=4))
# test_noise_silence
_add("noise16_stereo", ("random", 11, 20000, 16), 44100, 2, 16, LEVELS["8"])
_add("noise24_4ch_bs32", ("random", 12, 2000, 24), 44100, 4, 24, dict(block_size=32, max_lpc_order=8, max_residual_partition_order=5))
_add("silence16", ("list", [0] * 20000), 44100, 2, 16, LEVELS["8"])
_add("noise8_bs32768", ("random", 13, 40000, 8), 44100, 1, 8, dict(block_size=32768, max_lpc_order=12, max_residual_partition_order=15))
# test_fractional: lengths straddling a block boundary
for n in (4095, 4097):
    _add("fractional_%d" % n, ("synth", 1243, n), 44100, 2, 16, dict(block_size=2048, max_lpc_order=8, max_residual_partition_order=6, adaptive_mid_side=True))
# many frames so the UTF-8 frame number grows to 2 and 3 bytes (H7)
_add("framenum_utf8", ("synth", 1244, 16 * 2100), 44100, 1, 8, dict(block_size=16, max_lpc_order=2, max_residual_partition_order=1))
# BASELINE.json config #1 as written: the PCM of the reference's own fixture test/1m.flac (60 s of 44.1 kHz/16-bit
# stereo, 2,646,000 PCM frames = 645 full blocks + 4080; the fixture is digital silence, so every subframe is
# CONSTANT and the frame-number field grows from 1 to 2 bytes at frame 128), block_size 4096, max_lpc_order 8
for r in (3, 6):
    _add("config1_1m_flac_R%d" % r, ("flacfile", "1m.flac"), 44100, 2, 16,
         dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=r))


def case_pcm(case):
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
    if g[0] == "flacfile":
        # decoded by the compiled reference decoder where it exists (that is how the golden manifest was made),
        # else by the engine's own GPU decoder; the manifest's pcm_sha256 pins either
        with open(os.path.join(helpers.GOLDEN, "flac", g[1]), "rb") as fh:
            data = fh.read()
        if helpers.have_ref():
            return helpers.ref_decode(data)
        import b200flac
        return b200flac.decode(data)[1]
    if g[0] == "random":
        rng = np.random.RandomState(g[1])
        lo, hi = -(1 << (g[3] - 1)), (1 << (g[3] - 1))
        return helpers.pack_pcm(rng.randint(lo, hi, size=g[2] * ch).astype(np.int32), bps)
    raise ValueError(g)
