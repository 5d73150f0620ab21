"""CPU: the oracle restatement against (a) the committed golden manifest produced by the compiled
reference (tests/golden/golden.json, script tests/golden/make_golden.py) and (b), where
oracle/_ref has been built (this container; it also travels to the GPU box), the compiled
reference itself over a wider option grid, byte for byte."""
import hashlib
import json
import os

import numpy as np
import pytest

import helpers
from golden.golden_cases import CASES, LEVELS, case_pcm

with open(os.path.join(helpers.GOLDEN, "golden.json")) as _fh:
    GOLD = {c["name"]: c for c in json.load(_fh)["cases"]}


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_golden(case, built):
    g = GOLD[case["name"]]
    pcm = case_pcm(case)
    assert hashlib.sha256(pcm).hexdigest() == g["pcm_sha256"], "input generator drifted"
    flac = helpers.oracle_encode(pcm, case["rate"], case["channels"], case["bps"],
                                 helpers.options(**case["options"]))
    ff = helpers.first_frame_offset(flac)
    assert flac[ff:ff + 64].hex() == g["first_frame_bytes"]
    assert len(flac) == g["length"]
    assert hashlib.sha256(flac).hexdigest() == g["sha256"]


def test_streaminfo_md5_and_totals(built):
    pcm = helpers.synth_pcm(5, 2, 16, 10000)
    flac = helpers.oracle_encode(pcm, 44100, 2, 16, helpers.options(**LEVELS["5"]))
    si = helpers.streaminfo(flac)
    assert si["md5"] == hashlib.md5(pcm).digest()
    assert si["total_samples"] == 10000 and si["channels"] == 2 and si["bits_per_sample"] == 16
    assert si["min_block"] == 4096 and si["max_block"] == 4096 and si["sample_rate"] == 44100


def test_offsets_list_matches_frames(built):
    pcm = helpers.synth_pcm(6, 2, 16, 4096 * 3 + 17)
    flac, offs = helpers.oracle_encode(pcm, 44100, 2, 16, helpers.options(**LEVELS["4"]), want_offsets=True)
    ff = helpers.first_frame_offset(flac)
    assert [n for _, n in offs] == [4096, 4096, 4096, 17]
    for off, _ in offs:
        assert flac[ff + off:ff + off + 2] == b"\xff\xf8"
    si = helpers.streaminfo(flac)
    sizes = [b - a for (a, _), (b, _) in zip(offs, offs[1:])] + [len(flac) - ff - offs[-1][0]]
    assert si["min_frame"] == min(sizes) and si["max_frame"] == max(sizes)


needs_ref = pytest.mark.skipif(not helpers.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("bs", list(range(16, 34)))
def test_blocksizes_against_reference(bs, built):
    # test_formats.py:3685-3712 test_blocksizes: 32 random samples, block 16..33, many LPC orders
    rng = np.random.RandomState(100 + bs)
    pcm = helpers.pack_pcm(rng.randint(-32768, 32768, size=32).astype(np.int32), 16)
    for lpc in (0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32):
        o = helpers.options(block_size=bs, max_lpc_order=lpc, max_residual_partition_order=6, mid_side=True,
                            exhaustive_model_search=bool(lpc % 2))
        assert helpers.oracle_encode(pcm, 44100, 1, 16, o) == helpers.ref_encode(pcm, 44100, 1, 16, o), (bs, lpc)


@needs_ref
@pytest.mark.parametrize("lvl", sorted(LEVELS))
@pytest.mark.parametrize("shape", [(2, 16), (2, 24), (1, 8), (6, 24)])
def test_levels_against_reference(lvl, shape, built):
    ch, bps = shape
    pcm = helpers.synth_pcm(300 + int(lvl), ch, bps, 9000)
    o = helpers.options(**LEVELS[lvl])
    want = helpers.ref_encode(pcm, 48000, ch, bps, o)
    assert helpers.oracle_encode(pcm, 48000, ch, bps, o) == want
    assert helpers.ref_decode(want) == pcm


@needs_ref
def test_partition_order_beyond_block_against_reference(built):
    # partition orders whose partitions are shorter than the predictor order (H3 underflow levels)
    pcm = helpers.synth_pcm(77, 1, 16, 4096 * 2)
    for po in (10, 12, 15):
        o = helpers.options(block_size=4096, max_lpc_order=12, max_residual_partition_order=po)
        assert helpers.oracle_encode(pcm, 44100, 1, 16, o) == helpers.ref_encode(pcm, 44100, 1, 16, o)


@needs_ref
def test_nan_lpc_block_against_reference(built):
    # only the first and last samples non-zero -> windowed energy 0 -> NaN coefficients
    s = np.zeros(4096, dtype=np.int32)
    s[0], s[-1] = 1000, -700
    pcm = helpers.pack_pcm(s, 16)
    for dis_fixed in (False,):
        o = helpers.options(block_size=4096, max_lpc_order=8, max_residual_partition_order=4)
        assert helpers.oracle_encode(pcm, 44100, 1, 16, o) == helpers.ref_encode(pcm, 44100, 1, 16, o)
