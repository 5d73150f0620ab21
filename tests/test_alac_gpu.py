"""GPU parity of the ALAC encoder (python-audio-tools_b200/csrc/b200alac.cu) through the C ABI: mdat atoms and
frameset lists byte-identical to the CPU oracle (itself pinned to the compiled reference, tests/test_alac_oracle.py)
and to the golden manifest made from the reference binary."""
import hashlib
import json
import os

import numpy as np
import pytest

import helpers
from golden.alac_cases import ALAC_CASES, alac_case_pcm

pytestmark = pytest.mark.gpu

with open(os.path.join(helpers.GOLDEN, "alac_golden.json")) as _fh:
    GOLD = {c["name"]: c for c in json.load(_fh)["cases"]}


@pytest.mark.parametrize("case", ALAC_CASES, ids=[c["name"] for c in ALAC_CASES])
def test_alac_mdat_equals_reference_golden(case, tmp_path, built):
    import b200alac
    g = GOLD[case["name"]]
    pcm = alac_case_pcm(case)
    n = len(pcm) // (case["channels"] * (case["bps"] // 8))
    path = os.path.join(str(tmp_path), "o.m4a")
    b200alac.encode_mdat(path, pcm, n, b200alac.make_params(case["channels"], case["bps"], case["block_size"]))
    data = open(path, "rb").read()
    assert len(data) == g["length"] and hashlib.sha256(data).hexdigest() == g["sha256"]


@pytest.mark.parametrize("ch,bps,n,bs", [(2, 16, 4096 * 40 + 1234, 4096), (6, 24, 4096 * 5 + 5, 4096), (1, 16, 30000, 4096),
                                         (3, 16, 50000, 1024), (2, 24, 4096 * 9, 4096), (8, 16, 20000, 4096)])
def test_alac_framesets_equal_oracle(ch, bps, n, bs, built):
    import b200alac
    pcm = helpers.synth_pcm(800 + ch + bps, ch, bps, n)
    got, sizes, ms = b200alac.encode_framesets(pcm, n, b200alac.make_params(ch, bps, bs))
    want, want_sizes = helpers.oracle_alac_framesets(pcm, ch, bps, bs)
    assert sizes == want_sizes
    assert got == want


def test_alac_options_and_short_reads(built):
    """non-default history / k / leftweight range, and frameset lengths as a reader's short reads give them
    (alac.c:163-183): lengths below 10 go out uncompressed (:387), every frameset starts on its own byte"""
    import b200alac
    lens = [4096, 3, 4096, 9, 10, 11, 1000, 4096, 1]
    pcm = helpers.synth_pcm(41, 2, 16, sum(lens))
    p = b200alac.make_params(2, 16, 4096, initial_history=40, history_multiplier=20, maximum_k=10,
                             minimum_interlacing_leftweight=1, maximum_interlacing_leftweight=3)
    got, sizes, _ = b200alac.encode_framesets(pcm, sum(lens), p, frame_lengths=lens)
    want, want_sizes = helpers.oracle_alac_framesets(pcm, 2, 16, 4096, 40, 20, 10, 1, 3, frame_lengths=lens)
    assert sizes == want_sizes and got == want


def test_alac_zero_runs_and_escapes(built):
    """silence with clicks (zero-run codes, :1075-1092), full-scale noise (escape codes, :1109-1112), a constant
    offset (autocorrelation[0] != 0 with tiny errors) and digital silence (the all-zero special case, :767-776)"""
    import b200alac
    rng = np.random.RandomState(9)
    for bps in (16, 24):
        lo, hi = -(1 << (bps - 1)), (1 << (bps - 1))
        x = np.concatenate([np.zeros(9000, dtype=np.int32), rng.randint(lo, hi, size=9000).astype(np.int32),
                            np.full(9000, 1234, dtype=np.int32), np.zeros(3000, dtype=np.int32)])
        x[100] = 5; x[4000] = -7; x[4001] = 300
        pcm = helpers.pack_pcm(x, bps)
        n = len(x) // 2
        got, sizes, _ = b200alac.encode_framesets(pcm, n, b200alac.make_params(2, bps))
        want, want_sizes = helpers.oracle_alac_framesets(pcm, 2, bps)
        assert sizes == want_sizes and got == want


def test_alac_empty_and_errors(built, tmp_path):
    import b200alac
    got, sizes, _ = b200alac.encode_framesets(b"", 0, b200alac.make_params())
    assert got == b"" and sizes == []
    with pytest.raises(b200alac.B200AlacError, match="16 or 24"):
        b200alac.encode_framesets(b"\0" * 12, 6, b200alac.make_params(2, 8))
    with pytest.raises(b200alac.B200AlacError):
        b200alac.encode_framesets(b"\0" * 40, 10, b200alac.make_params(), frame_lengths=[4, 4])
    with pytest.raises(b200alac.B200AlacError):
        b200alac.encode_mdat(os.path.join(str(tmp_path), "no", "dir", "x.m4a"), b"\0" * 40, 10, b200alac.make_params())


def test_encode_alac_python_entry(built):
    """audiotools.encoders.encode_alac(file, pcmreader, block_size, initial_history, history_multiplier, maximum_k)
    (src/encoders/alac.c:30-214): the mdat atom in the file object, (frameset sizes, total PCM frames) back"""
    import io
    import audiotools
    from audiotools import encoders
    ch, bps, n = 2, 16, 4096 * 3 + 777
    pcm = helpers.synth_pcm(66, ch, bps, n)
    f = io.BytesIO(b"head")
    f.seek(4)
    sizes, total = encoders.encode_alac(f, audiotools.BufferedPCMReader(audiotools.PCMBytesReader(pcm, 44100, ch, 0x3, bps)),
                                        4096, 10, 40, 14)
    want = helpers.oracle_alac_mdat(pcm, ch, bps)
    _, want_sizes = helpers.oracle_alac_framesets(pcm, ch, bps)
    assert f.getvalue() == b"head" + want and sizes == want_sizes and total == n
    assert f.tell() == 8          # just after the rewritten size, like the reference's fsetpos + write
    with pytest.raises(ValueError, match="16 or 24"):
        encoders.encode_alac(io.BytesIO(), audiotools.PCMBytesReader(b"\0" * 100, 44100, 2, 0x3, 8), 4096, 10, 40, 14)
    with pytest.raises(TypeError):
        encoders.encode_alac(io.BytesIO(), audiotools.PCMBytesReader(pcm, 44100, ch, 0x3, bps))
