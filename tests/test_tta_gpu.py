"""GPU parity of the TTA encoder (python-audio-tools_b200/csrc/b200tta.cu) through the C ABI: whole files and
frame lists byte-identical to the CPU oracle (itself pinned to the compiled reference, tests/test_tta_oracle.py)
and to the golden manifest made from the reference binary."""
import hashlib
import json
import os

import numpy as np
import pytest

import helpers
from golden.tta_cases import TTA_CASES, tta_case_pcm

pytestmark = pytest.mark.gpu

with open(os.path.join(helpers.GOLDEN, "tta_golden.json")) as _fh:
    GOLD = {c["name"]: c for c in json.load(_fh)["cases"]}


@pytest.mark.parametrize("case", TTA_CASES, ids=[c["name"] for c in TTA_CASES])
def test_tta_file_equals_reference_golden(case, tmp_path, built):
    import b200tta
    g = GOLD[case["name"]]
    pcm = tta_case_pcm(case)
    n = len(pcm) // (case["channels"] * (case["bps"] // 8))
    path = os.path.join(str(tmp_path), "o.tta")
    b200tta.encode_file(path, pcm, n, case["rate"], case["channels"], case["bps"])
    data = open(path, "rb").read()
    assert len(data) == g["length"] and hashlib.sha256(data).hexdigest() == g["sha256"]


@pytest.mark.parametrize("rate,ch,bps,n", [(44100, 2, 16, 46080 * 7 + 1234), (96000, 6, 24, 100310 * 2 + 5), (8000, 1, 8, 30000),
                                            (48000, 3, 16, 150000), (44100, 2, 24, 46080), (44100, 8, 16, 50000)])
def test_tta_frames_equal_oracle(rate, ch, bps, n, built):
    import b200tta
    pcm = helpers.synth_pcm(700 + ch + bps, ch, bps, n)
    got, sizes, ms = b200tta.encode_frames(pcm, n, rate, ch, bps)
    want, want_sizes = helpers.oracle_tta_frames(pcm, rate, ch, bps)
    assert sizes == want_sizes
    assert got == want
    if helpers.have_tta_ref():
        assert helpers.ref_tta_decode(helpers.oracle_tta_file(pcm, rate, ch, bps)) == pcm


def test_tta_short_reads_and_odd_alignment(built):
    """frame lengths as a reader's short reads give them (tta.c:69-83): frames of 1, 3, 4097 ... PCM frames, so
    frame starts fall on every byte alignment of the output and descriptor planes are padded"""
    import b200tta
    rate, ch, bps = 44100, 2, 16
    lens = [1, 3, 4097, 46080, 2, 46079, 5, 1000, 7]
    pcm = helpers.synth_pcm(31, ch, bps, sum(lens))
    got, sizes, _ = b200tta.encode_frames(pcm, sum(lens), rate, ch, bps, frame_lengths=lens)
    want, want_sizes = helpers.oracle_tta_frames(pcm, rate, ch, bps, lens)
    assert sizes == want_sizes and got == want


def test_tta_long_unary_runs(built):
    """full-scale noise after silence: the adaptive Rice parameter starts at 10 (tta.c:193-196), so the first
    loud samples cost thousands of one-bits each -- the packer's run-of-ones path"""
    import b200tta
    rng = np.random.RandomState(5)
    for bps in (16, 24):
        lo, hi = -(1 << (bps - 1)), (1 << (bps - 1))
        x = np.concatenate([np.zeros(2000, dtype=np.int32), rng.randint(lo, hi, size=6000).astype(np.int32),
                            np.zeros(3000, dtype=np.int32), np.tile(np.array([hi - 1, lo], dtype=np.int32), 2000)])
        pcm = helpers.pack_pcm(x, bps)
        n = len(x) // 2
        got, sizes, _ = b200tta.encode_frames(pcm, n, 44100, 2, bps)
        want, want_sizes = helpers.oracle_tta_frames(pcm, 44100, 2, bps)
        assert sizes == want_sizes and got == want


def test_tta_empty_and_errors(built, tmp_path):
    import b200tta
    got, sizes, _ = b200tta.encode_frames(b"", 0, 44100, 2, 16)
    assert got == b"" and sizes == []
    with pytest.raises(b200tta.B200TtaError):
        b200tta.encode_frames(b"\0" * 12, 2, 44100, 2, 12)
    with pytest.raises(b200tta.B200TtaError):
        b200tta.encode_frames(b"\0" * 40, 10, 44100, 2, 16, frame_lengths=[4, 4])
    with pytest.raises(b200tta.B200TtaError):
        b200tta.encode_file(os.path.join(str(tmp_path), "no", "dir", "x.tta"), b"\0" * 40, 10, 44100, 2, 16)


def test_encode_tta_python_entry(built, tmp_path):
    """audiotools.encoders.encode_tta(file, pcmreader) (src/encoders/tta.c:31-117): frames to the file object,
    the list of frame sizes back; every read() of the reader is one frame -- a reader with short reads included"""
    import io
    import audiotools
    from audiotools import encoders
    rate, ch, bps, n = 44100, 2, 16, 46080 * 2 + 999
    pcm = helpers.synth_pcm(55, ch, bps, n)
    f = io.BytesIO()
    sizes = encoders.encode_tta(f, audiotools.BufferedPCMReader(audiotools.PCMBytesReader(pcm, rate, ch, 0x3, bps)))
    want, want_sizes = helpers.oracle_tta_frames(pcm, rate, ch, bps)
    assert f.getvalue() == want and sizes == want_sizes
    # keyword form, and a reader that hands out 1000 frames at a time whatever it is asked for
    class Short(audiotools.PCMBytesReader):
        def read(self, pcm_frames):
            return audiotools.PCMBytesReader.read(self, min(pcm_frames, 1000))
    f2 = io.BytesIO()
    sizes2 = encoders.encode_tta(file=f2, pcmreader=Short(pcm[:4 * 3500], rate, ch, 0x3, bps))
    want2, want_sizes2 = helpers.oracle_tta_frames(pcm[:4 * 3500], rate, ch, bps, [1000, 1000, 1000, 500])
    assert f2.getvalue() == want2 and sizes2 == want_sizes2
    with pytest.raises(TypeError):
        encoders.encode_tta(f2)
