"""CPU: the ALAC oracle (oracle/alac_oracle.c) against the golden manifest made from the COMPILED REFERENCE encoder
(tests/golden/alac_golden.json, tests/golden/make_alac_golden.py) and, where oracle/_ref exists, against the
reference binary itself -- whole mdat atoms, byte for byte (reference src/encoders/alac.c)."""
import hashlib
import json
import os
import struct

import pytest

import helpers
from golden.alac_cases import ALAC_CASES, alac_case_pcm

with open(os.path.join(helpers.GOLDEN, "alac_golden.json")) as _fh:
    GOLD = {c["name"]: c for c in json.load(_fh)["cases"]}


@pytest.mark.parametrize("case", ALAC_CASES, ids=[c["name"] for c in ALAC_CASES])
def test_oracle_matches_reference_golden(case):
    g = GOLD[case["name"]]
    pcm = alac_case_pcm(case)
    assert hashlib.sha256(pcm).hexdigest() == g["pcm_sha256"], "input generator drifted"
    data = helpers.oracle_alac_mdat(pcm, case["channels"], case["bps"], case["block_size"])
    assert len(data) == g["length"] and hashlib.sha256(data).hexdigest() == g["sha256"]


@pytest.mark.skipif(not helpers.have_alac_ref(), reason="oracle/_ref/alacenc not built (needs /root/reference)")
@pytest.mark.parametrize("ch,bps,n,bs", [(2, 16, 4096 * 3 + 17, 4096), (1, 24, 12345, 4096), (6, 16, 9000, 2048),
                                         (2, 24, 10000, 4096), (8, 24, 5000, 4096), (2, 16, 30, 16)])
def test_oracle_matches_compiled_reference(ch, bps, n, bs):
    pcm = helpers.synth_pcm(950 + ch, ch, bps, n)
    assert helpers.oracle_alac_mdat(pcm, ch, bps, bs) == helpers.ref_alac_encode(pcm, ch, bps, bs)


def test_mdat_layout_and_frameset_sizes():
    """ALACEncoder_encode_alac (alac.c:95-214): atom size = 8 + the framesets' sizes, every frameset byte aligned
    and closed by '111'; the first three bits of a stereo frameset are the channel count - 1"""
    pcm = helpers.synth_pcm(3, 2, 16, 4096 * 2 + 100)
    data = helpers.oracle_alac_mdat(pcm, 2, 16)
    size, tag = struct.unpack(">I4s", data[:8])
    assert tag == b"mdat" and size == len(data)
    frames, sizes = helpers.oracle_alac_framesets(pcm, 2, 16)
    assert frames == data[8:] and sum(sizes) == len(frames) and len(sizes) == 3
    pos = 0
    for s in sizes:
        assert frames[pos] >> 5 == 1          # two channels
        pos += s


def test_short_reads_become_short_framesets():
    lens = [4096, 100, 4096, 7, 3000]
    pcm = helpers.synth_pcm(2, 2, 16, sum(lens))
    frames, sizes = helpers.oracle_alac_framesets(pcm, 2, 16, frame_lengths=lens)
    assert len(sizes) == len(lens) and sum(sizes) == len(frames)
    pos, off = 0, 0
    for n, s in zip(lens, sizes):
        one, one_s = helpers.oracle_alac_framesets(pcm[pos * 4:(pos + n) * 4], 2, 16, frame_lengths=[n])
        assert frames[off:off + s] == one and one_s == [s]
        pos += n
        off += s


def test_alac_abi_exports_every_declared_symbol(built):
    import re
    import b200alac
    import b200flac
    hdr = open(os.path.join(helpers.ROOT, "include", "b200alac.h")).read()
    names = set(re.findall(r"\b(b200alac_[a-z_0-9]+)\s*\(", hdr))
    assert names >= {"b200alac_encode_framesets", "b200alac_encode_device", "b200alac_encode_mdat", "b200alac_output_bound",
                     "b200alac_last_error", "b200alac_free"}
    L = b200alac.lib()
    for n in names:
        assert hasattr(L, n), n
    if b200flac.device_count() == 0:
        with pytest.raises(b200alac.B200AlacError, match="no CPU fallback"):
            b200alac.encode_framesets(b"\0" * 400, 100, b200alac.make_params())
