"""CPU: the TTA oracle (oracle/tta_oracle.c) against the golden manifest made from the COMPILED REFERENCE encoder
(tests/golden/tta_golden.json, tests/golden/make_tta_golden.py) and, where oracle/_ref exists, against the
reference binary itself -- whole files, byte for byte (reference src/encoders/tta.c)."""
import hashlib
import json
import os
import struct
import zlib

import numpy as np
import pytest

import helpers
from golden.tta_cases import TTA_CASES, tta_case_pcm

with open(os.path.join(helpers.GOLDEN, "tta_golden.json")) as _fh:
    GOLD = {c["name"]: c for c in json.load(_fh)["cases"]}


@pytest.mark.parametrize("case", TTA_CASES, ids=[c["name"] for c in TTA_CASES])
def test_oracle_matches_reference_golden(case):
    g = GOLD[case["name"]]
    pcm = tta_case_pcm(case)
    assert hashlib.sha256(pcm).hexdigest() == g["pcm_sha256"], "input generator drifted"
    data = helpers.oracle_tta_file(pcm, case["rate"], case["channels"], case["bps"])
    assert len(data) == g["length"] and hashlib.sha256(data).hexdigest() == g["sha256"]


@pytest.mark.skipif(not helpers.have_tta_ref(), reason="oracle/_ref/ttaenc not built (needs /root/reference)")
@pytest.mark.parametrize("rate,ch,bps,n", [(44100, 2, 16, 46080 * 2 + 17), (48000, 1, 24, 12345), (32000, 4, 8, 40000),
                                            (96000, 2, 24, 100310), (44100, 5, 16, 46081)])
def test_oracle_matches_compiled_reference(rate, ch, bps, n):
    pcm = helpers.synth_pcm(900 + ch, ch, bps, n)
    ours = helpers.oracle_tta_file(pcm, rate, ch, bps)
    ref = helpers.ref_tta_encode(pcm, rate, ch, bps)
    assert ours == ref
    assert helpers.ref_tta_decode(ours) == pcm


def test_file_layout_and_crcs():
    """write_header / write_seektable / encode_frame (tta.c:562-595,252-255): every CRC-32 in the file is the
    standard reflected CRC-32 (zlib's) of the bytes it covers, sizes in the seektable add up to the file"""
    rate, ch, bps, n = 44100, 2, 16, 100000
    pcm = helpers.synth_pcm(1, ch, bps, n)
    data = helpers.oracle_tta_file(pcm, rate, ch, bps)
    assert data[:4] == b"TTA1"
    fmt, c, b, r, total, crc = struct.unpack("<HHHIII", data[4:22])
    assert (fmt, c, b, r, total) == (1, ch, bps, rate, n) and crc == zlib.crc32(data[:18])
    block = (rate * 256) // 245
    nf = (n + block - 1) // block
    sizes = struct.unpack("<%dI" % nf, data[22:22 + 4 * nf])
    assert struct.unpack("<I", data[22 + 4 * nf:26 + 4 * nf])[0] == zlib.crc32(data[22:22 + 4 * nf])
    pos = 26 + 4 * nf
    for s in sizes:
        frame = data[pos:pos + s]
        assert struct.unpack("<I", frame[-4:])[0] == zlib.crc32(frame[:-4])
        pos += s
    assert pos == len(data)
    frames, sz = helpers.oracle_tta_frames(pcm, rate, ch, bps)
    assert frames == data[26 + 4 * nf:] and tuple(sz) == sizes


def test_short_reads_become_short_frames():
    """the reference encodes whatever length pcmreader->read() returns as one frame (tta.c:69-83)"""
    rate, ch, bps = 44100, 2, 16
    lens = [46080, 100, 46080, 7, 3000]
    pcm = helpers.synth_pcm(2, ch, bps, sum(lens))
    frames, sizes = helpers.oracle_tta_frames(pcm, rate, ch, bps, lens)
    assert len(sizes) == len(lens) and sum(sizes) == len(frames)
    pos, off = 0, 0
    for n, s in zip(lens, sizes):
        one, one_s = helpers.oracle_tta_frames(pcm[pos * 4:(pos + n) * 4], rate, ch, bps, [n])
        assert frames[off:off + s] == one and one_s == [s]
        pos += n
        off += s


def test_tta_abi_exports_every_declared_symbol(built):
    """the library exports what include/b200tta.h declares; without a GPU every entry point fails loudly"""
    import re
    import b200flac
    import b200tta
    hdr = open(os.path.join(helpers.ROOT, "include", "b200tta.h")).read()
    names = set(re.findall(r"\b(b200tta_[a-z_0-9]+)\s*\(", hdr))
    assert names >= {"b200tta_encode_frames", "b200tta_encode_device", "b200tta_encode_file", "b200tta_block_size",
                     "b200tta_output_bound", "b200tta_last_error", "b200tta_free"}
    L = b200tta.lib()
    for n in names:
        assert hasattr(L, n), n
    assert b200tta.block_size(44100) == 46080 and b200tta.block_size(96000) == 100310
    if b200flac.device_count() == 0:
        with pytest.raises(b200tta.B200TtaError, match="no CPU fallback"):
            b200tta.encode_frames(b"\0" * 400, 100, 44100, 2, 16)
