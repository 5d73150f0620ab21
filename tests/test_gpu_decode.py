"""GPU FLAC decode / verify (reference src/decoders/flac.c:174-286, 569-1270, 1340-1510).

Streams come from the CPU oracle encoder (byte-identical to the reference encoder's) at every
compression level and shape; the GPU decoder must return the PCM that went in (and what the compiled
reference decoder returns), the same frame list, and the reference's error for damaged streams.
A hand-assembled stream covers what the reference encoder never writes: escape-coded partitions,
Rice parameter 15 with escape 0, wasted bits on a side channel, an 8-bit block-size escape."""
import hashlib
import os
import struct

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

LEVELS = {
    "0": dict(block_size=1152, max_lpc_order=0, max_residual_partition_order=3),
    "3": dict(block_size=1152, max_lpc_order=6, max_residual_partition_order=3, adaptive_mid_side=True),
    "5": dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=5, mid_side=True),
    "8": dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True, exhaustive_model_search=True),
    "l32": dict(block_size=4096, max_lpc_order=32, max_residual_partition_order=8, mid_side=True),
}


def _decode(flac, **kw):
    import b200flac
    return b200flac.decode(flac, **kw)


@pytest.mark.parametrize("level", sorted(LEVELS))
@pytest.mark.parametrize("rate,ch,bps,n", [(44100, 2, 16, 44100 * 3 + 17), (96000, 2, 24, 50000), (8000, 1, 8, 9000),
                                           (96000, 6, 24, 4608 * 5 + 100)])
def test_decode_round_trip(level, rate, ch, bps, n, built):
    pcm = helpers.synth_pcm(11 + ch + bps, ch, bps, n)
    opts = helpers.options(**LEVELS[level])
    flac, offs = helpers.oracle_encode(pcm, rate, ch, bps, opts, want_offsets=True)
    info, got, frames, ms = _decode(flac, want_frames=True)
    assert (info.sample_rate, info.channels, info.bits_per_sample, info.total_pcm_frames) == (rate, ch, bps, n)
    assert bytes(info.md5) == hashlib.md5(pcm).digest()
    assert got == pcm
    assert frames == offs                      # same (byte offset, PCM frames) list as the encoder reported
    if helpers.have_ref() and level in ("5", "l32"):
        assert helpers.ref_decode(flac) == got


@pytest.mark.parametrize("name", ["silence", "constant", "noise", "wasted", "full_scale", "tiny", "one_sample"])
def test_decode_special_blocks(name, built):
    rng = np.random.default_rng(5)
    if name == "silence":
        pcm, ch, bps = bytes(4 * 10000), 2, 16
    elif name == "constant":
        pcm, ch, bps = helpers.pack_pcm(np.full((9000, 2), -1234, dtype=np.int32), 16), 2, 16
    elif name == "noise":       # VERBATIM subframes
        pcm, ch, bps = helpers.pack_pcm(rng.integers(-32768, 32768, size=(12000, 2)).astype(np.int32), 16), 2, 16
    elif name == "wasted":
        pcm, ch, bps = helpers.wasted_bps16(20000), 2, 16
    elif name == "full_scale":
        pat = helpers.full_scale_patterns(24)[4]
        pcm, ch, bps = helpers.pack_pcm(np.array((pat * 3000)[:16000], dtype=np.int32).reshape(-1, 2), 24), 2, 24
    elif name == "tiny":
        pcm, ch, bps = helpers.synth_pcm(3, 2, 16, 10), 2, 16
    else:
        pcm, ch, bps = helpers.synth_pcm(3, 1, 16, 1), 1, 16
    flac = helpers.oracle_encode(pcm, 44100, ch, bps, helpers.options(**LEVELS["8"]))
    info, got = _decode(flac)
    assert got == pcm
    if helpers.have_ref():
        assert helpers.ref_decode(flac) == got


def test_decode_empty_stream(built):
    flac = helpers.oracle_encode(b"", 44100, 2, 16, helpers.options())
    info, got = _decode(flac)
    assert info.total_pcm_frames == 0 and got == b""


def test_decode_short_frames_mid_stream(tmp_path, built):
    """frames of unequal length in one stream (the stream layer's end_block)"""
    import b200flac
    p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=8, max_residual_partition_order=5)
    path = os.path.join(str(tmp_path), "s.flac")
    pcm = helpers.synth_pcm(9, 2, 16, 30000)
    s = b200flac.Stream(path, p)
    for a, b in ((0, 5000), (5000, 5001), (5001, 20000), (20000, 30000)):
        s.write(pcm[a * 4:b * 4])
        s.end_block()
    offs = s.close()
    flac = open(path, "rb").read()
    info, got, frames, ms = _decode(flac, want_frames=True)
    assert got == pcm and frames == offs
    assert len({f[1] for f in frames}) > 2


def _flip(data, pos, mask=0x10):
    b = bytearray(data)
    b[pos] ^= mask
    return bytes(b)


def _ref_error(flac):
    try:
        helpers.ref_decode(flac)
    except RuntimeError as e:
        return str(e)
    return None


def test_decode_damaged_streams_fail_like_the_reference(built):
    pcm = helpers.synth_pcm(21, 2, 16, 4096 * 6 + 50)
    flac, offs = helpers.oracle_encode(pcm, 44100, 2, 16, helpers.options(**LEVELS["5"]), want_offsets=True)
    first = helpers.first_frame_offset(flac)
    cases = [
        (_flip(flac, first + offs[2][0] + 40), ValueError, "invalid checksum in frame"),            # payload bit
        (_flip(flac, first + offs[3][0] + 2), ValueError, "invalid checksum in frame header"),      # header field
        (_flip(flac, first + offs[1][0] + 1, 0x02), ValueError, "invalid reserved bit"),
        (_flip(flac, first + offs[1][0], 0x80), ValueError, "invalid sync code"),
        (_flip(flac, 8 + 18 + 3), ValueError, "MD5 mismatch at end of stream"),                      # STREAMINFO MD5
        (flac[:first + offs[5][0] + 3], IOError, "EOF reading frame"),
        (b"fLaX" + flac[4:], ValueError, "not a FLAC file"),
    ]
    for bad, exc, text in cases:
        with pytest.raises(exc, match=text):
            _decode(bad)
        if helpers.have_ref() and text not in ("not a FLAC file",):
            err = _ref_error(bad)
            assert err is not None
            want = {"EOF reading frame": "I/O Error reading frame"}.get(text, text)
            assert want in err, (text, err)


class Bits(object):
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, value, bits):
        self.v = (self.v << bits) | (value & ((1 << bits) - 1))
        self.n += bits

    def unary(self, zeros):
        self.put(1, zeros + 1)

    def align(self):
        if self.n % 8:
            self.put(0, 8 - self.n % 8)

    def bytes(self):
        return self.v.to_bytes(self.n // 8, "big")


def _crc(data, poly, width):
    crc, top, mask = 0, 1 << (width - 1), (1 << width) - 1
    for byte in data:
        crc ^= byte << (width - 8)
        for _ in range(8):
            crc = ((crc << 1) ^ poly) & mask if crc & top else (crc << 1) & mask
    return crc


def test_decode_hand_made_frame(built):
    """one left/side frame of 16 samples: FIXED order 1 with an escape-coded partition and a Rice partition
    with parameter 15 whose escape field is 0 (the reference keeps Rice coding, flac.c:1180-1186,1195), then a
    side channel with 2 wasted bits as FIXED order 0 at coding method 1; block size through the 8-bit escape"""
    n = 16
    left = [100, 103, 99, 120, -70, -71, -69, 0, 5, 6, 7, 8, 9, 10, 11, 12]
    side = [4 * v for v in (1, -2, 3, -4, 5, -6, 7, -8, 0, 0, 1, 1, -1, -1, 2, -2)]
    b = Bits()
    b.put(0x3FFE, 14); b.put(0, 1); b.put(0, 1)
    b.put(6, 4); b.put(9, 4); b.put(8, 4); b.put(4, 3); b.put(0, 1)      # 8-bit block size, 44.1 kHz, left/side, 16 bit
    b.put(0, 8)                                                           # frame number 0
    b.put(n - 1, 8)
    b.put(_crc(b.bytes(), 0x07, 8), 8)
    # subframe 0: FIXED order 1, 16 bits, partition order 1: [7 residuals escape-coded at 9 bits][8 residuals Rice 15 / escape 0]
    b.put(0, 1); b.put(0b001001, 6); b.put(0, 1)
    b.put(left[0], 16)
    res = [left[i] - left[i - 1] for i in range(1, n)]
    b.put(0, 2); b.put(1, 4)
    b.put(15, 4); b.put(9, 5)
    for r in res[:7]:
        b.put(r, 9)
    b.put(15, 4); b.put(0, 5)
    for r in res[7:]:
        u = 2 * r if r >= 0 else -2 * r - 1
        b.unary(u >> 15); b.put(u, 15)
    # subframe 1: side at 17 bits, 2 wasted bits, FIXED order 0, coding method 1, partition order 0, Rice 3
    b.put(0, 1); b.put(0b001000, 6); b.put(1, 1); b.unary(1)
    b.put(1, 2); b.put(0, 4); b.put(3, 5)
    for v in side:
        r = v >> 2
        u = 2 * r if r >= 0 else -2 * r - 1
        b.unary(u >> 3); b.put(u, 3)
    b.align()
    frame = b.bytes()
    frame += struct.pack(">H", _crc(frame, 0x8005, 16))
    right = [l - s for l, s in zip(left, side)]
    pcm = helpers.pack_pcm(np.array(list(zip(left, right)), dtype=np.int32), 16)
    si = struct.pack(">HH", 16, 16) + b"\x00" * 6 + ((44100 << 44) | (1 << 41) | (15 << 36) | n).to_bytes(8, "big") + hashlib.md5(pcm).digest()
    flac = b"fLaC" + bytes([0x80, 0, 0, 34]) + si + frame
    if helpers.have_ref():
        assert helpers.ref_decode(flac) == pcm
    info, got = _decode(flac)
    assert got == pcm


def test_verify_file(tmp_path, built):
    import b200flac
    pcm = helpers.synth_pcm(2, 2, 16, 100000)
    flac = helpers.oracle_encode(pcm, 44100, 2, 16, helpers.options(**LEVELS["8"]))
    good, bad = os.path.join(str(tmp_path), "g.flac"), os.path.join(str(tmp_path), "b.flac")
    open(good, "wb").write(flac)
    open(bad, "wb").write(_flip(flac, len(flac) - 100))
    assert b200flac.verify_file(good)
    with pytest.raises(ValueError, match="invalid checksum in frame"):
        b200flac.verify_file(bad)


def test_flacdecoder_reads_frame_by_frame(tmp_path, built):
    """decoders.FlacDecoder (src/decoders/flac.c:28-286): PCMReader attributes, one FLAC frame per read(),
    an empty FrameList at the end, ValueError after close()"""
    import audiotools
    import audiotools.decoders
    n = 4096 * 7 + 99
    pcm = helpers.synth_pcm(8, 2, 16, n)
    path = os.path.join(str(tmp_path), "d.flac")
    flac = audiotools.FlacAudio.from_pcm(path, audiotools.PCMBytesReader(pcm, 44100, 2, 0x3, 16), "6")
    d = flac.to_pcm()
    assert (d.sample_rate, d.channels, d.bits_per_sample, d.channel_mask) == (44100, 2, 16, 0x3)
    sizes, got = [], []
    while True:
        f = d.read(4096)
        if f.frames == 0:
            break
        sizes.append(f.frames)
        got.extend(list(f))
    assert sizes == [4096] * 7 + [99] and got == list(helpers.unpack_pcm(pcm, 16))
    assert d.read(4096).frames == 0
    assert [o[1] for o in d.offsets()] == sizes
    d.close()
    with pytest.raises(ValueError, match="cannot read closed stream"):
        d.read(4096)
    assert flac.verify() is True


def test_flacdecoder_seek_follows_the_seektable(tmp_path, built):
    """FlacDecoder.seek (src/decoders/flac.c:288-356): the latest SEEKTABLE point at or before the requested PCM
    frame -- from_pcm writes one every 10 s --, its PCM frame number returned, reading resumes at that frame, the
    MD5 is only validated when decoding restarts from 0; without a table every seek lands on frame 0"""
    import audiotools
    rate, n = 8000, 8000 * 35 + 123
    pcm = helpers.synth_pcm(12, 1, 16, n)
    path = os.path.join(str(tmp_path), "s.flac")
    flac = audiotools.FlacAudio.from_pcm(path, audiotools.PCMBytesReader(pcm, rate, 1, 0x4, 16), "5")
    samples = helpers.unpack_pcm(pcm, 16)
    d = flac.to_pcm()
    # seek points at 0, 80000, 160000, 240000 PCM frames -> the 4096-frame blocks containing them
    starts = [(t // 4096) * 4096 for t in (0, 80000, 160000, 240000)]
    assert starts == [0, 77824, 159744, 237568]
    # (a point's sample number is its frame's first sample, so 159744..159999 already belong to the third point)
    for want_frame, target in ((starts[0], 0), (starts[0], 4096 * 3), (starts[1], 80000), (starts[1], 159743),
                               (starts[2], 159744), (starts[2], 160000), (starts[3], n + 5000)):
        assert d.seek(target) == want_frame
        f = d.read(4096)
        assert list(f) == list(samples[want_frame:want_frame + f.frames]) and f.frames == min(4096, n - want_frame)
    # read to the end after a seek into the middle: no MD5 verdict (validation is off), an empty FrameList ends it
    d.seek(200000)
    total = 0
    while True:
        f = d.read(4096)
        if f.frames == 0:
            break
        total += f.frames
    assert total == n - starts[2]
    # back to the start: the whole stream again, MD5 checked
    assert d.seek(0) == 0
    total = 0
    while True:
        f = d.read(4096)
        if f.frames == 0:
            break
        total += f.frames
    assert total == n
    with pytest.raises(ValueError, match="negative"):
        d.seek(-1)
    d.close()
    with pytest.raises(ValueError, match="cannot seek closed stream"):
        d.seek(0)
    # a stream without a SEEKTABLE (the encoder's own output): every seek goes to the start
    raw = os.path.join(str(tmp_path), "raw.flac")
    open(raw, "wb").write(helpers.oracle_encode(pcm, rate, 1, 16, helpers.options()))
    d = audiotools.FlacAudio(raw).to_pcm()
    assert d.seek(100000) == 0 and list(d.read(4096)) == list(samples[:4096])


def test_flacdecoder_md5_mismatch_at_end_of_stream(tmp_path, built):
    import audiotools
    import audiotools.decoders
    pcm = helpers.synth_pcm(8, 2, 16, 10000)
    flac = helpers.oracle_encode(pcm, 44100, 2, 16, helpers.options())
    path = os.path.join(str(tmp_path), "m.flac")
    open(path, "wb").write(_flip(flac, 8 + 18 + 5))
    d = audiotools.FlacAudio(path).to_pcm()
    frames = 0
    with pytest.raises(ValueError, match="MD5 mismatch at end of stream"):
        while True:
            f = d.read(4096)
            assert f.frames > 0        # every frame is delivered before the mismatch is reported
            frames += f.frames
    assert frames == 10000
    with pytest.raises(audiotools.InvalidFLAC, match="MD5 mismatch at end of stream"):
        audiotools.FlacAudio(path).verify()


def test_transcode_round_trip_through_both_engines(tmp_path, built):
    """WAVE -> FLAC (file feed) -> PCM (GPU decoder) == the WAVE's data chunk"""
    import audiotools
    import struct as st
    n = 4096 * 20 + 7
    pcm = helpers.synth_pcm(77, 2, 24, n)
    wav = os.path.join(str(tmp_path), "t.wav")
    fmt = st.pack("<HHIIHH", 1, 2, 96000, 96000 * 6, 6, 24)
    open(wav, "wb").write(b"RIFF" + st.pack("<I", 4 + 8 + len(fmt) + 8 + len(pcm)) + b"WAVE" + b"fmt " + st.pack("<I", len(fmt)) + fmt +
                          b"data" + st.pack("<I", len(pcm)) + pcm)
    path = os.path.join(str(tmp_path), "t.flac")
    flac = audiotools.FlacAudio.from_pcm(path, audiotools.WaveAudio(wav).to_pcm(), "8")
    import b200flac
    info, got = b200flac.decode(open(path, "rb").read())
    assert got == pcm and info.channel_mask == 0x3


def test_standalone_decoder_driver_matches_reference_driver(tmp_path, built):
    """`b200flacdec file.flac > pcm` against `flacdec` (the compiled reference's driver,
    src/decoders/flac.c:1340-1527): same PCM and exit status for good streams, same stderr line and exit
    status for damaged ones (the reference additionally writes the frames before the damage)"""
    import subprocess
    exe = os.path.join(helpers.ROOT, "python-audio-tools_b200", "b200flacdec")
    ref = os.path.join(helpers.ROOT, "oracle", "_ref", "flacdec")
    pcm = helpers.synth_pcm(31, 2, 16, 4096 * 5 + 9)
    flac, offs = helpers.oracle_encode(pcm, 44100, 2, 16, helpers.options(**LEVELS["8"]), want_offsets=True)
    first = helpers.first_frame_offset(flac)
    variants = {"good": flac, "crc16": _flip(flac, first + offs[2][0] + 50), "crc8": _flip(flac, first + offs[2][0] + 3),
                "md5": _flip(flac, 8 + 18 + 9), "cut": flac[:first + offs[4][0] + 10], "notflac": b"OggS" + flac[4:]}
    for name, data in variants.items():
        path = os.path.join(str(tmp_path), name + ".flac")
        open(path, "wb").write(data)
        r = subprocess.run([exe, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if name == "good":
            assert r.returncode == 0 and r.stdout == pcm and r.stderr == b""
        else:
            assert r.returncode == 1
        if name == "md5":
            assert r.stdout == pcm          # reported after the PCM, as in the reference
        if helpers.have_ref():
            q = subprocess.run([ref, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            assert (q.returncode, q.stderr) == (r.returncode, r.stderr), name
            if name in ("good", "md5"):
                assert q.stdout == r.stdout
    r = subprocess.run([exe, os.path.join(str(tmp_path), "missing.flac")], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1 and r.stderr.startswith(b"*** ") and b"No such file" in r.stderr


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_decode_fuzz_same_verdict_and_message_as_the_reference(seed, tmp_path, built):
    """random damage (bit flips, byte overwrites, truncation, zeroed runs, duplicated runs) anywhere after the
    metadata: the GPU decoder and the compiled reference decoder agree on accept/reject, on the PCM when
    they accept, and on the error text when they reject -- which also drives garbage through every
    speculative path of the frame kernel.  DECODE_FUZZ_CASES sets the number of cases per seed (default 150;
    2400 cases over six seeds were run when this went in)."""
    import random
    import subprocess
    if not helpers.have_ref():
        pytest.skip("needs oracle/_ref/flacdec")
    ref = os.path.join(helpers.ROOT, "oracle", "_ref", "flacdec")
    rng = random.Random(seed)
    pcm = helpers.synth_pcm(5, 2, 16, 4096 * 6 + 100)
    flac = helpers.oracle_encode(pcm, 44100, 2, 16, helpers.options(**LEVELS["5"]))
    first = helpers.first_frame_offset(flac)
    path = os.path.join(str(tmp_path), "f.flac")
    for case in range(int(os.environ.get("DECODE_FUZZ_CASES", "150"))):
        b = bytearray(flac)
        kind = rng.choice(["flip", "flip", "flip", "byte", "cut", "zero", "dup"])
        if kind == "flip":
            for _ in range(rng.choice([1, 1, 2, 5])):
                b[rng.randrange(first, len(b))] ^= 1 << rng.randrange(8)
        elif kind == "byte":
            b[rng.randrange(first, len(b))] = rng.randrange(256)
        elif kind == "cut":
            b = b[:rng.randrange(first, len(b))]
        elif kind == "zero":
            p, n = rng.randrange(first, len(b)), rng.randrange(1, 200)
            b[p:p + n] = bytes(min(n, len(b) - p))
        else:
            p, n = rng.randrange(first, len(b)), rng.randrange(1, 50)
            b[p:p] = b[p:p + n]
        data = bytes(b)
        open(path, "wb").write(data)
        r = subprocess.run([ref, path], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        try:
            info, got = _decode(data)
            mine = None
        except (ValueError, IOError) as e:
            got, mine = None, str(e)
        if r.returncode == 0:
            assert mine is None and got == r.stdout, (seed, case, kind, mine)
        else:
            want = r.stderr.decode().strip().replace("*** Error: ", "").replace("*** ", "")
            want = {"I/O Error reading frame": "EOF reading frame"}.get(want, want)
            assert mine == want, (seed, case, kind)


def _decode_golden():
    import json
    with open(os.path.join(helpers.ROOT, "tests", "golden", "decode_golden.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", sorted(_decode_golden()))
def test_decode_reference_fixtures(name, built):
    """the reference's own FLAC fixtures (test/*.flac: written by other encoders -- libFLAC tones, every
    subframe type, metadata in unusual order, ID3 prefix, blank MD5, 32768-sample blocks) through the
    standalone driver: exit status, stderr and PCM equal what the compiled reference decoder produced
    (tests/golden/make_decode_golden.py)"""
    import subprocess
    want = _decode_golden()[name]
    exe = os.path.join(helpers.ROOT, "python-audio-tools_b200", "b200flacdec")
    r = subprocess.run([exe, os.path.join(helpers.ROOT, "tests", "golden", "flac", name)], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE)
    assert (r.returncode, r.stderr.decode()) == (want["rc"], want["stderr"])
    assert len(r.stdout) == want["pcm_bytes"] and hashlib.sha256(r.stdout).hexdigest() == want["pcm_sha256"]


@pytest.mark.parametrize("name", ["flac-id3.flac", "flac-id3-2.flac", "flac-disordered.flac"])
def test_flacaudio_reads_id3_prefixed_and_disordered_fixtures(name, built):
    """FlacAudio skips ID3v2 tags in front of the stream (flac.py:2420-2462, id3.py:264-311) and finds a
    STREAMINFO that is not the first block; to_pcm() then decodes from the stream offset, and the decoded
    PCM hashes to the STREAMINFO MD5"""
    import audiotools
    a = audiotools.FlacAudio(os.path.join(helpers.ROOT, "tests", "golden", "flac", name))
    d = a.to_pcm()
    assert (d.sample_rate, d.channels, d.bits_per_sample) == (a.sample_rate(), a.channels(), a.bits_per_sample())
    h, frames = hashlib.md5(), 0
    while True:
        f = d.read(4096)
        if f.frames == 0:
            break
        frames += f.frames
        h.update(f.to_bytes(False, True))
    assert frames == a.total_frames() and h.digest() == a.__md5__ and a.__md5__ != bytes(16)
    d.close()


def test_device_chain_and_host_walk_agree(built, monkeypatch):
    """the frame chain is normally resolved on the device (hash table + pointer jumping); the host walk that
    words the errors of damaged streams must give the same frames and PCM on a good one"""
    pcm = helpers.synth_pcm(61, 2, 16, 4096 * 40 + 11)
    flac, offs = helpers.oracle_encode(pcm, 44100, 2, 16, helpers.options(**LEVELS["5"]), want_offsets=True)
    info, got, frames, ms = _decode(flac, want_frames=True)
    monkeypatch.setenv("B200FLAC_DEC_HOST_CHAIN", "1")
    info2, got2, frames2, ms2 = _decode(flac, want_frames=True)
    assert got == got2 == pcm and frames == frames2 == offs
