"""File-backed PCM sources: audiotools.wav.WaveReader / audiotools.aiff.AiffReader
(reference audiotools/wav.py:421-553, aiff.py:350-482), their C restatements b200flac_wave_probe /
b200flac_aiff_probe, and the direct file feed b200flac_stream_write_file.

CPU: the Python readers and the C probes agree field for field and error for error on well-formed and
malformed containers (the reference's checks, including the `fmt ` remainder it leaves unread).
GPU: FlacAudio.from_pcm over a WAVE/AIFF reader writes the same file with the file feed as with one
FrameList per read(), and its frames equal the oracle's, byte for byte."""
import hashlib
import os
import struct

import pytest

import helpers

PCM_GUID = b"\x01\x00\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"


def _at():
    import audiotools
    return audiotools


def chunk(cid, body, big=False, pad=True):
    out = cid + struct.pack(">I" if big else "<I", len(body)) + body
    if pad and len(body) % 2:
        out += b"\x00"
    return out


def fmt_plain(channels, rate, bps, compression=1):
    return struct.pack("<HHIIHH", compression, channels, rate, rate * channels * bps // 8, channels * bps // 8, bps)


def fmt_extensible(channels, rate, bps, mask, guid=PCM_GUID):
    return fmt_plain(channels, rate, bps, 0xFFFE) + struct.pack("<HHI", 22, bps, mask) + guid


def wave(chunks, size=None):
    body = b"WAVE" + b"".join(chunks)
    return b"RIFF" + struct.pack("<I", len(body) if size is None else size) + body


def comm(channels, frames, bps, rate_bytes=b"\x40\x0e\xac\x44\x00\x00\x00\x00\x00\x00"):  # 44100 Hz
    return struct.pack(">HIH", channels, frames, bps) + rate_bytes


def aiff(chunks, size=None):
    body = b"AIFF" + b"".join(chunks)
    return b"FORM" + struct.pack(">I", len(body) if size is None else size) + body


def le_to_wave_data(pcm, bps):
    """signed little-endian PCM -> the bytes a WAVE data chunk holds (8-bit is unsigned)"""
    return bytes((b ^ 0x80) for b in pcm) if bps == 8 else pcm


def le_to_aiff_data(pcm, bps):
    n = bps // 8
    return b"".join(pcm[i:i + n][::-1] for i in range(0, len(pcm), n))


PCM16 = helpers.synth_pcm(7, 2, 16, 300)
WAVE_CASES = {
    "plain16": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16)), chunk(b"data", PCM16)]),
    "mono8": wave([chunk(b"fmt ", fmt_plain(1, 8000, 8)), chunk(b"data", bytes(range(200)))]),
    "six24_plain": wave([chunk(b"fmt ", fmt_plain(6, 96000, 24)), chunk(b"data", bytes(18 * 10))]),
    "seven_plain": wave([chunk(b"fmt ", fmt_plain(7, 48000, 16)), chunk(b"data", bytes(14 * 4))]),
    "extensible": wave([chunk(b"fmt ", fmt_extensible(6, 96000, 24, 0x60F)), chunk(b"data", bytes(18 * 10))]),
    "odd_list_before_data": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16)), chunk(b"LIST", b"abc"), chunk(b"data", PCM16)]),
    "partial_last_frame": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16)), chunk(b"data", PCM16 + b"\x01\x02\x03")]),
    "two_fmt_last_wins": wave([chunk(b"fmt ", fmt_plain(1, 8000, 8)), chunk(b"fmt ", fmt_plain(2, 44100, 16)), chunk(b"data", PCM16)]),
    "data_claims_more_than_file": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16)), b"data" + struct.pack("<I", 4000) + PCM16[:400]]),
    # malformed
    "fmt18_remainder_unread": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16) + b"\x00\x00"), chunk(b"data", PCM16)]),
    "short_header": b"RIFF\x10\x00",
    "not_riff": b"RIFX" + wave([chunk(b"fmt ", fmt_plain(2, 44100, 16)), chunk(b"data", PCM16)])[4:],
    "not_wave": b"RIFF\x20\x00\x00\x00WAVX" + bytes(28),
    "bad_chunk_id": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16)), chunk(b"da\x01a", PCM16)]),
    "data_before_fmt": wave([chunk(b"data", PCM16), chunk(b"fmt ", fmt_plain(2, 44100, 16))]),
    "no_data": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16)), chunk(b"LIST", b"abcd")]),
    "chunk_header_cut": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16))], size=4 + 24 + 20) + b"dat",
    "odd_chunk_without_pad": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16))], size=200) + chunk(b"LIST", b"abc", pad=False),
    "adpcm": wave([chunk(b"fmt ", fmt_plain(2, 44100, 16, compression=2)), chunk(b"data", PCM16)]),
    "bad_sub_format": wave([chunk(b"fmt ", fmt_extensible(2, 44100, 16, 3, guid=bytes(16))), chunk(b"data", PCM16)]),
    "fmt_cut": b"RIFF\x40\x00\x00\x00WAVEfmt \x10\x00\x00\x00\x01\x00\x02",
    "tiny_size": b"RIFF\x02\x00\x00\x00WAVE" + chunk(b"fmt ", fmt_plain(2, 44100, 16)) + chunk(b"data", PCM16),
}
AIFF_CASES = {
    "plain16": aiff([chunk(b"COMM", comm(2, 300, 16), True), chunk(b"SSND", bytes(8) + le_to_aiff_data(PCM16, 16), True)]),
    "mono24_96k": aiff([chunk(b"COMM", comm(1, 10, 24, b"\x40\x0f\xbb\x80\x00\x00\x00\x00\x00\x00"), True),
                        chunk(b"SSND", bytes(8 + 30), True)]),
    "six": aiff([chunk(b"COMM", comm(6, 4, 16), True), chunk(b"SSND", bytes(8 + 48), True)]),
    "odd_chunk_first": aiff([chunk(b"NAME", b"abc", True), chunk(b"COMM", comm(2, 300, 16), True),
                             chunk(b"SSND", bytes(8) + le_to_aiff_data(PCM16, 16), True)]),
    "zero_rate": aiff([chunk(b"COMM", comm(2, 1, 16, bytes(10)), True), chunk(b"SSND", bytes(12), True)]),
    # malformed
    "short_header": b"FORM\x00",
    "not_form": b"FORX" + bytes(20),
    "not_aiff": b"FORM\x00\x00\x00\x20AIFC" + bytes(28),
    "bad_chunk_id": aiff([chunk(b"CO\x7fM", comm(2, 300, 16), True)]),
    "ssnd_before_comm": aiff([chunk(b"SSND", bytes(16), True), chunk(b"COMM", comm(2, 2, 16), True)]),
    "no_ssnd": aiff([chunk(b"COMM", comm(2, 300, 16), True)]),
    "zero_channels": aiff([chunk(b"COMM", comm(0, 300, 16), True), chunk(b"SSND", bytes(16), True)]),
    "comm_cut": b"FORM\x00\x00\x00\x40AIFFCOMM\x00\x00\x00\x12\x00\x02",
    "odd_chunk_without_pad": aiff([chunk(b"COMM", comm(2, 300, 16), True)], size=200) + chunk(b"NAME", b"abc", True, pad=False),
}


def _python_side(reader_cls, path):
    try:
        r = reader_cls(path)
    except ValueError as e:
        return ("ValueError", str(e))
    except IOError as e:
        return ("IOError", str(e))
    except ZeroDivisionError:
        return ("ValueError", "integer division or modulo by zero")
    offset = r.file.tell()
    out = (r.sample_rate, r.channels, r.bits_per_sample, r.channel_mask, offset, r.total_pcm_frames)
    r.close()
    return out


def _c_side(probe, path):
    try:
        s = probe(path)
    except ValueError as e:
        return ("ValueError", str(e))
    except IOError as e:
        return ("IOError", str(e))
    return (s.sample_rate, s.channels, s.bits_per_sample, s.channel_mask, s.data_offset, s.total_pcm_frames)


@pytest.mark.parametrize("name", sorted(WAVE_CASES))
def test_wave_reader_and_c_probe_agree(name, tmp_path, built):
    import b200flac
    path = os.path.join(str(tmp_path), name + ".wav")
    open(path, "wb").write(WAVE_CASES[name])
    py = _python_side(_at().wav.WaveReader, path)
    assert _c_side(b200flac.wave_probe, path) == py
    expect = {"plain16": (44100, 2, 16, 0x3, 44, 300), "mono8": (8000, 1, 8, 0x4, 44, 200),
              "six24_plain": (96000, 6, 24, 0x3F, 44, 10), "seven_plain": (48000, 7, 16, 0, 44, 4),
              "extensible": (96000, 6, 24, 0x60F, 68, 10), "odd_list_before_data": (44100, 2, 16, 0x3, 56, 300),
              "partial_last_frame": (44100, 2, 16, 0x3, 44, 300), "two_fmt_last_wins": (44100, 2, 16, 0x3, 68, 300),
              "data_claims_more_than_file": (44100, 2, 16, 0x3, 44, 1000),
              "fmt18_remainder_unread": ("ValueError", "invalid RIFF WAVE chunk ID"),
              "short_header": ("ValueError", "invalid RIFF WAVE file"), "not_riff": ("ValueError", "not a RIFF WAVE file"),
              "not_wave": ("ValueError", "invalid RIFF WAVE file"), "bad_chunk_id": ("ValueError", "invalid RIFF WAVE chunk ID"),
              "data_before_fmt": ("ValueError", "data chunk found before fmt"), "no_data": ("ValueError", "data chunk not found"),
              "chunk_header_cut": ("ValueError", "invalid RIFF WAVE file"),
              "odd_chunk_without_pad": ("ValueError", "invalid RIFF WAVE chunk ID"),
              "adpcm": ("ValueError", "unsupported WAVE compression"), "bad_sub_format": ("ValueError", "invalid WAVE sub-format"),
              "fmt_cut": ("IOError", "I/O error reading stream"), "tiny_size": ("ValueError", "data chunk not found")}[name]
    assert py == expect


@pytest.mark.parametrize("name", sorted(AIFF_CASES))
def test_aiff_reader_and_c_probe_agree(name, tmp_path, built):
    import b200flac
    path = os.path.join(str(tmp_path), name + ".aiff")
    open(path, "wb").write(AIFF_CASES[name])
    py = _python_side(_at().aiff.AiffReader, path)
    assert _c_side(b200flac.aiff_probe, path) == py
    expect = {"plain16": (44100, 2, 16, 0x3, 54, 300), "mono24_96k": (96000, 1, 24, 0x4, 54, 10),
              "six": (44100, 6, 16, 0, 54, 4), "odd_chunk_first": (44100, 2, 16, 0x3, 66, 300),
              "zero_rate": (0, 2, 16, 0x3, 54, 1),
              "short_header": ("ValueError", "invalid AIFF file"), "not_form": ("ValueError", "not an AIFF file"),
              "not_aiff": ("ValueError", "invalid AIFF file"), "bad_chunk_id": ("ValueError", "invalid AIFF chunk ID"),
              "ssnd_before_comm": ("ValueError", "SSND chunk found before fmt"), "no_ssnd": ("ValueError", "SSND chunk not found"),
              "zero_channels": ("ValueError", "ambiguous channel assignment"), "comm_cut": ("IOError", "I/O error reading stream"),
              "odd_chunk_without_pad": ("ValueError", "invalid AIFF chunk")}[name]
    assert py == expect


def test_probe_missing_file_is_ioerror(tmp_path, built):
    import b200flac
    with pytest.raises(IOError):
        b200flac.wave_probe(os.path.join(str(tmp_path), "nope.wav"))
    with pytest.raises(IOError):
        _at().wav.WaveReader(os.path.join(str(tmp_path), "nope.wav"))


def test_wave_reader_framelists(tmp_path, built):
    """read(): 8-bit WAVE is unsigned, 16/24-bit signed little-endian (wav.py:523-527); the stream ends
    with an empty FrameList after exactly total_pcm_frames frames; seek() repositions"""
    at = _at()
    for bps, ch in ((8, 1), (16, 2), (24, 3)):
        pcm = helpers.synth_pcm(3, ch, bps, 1000)
        path = os.path.join(str(tmp_path), "r%d.wav" % bps)
        open(path, "wb").write(wave([chunk(b"fmt ", fmt_plain(ch, 44100, bps)), chunk(b"data", le_to_wave_data(pcm, bps) + (b"\x55" if ch * bps > 8 else b""))]))
        r = at.wav.WaveAudio(path).to_pcm()
        got = []
        while True:
            f = r.read(333)
            if f.frames == 0:
                break
            assert f.frames <= 333 and f.channels == ch and f.bits_per_sample == bps
            got.extend(list(f))
        assert got == list(helpers.unpack_pcm(pcm, bps))
        assert r.seek(990) == 990 and r.read(4096).frames == 10 and r.seek(5000) == 1000
        span = r.b200_file_span()
        assert span[0] == path and span[2] == 0 and span[3] == (2 if bps == 8 else 0)
        r.close()


def test_aiff_reader_framelists(tmp_path, built):
    at = _at()
    for bps, ch in ((8, 1), (16, 2), (24, 2)):
        pcm = helpers.synth_pcm(4, ch, bps, 500)
        path = os.path.join(str(tmp_path), "r%d.aiff" % bps)
        open(path, "wb").write(aiff([chunk(b"COMM", comm(ch, 500, bps), True),
                                     chunk(b"SSND", bytes(8) + le_to_aiff_data(pcm, bps), True)]))
        r = at.aiff.AiffAudio(path).to_pcm()
        got = []
        while True:
            f = r.read(128)
            if f.frames == 0:
                break
            got.extend(list(f))
        assert got == list(helpers.unpack_pcm(pcm, bps))
        r.close()


def test_truncated_data_chunk_raises_like_the_reference(tmp_path, built):
    at = _at()
    path = os.path.join(str(tmp_path), "t.wav")
    open(path, "wb").write(WAVE_CASES["data_claims_more_than_file"])
    r = at.wav.WaveReader(path)
    assert r.read(100).frames == 100
    with pytest.raises(IOError, match="premature end of data chunk"):
        r.read(4096)


# ------------------------------------------------------------------------------------------ GPU
def _write_container(kind, path, pcm, rate, ch, bps, mask=None, extra=False):
    if kind == "wave":
        fmt = fmt_extensible(ch, rate, bps, mask) if mask is not None else fmt_plain(ch, rate, bps)
        chunks = [chunk(b"fmt ", fmt)] + ([chunk(b"LIST", b"INFOabc")] if extra else []) + \
                 [chunk(b"data", le_to_wave_data(pcm, bps)), chunk(b"tail", b"xyz")]
        open(path, "wb").write(wave(chunks))
    else:
        import audiotools.aiff  # noqa: F401
        rate_bytes = {44100: b"\x40\x0e\xac\x44\x00\x00\x00\x00\x00\x00", 96000: b"\x40\x0f\xbb\x80\x00\x00\x00\x00\x00\x00"}[rate]
        n = len(pcm) // (ch * bps // 8)
        chunks = ([chunk(b"NAME", b"abc", True)] if extra else []) + \
                 [chunk(b"COMM", comm(ch, n, bps, rate_bytes), True), chunk(b"SSND", bytes(8) + le_to_aiff_data(pcm, bps), True)]
        open(path, "wb").write(aiff(chunks))


@pytest.mark.gpu
@pytest.mark.parametrize("kind,rate,ch,bps,n,level", [
    ("wave", 44100, 2, 16, 4096 * 2100 + 77, "8"),     # > one 2048-block batch: the feed crosses lanes
    ("wave", 44100, 1, 8, 50001, "5"),                  # unsigned 8-bit
    ("wave", 96000, 6, 24, 4608 * 40 + 5, "6"),         # plain 6-channel: mask 0x3F from the channel count
    ("wave", 44100, 2, 24, 100000, "8"),
    ("aiff", 44100, 2, 16, 4096 * 30 + 1, "8"),         # big-endian source
    ("aiff", 96000, 2, 24, 77777, "4"),
    ("aiff", 44100, 1, 8, 4096, "0"),
])
def test_from_pcm_file_feed_equals_framelist_feed(kind, rate, ch, bps, n, level, tmp_path, built, monkeypatch):
    at = _at()
    import audiotools.wav
    import audiotools.aiff
    pcm = helpers.synth_pcm(90 + bps + ch, ch, bps, n)
    src = os.path.join(str(tmp_path), "in." + ("wav" if kind == "wave" else "aiff"))
    _write_container(kind, src, pcm, rate, ch, bps, extra=True)
    audio = audiotools.wav.WaveAudio(src) if kind == "wave" else audiotools.aiff.AiffAudio(src)
    assert (audio.sample_rate(), audio.channels(), audio.bits_per_sample(), audio.total_frames()) == (rate, ch, bps, n)

    fast = os.path.join(str(tmp_path), "fast.flac")
    reader = audio.to_pcm()
    at.FlacAudio.from_pcm(fast, reader, level)
    assert reader.remaining_pcm_frames == 0          # left as if read to the end
    data = open(fast, "rb").read()

    monkeypatch.setenv("B200FLAC_FILE_FEED", "0")
    slow = os.path.join(str(tmp_path), "slow.flac")
    at.FlacAudio.from_pcm(slow, audio.to_pcm(), level)
    assert open(slow, "rb").read() == data

    si = helpers.streaminfo(data)
    assert si["md5"] == hashlib.md5(pcm).digest() and si["total_samples"] == n
    opts = helpers.options(**{k: v for k, v in at.FlacAudio.ENCODING_OPTIONS[level].items()})
    want = helpers.oracle_encode(pcm, rate, ch, bps, opts)
    assert data[helpers.first_frame_offset(data):] == want[helpers.first_frame_offset(want):]
    if helpers.have_ref():
        assert helpers.ref_decode(data) == pcm


@pytest.mark.gpu
def test_encode_wave_one_call_equals_stream_write(tmp_path, built):
    """b200flac_encode_wave (file to file in C) == b200flac_encode_file of the same PCM"""
    import b200flac
    n = 4096 * 50 + 123
    pcm = helpers.synth_pcm(5, 2, 16, n)
    src = os.path.join(str(tmp_path), "in.wav")
    _write_container("wave", src, pcm, 44100, 2, 16, mask=0x3)
    p = b200flac.make_params(1, 1, 8, block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True)
    a = os.path.join(str(tmp_path), "a.flac")
    info, offsets = b200flac.encode_container(a, src, p, "wave")
    assert (info.sample_rate, info.channels, info.bits_per_sample, info.channel_mask, info.total_pcm_frames) == (44100, 2, 16, 3, n)
    p2 = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True)
    b = os.path.join(str(tmp_path), "b.flac")
    b200flac.encode_file(b, p2, pcm, n)
    assert open(a, "rb").read() == open(b, "rb").read()
    assert len(offsets) == 51 and offsets[0] == (0, 4096) and offsets[-1][1] == 123


@pytest.mark.gpu
def test_truncated_wave_fails_like_the_reader(tmp_path, built, monkeypatch):
    """a data chunk that claims more than the file holds: from_pcm raises EncodingError and removes the
    partial file with either feed (wav.py:516-518 -> flac.py:1840-1845)"""
    at = _at()
    import audiotools.wav
    src = os.path.join(str(tmp_path), "t.wav")
    pcm = helpers.synth_pcm(1, 2, 16, 20000)
    open(src, "wb").write(wave([chunk(b"fmt ", fmt_plain(2, 44100, 16)), b"data" + struct.pack("<I", 4 * 30000) + pcm]))
    for knob in ("1", "0"):
        monkeypatch.setenv("B200FLAC_FILE_FEED", knob)
        out = os.path.join(str(tmp_path), "t%s.flac" % knob)
        with pytest.raises(at.EncodingError, match="premature end of data chunk"):
            at.FlacAudio.from_pcm(out, audiotools.wav.WaveReader(src), "8")
        assert not os.path.exists(out)


# ------------------------------------------------------------------------------------------ WAVE output
def test_wave_header_matches_reference_layout(built):
    """wave_header (wav.py:357-418): plain fmt for <= 2 channels at <= 16 bits, WAVEFORMATEXTENSIBLE otherwise with
    the default channel mask when none is given; parsed back by WaveReader's own fmt parser"""
    import io
    at = _at()
    h = at.wav.wave_header(44100, 2, 0x3, 16, 1000)
    assert h == (b"RIFF" + struct.pack("<I", 4 + 8 + 16 + 8 + 4000) + b"WAVE" + b"fmt " + struct.pack("<I", 16) +
                 fmt_plain(2, 44100, 16) + b"data" + struct.pack("<I", 4000))
    h = at.wav.wave_header(96000, 6, 0, 24, 7)
    assert h[20:22] == b"\xfe\xff" and len(h) == 12 + 8 + 40 + 8
    assert at.wav.parse_fmt(io.BytesIO(h[20:60])) == (6, 96000, 24, 0x3F)
    assert struct.unpack("<I", h[4:8])[0] == 4 + 8 + 40 + 8 + 126 and struct.unpack("<I", h[64:68])[0] == 126
    h = at.wav.wave_header(8000, 1, 0x4, 8, 3)           # odd data size: counted padded in the RIFF size
    assert struct.unpack("<I", h[4:8])[0] == 4 + 8 + 16 + 8 + 3 + 1
    assert at.wav.parse_fmt(io.BytesIO(at.wav.wave_header(44100, 2, 0x3, 24, 1)[20:60])) == (2, 44100, 24, 0x3)
    with pytest.raises(ValueError, match="total size too large for wave file"):
        at.wav.wave_header(44100, 2, 0x3, 16, 2 ** 30)


def test_wave_from_pcm_writes_what_the_reader_reads(tmp_path, built):
    """WaveAudio.from_pcm (wav.py:660-729) with and without total_pcm_frames; the pad byte follows the parity of
    the frame count (the reference's rule); a wrong total removes the file"""
    at = _at()
    for bps, ch, n in ((16, 2, 1001), (8, 1, 333), (24, 3, 500)):
        pcm = helpers.synth_pcm(6, ch, bps, n)
        for total in (None, n):
            path = os.path.join(str(tmp_path), "o%d_%s.wav" % (bps, total))
            w = at.WaveAudio.from_pcm(path, at.PCMBytesReader(pcm, 44100, ch, 0, bps), total_pcm_frames=total)
            assert (w.sample_rate(), w.channels(), w.bits_per_sample(), w.total_frames()) == (44100, ch, bps, n)
            data = open(path, "rb").read()
            head = at.wav.wave_header(44100, ch, 0, bps, n)
            assert data[:len(head)] == head
            assert data[len(head):len(head) + len(pcm)] == le_to_wave_data(pcm, bps)
            assert len(data) == len(head) + len(pcm) + (n % 2)
    with pytest.raises(at.EncodingError, match="total_pcm_frames mismatch"):
        at.WaveAudio.from_pcm(os.path.join(str(tmp_path), "bad.wav"), at.PCMBytesReader(pcm, 44100, 3, 0, 24), total_pcm_frames=499)
    assert not os.path.exists(os.path.join(str(tmp_path), "bad.wav"))


@pytest.mark.gpu
@pytest.mark.parametrize("rate,ch,bps,n", [(44100, 2, 16, 4096 * 4 + 33), (8000, 1, 8, 5001), (96000, 6, 24, 4608 * 3 + 10)])
def test_flac_to_wave_c_call_equals_python_path(rate, ch, bps, n, tmp_path, built):
    """b200flac_decode_to_wave == WaveAudio.from_pcm(wave, FlacAudio(flac).to_pcm()), byte for byte, and
    WAVE -> FLAC -> WAVE returns the WAVE it started from"""
    import b200flac
    at = _at()
    pcm = helpers.synth_pcm(44, ch, bps, n)
    src = os.path.join(str(tmp_path), "src.wav")
    at.WaveAudio.from_pcm(src, at.PCMBytesReader(pcm, rate, ch, 0, bps))
    flac = os.path.join(str(tmp_path), "x.flac")
    at.FlacAudio.from_pcm(flac, at.WaveAudio(src).to_pcm(), "8")
    a, b = os.path.join(str(tmp_path), "a.wav"), os.path.join(str(tmp_path), "b.wav")
    b200flac.decode_to_wave(flac, a)
    at.WaveAudio.from_pcm(b, at.FlacAudio(flac).to_pcm())
    assert open(a, "rb").read() == open(b, "rb").read() == open(src, "rb").read()
