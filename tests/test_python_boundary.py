"""The drop-in boundary: audiotools.pcm.FrameList, audiotools.encoders.encode_flac and
FlacAudio.from_pcm under Python 3 (reference: src/pcm.c, src/encoders/flac.c:43-307,
audiotools/flac.py:1695-1845; tests restate test/test_formats.py FlacFileTest)."""
import hashlib
import inspect
import os

import numpy as np
import pytest

import helpers


def _at():
    import audiotools
    return audiotools


# ------------------------------------------------------------------ CPU-only checks
def test_framelist_basics(built):
    from audiotools import pcm
    f = pcm.FrameList(bytes([1, 0, 254, 255, 3, 0, 4, 0]), 2, 16, False, True)
    assert (f.frames, f.channels, f.bits_per_sample, len(f)) == (2, 2, 16, 4)
    assert list(f) == [1, -2, 3, 4]
    assert f.to_bytes(False, True) == bytes([1, 0, 254, 255, 3, 0, 4, 0])
    assert f.to_bytes(True, True) == bytes([0, 1, 255, 254, 0, 3, 0, 4])
    g = pcm.from_list([5, 6], 2, 16, True)
    h = f + g
    assert list(h) == [1, -2, 3, 4, 5, 6] and h.frames == 3
    a, b = h.split(1)
    assert list(a) == [1, -2] and list(b) == [3, 4, 5, 6]
    a, b = h.split(10)
    assert a == h and b.frames == 0
    assert list(h.channel(1)) == [-2, 4, 6] and list(h.frame(2)) == [5, 6]
    f24 = pcm.FrameList(bytes([0xFF, 0xFF, 0xFF, 0x00, 0x00, 0x80]), 1, 24, False, True)
    assert list(f24) == [-1, -8388608]
    with pytest.raises(ValueError):
        pcm.FrameList(b"\0\0\0", 2, 16, False, True)
    with pytest.raises(ValueError):
        f + pcm.from_list([1], 1, 16, True)


def test_buffered_pcmreader_exact_reads(built):
    at = _at()
    data = helpers.synth_pcm(3, 2, 16, 10000)

    class Odd(at.PCMBytesReader):
        def read(self, n):
            return at.PCMBytesReader.read(self, 777)
    r = at.BufferedPCMReader(Odd(data, 44100, 2, 0x3, 16))
    got, sizes = b"", []
    while True:
        f = r.read(4096)
        if f.frames == 0:
            break
        sizes.append(f.frames)
        got += f.to_bytes(False, True)
    assert sizes == [4096, 4096, 1808] and got == data


def test_encode_flac_signature_matches_reference(built):
    from audiotools import encoders
    doc = encoders.encode_flac.__doc__
    # kwlist of src/encoders/flac.c:52-67, in order
    names = ["filename", "pcmreader", "block_size", "max_lpc_order", "min_residual_partition_order",
             "max_residual_partition_order", "mid_side", "adaptive_mid_side", "exhaustive_model_search",
             "disable_verbatim_subframes", "disable_constant_subframes", "disable_fixed_subframes",
             "disable_lpc_subframes", "padding_size"]
    pos = [doc.index(n) for n in names]
    assert pos == sorted(pos)
    at = _at()
    sig = inspect.signature(at.FlacAudio.from_pcm)
    assert list(sig.parameters) == ["filename", "pcmreader", "compression", "total_pcm_frames", "encoding_function"]
    with pytest.raises(TypeError):
        encoders.encode_flac("x.flac")  # required arguments, format "sO&IIII|..."


def test_encode_flac_without_gpu_fails_loudly(built, tmp_path):
    at = _at()
    from audiotools import encoders
    if encoders.b200_device_count() > 0:
        pytest.skip("a GPU is present")
    r = at.PCMBytesReader(helpers.synth_pcm(1, 2, 16, 100), 44100, 2, 0x3, 16)
    # an IOError, so that FlacAudio.from_pcm's EncodingError contract holds (flac.py:1833-1845)
    with pytest.raises(IOError) as e:
        encoders.encode_flac(os.path.join(str(tmp_path), "a.flac"), r, 4096, 8, 0, 5)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(at.EncodingError):
        at.FlacAudio.from_pcm(os.path.join(str(tmp_path), "b.flac"),
                              at.PCMBytesReader(helpers.synth_pcm(1, 2, 16, 100), 44100, 2, 0x3, 16))
    assert not os.path.exists(os.path.join(str(tmp_path), "b.flac"))   # from_pcm unlinks the partial file
    with pytest.raises(IOError):
        encoders.encode_flac(os.path.join(str(tmp_path), "missing", "a.flac"), r, 4096, 8, 0, 5)


# ------------------------------------------------------------------ GPU checks
@pytest.mark.gpu
@pytest.mark.parametrize("level", list("012345678"))
def test_from_pcm_every_compression_level(level, tmp_path, built):
    """FlacFileTest.test_option_variations / __test_reader__ (test_formats.py:3578-3620,3773-3795):
    STREAMINFO MD5 == reader MD5, decoder output == input; plus frames == the oracle's, byte for byte"""
    at = _at()
    n = 44100 * 12 + 321
    pcm = helpers.synth_pcm(40 + int(level), 2, 16, n)
    path = os.path.join(str(tmp_path), "l%s.flac" % level)
    flac = at.FlacAudio.from_pcm(path, at.PCMBytesReader(pcm, 44100, 2, 0x3, 16), level, total_pcm_frames=n)
    data = open(path, "rb").read()
    si = helpers.streaminfo(data)
    assert si["md5"] == hashlib.md5(pcm).digest() and si["total_samples"] == n
    assert (flac.sample_rate(), flac.channels(), flac.bits_per_sample(), flac.total_frames()) == (44100, 2, 16, n)
    opts = helpers.options(**{k: v for k, v in at.FlacAudio.ENCODING_OPTIONS[level].items()})
    want, offs = helpers.oracle_encode(pcm, 44100, 2, 16, opts, want_offsets=True)
    ff_got, ff_want = helpers.first_frame_offset(data), helpers.first_frame_offset(want)
    assert data[ff_got:] == want[ff_want:]
    # metadata: STREAMINFO, SEEKTABLE, VORBIS_COMMENT, PADDING in the reference's preferred order
    blocks = [b[0] for b in flac.get_metadata().block_list]
    assert blocks == [0, 3, 4, 1]
    seek = flac.get_metadata().get_blocks(3)[0][1]
    assert len(seek) == 18 * 2  # one seek point per 10 s
    first_point = (int.from_bytes(seek[0:8], "big"), int.from_bytes(seek[8:16], "big"), int.from_bytes(seek[16:18], "big"))
    assert first_point == (0, 0, offs[0][1])
    if helpers.have_ref():
        assert helpers.ref_decode(data) == pcm
    # the C finalisation (default) and the Python restatement of flac.py:1811-1832 write the same file
    at.FlacAudio.NATIVE_FINALIZE = False
    try:
        path2 = os.path.join(str(tmp_path), "p%s.flac" % level)
        at.FlacAudio.from_pcm(path2, at.PCMBytesReader(pcm, 44100, 2, 0x3, 16), level, total_pcm_frames=n)
    finally:
        at.FlacAudio.NATIVE_FINALIZE = True
    assert open(path2, "rb").read() == data


@pytest.mark.gpu
def test_encode_flac_returns_reference_offsets(tmp_path, built):
    at = _at()
    from audiotools import encoders
    pcm = helpers.synth_pcm(9, 2, 24, 30000)
    path = os.path.join(str(tmp_path), "o.flac")
    offs = encoders.encode_flac(path, pcmreader=at.BufferedPCMReader(at.PCMBytesReader(pcm, 96000, 2, 0x3, 24)),
                                block_size=4096, max_lpc_order=12, min_residual_partition_order=0,
                                max_residual_partition_order=8, mid_side=True, exhaustive_model_search=True,
                                padding_size=1000)
    o = helpers.options(block_size=4096, max_lpc_order=12, max_residual_partition_order=8, mid_side=True,
                        exhaustive_model_search=True, padding_size=1000)
    want, want_offs = helpers.oracle_encode(pcm, 96000, 2, 24, o, want_offsets=True)
    assert open(path, "rb").read() == want
    assert offs == want_offs and all(isinstance(t, tuple) for t in offs)


@pytest.mark.gpu
def test_unbuffered_short_reads_make_short_frames(tmp_path, built):
    """the C entry encodes whatever read() returns (flac.c:247,525)"""
    at = _at()
    from audiotools import encoders
    pcm = helpers.synth_pcm(10, 1, 16, 3000)

    class Short(at.PCMBytesReader):
        def read(self, n):
            return at.PCMBytesReader.read(self, 700)
    offs = encoders.encode_flac(os.path.join(str(tmp_path), "s.flac"), Short(pcm, 44100, 1, 0x4, 16), 1024, 8, 0, 4)
    assert [n for _, n in offs] == [700, 700, 700, 700, 200]


@pytest.mark.gpu
def test_invalid_from_pcm_unlinks_partial_file(tmp_path, built):
    """test_formats.py:721-757 test_invalid_from_pcm with ERROR_PCM_Reader (:60-100)"""
    at = _at()

    class ErrorReader(at.PCMBytesReader):
        def __init__(self, *a):
            at.PCMBytesReader.__init__(self, *a)
            self.calls = 0

        def read(self, n):
            self.calls += 1
            if self.calls > 2:
                raise ValueError("reader failed")
            return at.PCMBytesReader.read(self, n)
    pcm = helpers.synth_pcm(11, 2, 16, 3000000)
    path = os.path.join(str(tmp_path), "err.flac")
    with pytest.raises(at.EncodingError):
        at.FlacAudio.from_pcm(path, ErrorReader(pcm, 44100, 2, 0x3, 16), "5")
    assert not os.path.exists(path)

    class NotFrameList(at.PCMBytesReader):
        def read(self, n):
            return b"bytes are not FrameLists"
    from audiotools import encoders
    with pytest.raises(TypeError):
        encoders.encode_flac(path, NotFrameList(pcm, 44100, 2, 0x3, 16), 4096, 8, 0, 5)


@pytest.mark.gpu
def test_from_pcm_multichannel_adds_channel_mask(tmp_path, built):
    at = _at()
    pcm = helpers.synth_pcm(12, 6, 24, 20000)
    path = os.path.join(str(tmp_path), "m.flac")
    flac = at.FlacAudio.from_pcm(path, at.PCMBytesReader(pcm, 96000, 6, 0, 24), "6")
    vc = flac.get_metadata().get_blocks(4)[0][1]
    assert b"WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x003F" in vc
    if helpers.have_ref():
        assert helpers.ref_decode(open(path, "rb").read()) == pcm


def test_finalize_flac_metadata_extension_entry(tmp_path, built):
    """audiotools.encoders.finalize_flac_metadata (host only): same file as the Python tail of from_pcm"""
    import shutil
    at = _at()
    from audiotools import encoders, flac as aflac
    o = helpers.options(block_size=1152, max_lpc_order=0, max_residual_partition_order=2, padding_size=4096)
    pcm = helpers.synth_pcm(5, 2, 16, 44100 * 21)
    data, offsets = helpers.oracle_encode(pcm, 44100, 2, 16, o, want_offsets=True)
    a, b = os.path.join(str(tmp_path), "a.flac"), os.path.join(str(tmp_path), "b.flac")
    with open(a, "wb") as fh:
        fh.write(data)
    shutil.copy(a, b)
    f = aflac.FlacAudio(a)
    md = f.get_metadata()
    md.add_block(f.seektable(list(offsets), 441000))
    f.update_metadata(md)
    encoders.finalize_flac_metadata(b, offsets, seekpoint_interval=441000)
    assert open(b, "rb").read() == open(a, "rb").read()
    assert [blk[0] for blk in at.FlacAudio(b).get_metadata().block_list] == [0, 3, 4, 1]
    with pytest.raises(IOError):
        encoders.finalize_flac_metadata(os.path.join(str(tmp_path), "nope.flac"), offsets)
    with pytest.raises(TypeError):
        encoders.finalize_flac_metadata(b, [("x", 1)])


@pytest.mark.gpu
def test_encode_flac_files_equals_encode_flac(tmp_path, built):
    """audiotools.encoders.encode_flac_files (many tracks in one call): every file is byte
    for byte what encode_flac writes for the same reader; readers are closed; mismatched readers are a ValueError"""
    at = _at()
    from audiotools import encoders
    lengths = [30000, 4096 * 3, 1, 4097, 55555]
    pcms = [helpers.synth_pcm(40 + i, 2, 16, n) for i, n in enumerate(lengths)]
    opts = dict(block_size=4096, max_lpc_order=12, min_residual_partition_order=0, max_residual_partition_order=6,
                mid_side=True, exhaustive_model_search=True, padding_size=2000)
    want = []
    for i, pcm in enumerate(pcms):
        path = os.path.join(str(tmp_path), "one_%d.flac" % i)
        encoders.encode_flac(path, pcmreader=at.BufferedPCMReader(at.PCMBytesReader(pcm, 44100, 2, 0x3, 16)), **opts)
        want.append(open(path, "rb").read())

    class Closing(at.PCMBytesReader):
        closed = False

        def close(self):
            self.closed = True
    names = [os.path.join(str(tmp_path), "many_%d.flac" % i) for i in range(len(pcms))]
    readers = [Closing(pcm, 44100, 2, 0x3, 16) for pcm in pcms]
    assert encoders.encode_flac_files(names, readers, host_threads=2, **opts) is None
    for i, name in enumerate(names):
        assert open(name, "rb").read() == want[i], "track %d" % i
    assert all(r.closed for r in readers)
    with pytest.raises(ValueError):
        encoders.encode_flac_files(names[:2], [at.PCMBytesReader(pcms[0], 44100, 2, 0x3, 16),
                                               at.PCMBytesReader(pcms[1], 48000, 2, 0x3, 16)], **opts)
    with pytest.raises(ValueError):
        encoders.encode_flac_files(names[:2], readers[:1], **opts)
    with pytest.raises(IOError):
        encoders.encode_flac_files([os.path.join(str(tmp_path), "no_such_dir", "x.flac")],
                                   [at.PCMBytesReader(pcms[0], 44100, 2, 0x3, 16)], **opts)
