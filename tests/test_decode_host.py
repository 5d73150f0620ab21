"""CPU: the host side of the decode layer -- STREAMINFO / channel mask as flacdec_read_metadata reads them
(src/decoders/flac.c:569-708, 509-566), the reference's metadata errors, and the no-fallback rule."""
import io
import struct

import pytest

import helpers


def _flac(channels=2, bps=16, rate=44100, n=1000, comments=(), extra_blocks=()):
    pcm = helpers.synth_pcm(1, channels, bps, n)
    data = helpers.oracle_encode(pcm, rate, channels, bps, helpers.options())
    if not comments and not extra_blocks:
        return data
    # rebuild the metadata: STREAMINFO, a VORBIS_COMMENT with the given entries, the extra blocks, then the frames
    first = helpers.first_frame_offset(data)
    vendor = b"test"
    body = struct.pack("<I", len(vendor)) + vendor + struct.pack("<I", len(comments)) + \
        b"".join(struct.pack("<I", len(c)) + c for c in comments)
    blocks = [(4, body)] + list(extra_blocks)
    out = data[:4] + bytes([0]) + data[5:8] + data[8:42]
    for i, (bid, payload) in enumerate(blocks):
        last = 0x80 if i == len(blocks) - 1 else 0
        out += bytes([last | bid]) + len(payload).to_bytes(3, "big") + payload
    return out + data[first:]


def test_streaminfo_fields(built):
    import b200flac
    data = _flac(2, 16, 44100, 12345)
    info = b200flac.read_streaminfo(data)
    si = helpers.streaminfo(data)
    assert (info.sample_rate, info.channels, info.bits_per_sample, info.total_pcm_frames) == (44100, 2, 16, 12345)
    assert bytes(info.md5) == si["md5"] and info.first_frame_offset == helpers.first_frame_offset(data)
    assert info.min_block_size == info.max_block_size == 4096 and info.channel_mask == 0x3


@pytest.mark.parametrize("channels,mask", [(1, 0x4), (2, 0x3), (3, 0x7), (4, 0x33), (5, 0x37), (6, 0x3F), (7, 0x70F), (8, 0x63F)])
def test_default_channel_mask_by_count(channels, mask, built):
    import b200flac
    assert b200flac.read_streaminfo(_flac(channels, 16, 48000, 100)).channel_mask == mask


def test_vorbis_comment_channel_mask_override(built):
    import b200flac
    # taken when its bit count equals the channel count; compared in upper case; otherwise ignored
    assert b200flac.read_streaminfo(_flac(6, 24, 96000, 100, [b"TITLE=x", b"WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x060F"])).channel_mask == 0x60F
    assert b200flac.read_streaminfo(_flac(6, 24, 96000, 100, [b"waveformatextensible_channel_mask=0x060f"])).channel_mask == 0x60F
    assert b200flac.read_streaminfo(_flac(6, 24, 96000, 100, [b"WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x0003"])).channel_mask == 0x3F
    assert b200flac.read_streaminfo(_flac(2, 16, 44100, 100, [b"WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x0030"], [(1, bytes(10))])).channel_mask == 0x30


def test_metadata_errors_like_the_reference(built):
    import b200flac
    data = _flac()
    with pytest.raises(ValueError, match="not a FLAC file"):
        b200flac.read_streaminfo(b"RIFF" + data[4:])
    with pytest.raises(IOError, match="EOF while reading metadata"):
        b200flac.read_streaminfo(data[:30])
    with pytest.raises(IOError, match="EOF while reading metadata"):
        b200flac.read_streaminfo(data[:4] + bytes([0x00]) + data[5:60])   # last-block flag cleared, then nothing


def test_decode_without_gpu_fails_loudly(built):
    import audiotools.decoders
    import b200flac
    if b200flac.device_count() > 0:
        pytest.skip("a CUDA device is present")
    d = audiotools.decoders.FlacDecoder(io.BytesIO(_flac()))
    assert (d.sample_rate, d.channels, d.bits_per_sample, d.channel_mask) == (44100, 2, 16, 0x3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d.read(4096)


def test_reference_fixture_streaminfo(built):
    """STREAMINFO of the reference's fixtures: metadata blocks in any order, channel masks by channel count
    (test/flac-nomask*.flac, flac-disordered.flac), an ID3 prefix is not FLAC for the C entry points"""
    import os
    import b200flac
    d = os.path.join(helpers.ROOT, "tests", "golden", "flac")
    got = {}
    for name in ("flac-allframes.flac", "flac-disordered.flac", "flac-nomask1.flac", "flac-nomask3.flac", "1h.flac"):
        i = b200flac.read_streaminfo(open(os.path.join(d, name), "rb").read())
        got[name] = (i.sample_rate, i.channels, i.bits_per_sample, i.total_pcm_frames, i.max_block_size, i.channel_mask)
    assert got == {"flac-allframes.flac": (44100, 1, 16, 80, 4096, 0x4), "flac-disordered.flac": (44100, 2, 16, 304844, 4096, 0x3),
                   "flac-nomask1.flac": (44100, 6, 16, 44100, 4096, 0x3F), "flac-nomask3.flac": (44100, 2, 24, 44100, 4096, 0x3),
                   "1h.flac": (8000, 2, 16, 28800000, 32768, 0x3)}
    with pytest.raises(ValueError, match="not a FLAC file"):
        b200flac.read_streaminfo(open(os.path.join(d, "flac-id3.flac"), "rb").read())


def test_crafted_total_is_rejected_before_allocation(built, tmp_path):
    """a 42-byte file whose STREAMINFO claims 2^36 - 1 PCM frames of 8 x 24-bit channels (1.6 TB): the C entry
    points report what the reference's frame loop would hit -- EOF -- instead of sizing a buffer from the
    header (round-1 advisor finding); needs no GPU"""
    import os
    import b200flac
    si = bytearray(34)
    si[0:2] = si[2:4] = (4096).to_bytes(2, "big")
    v = (96000 << 44) | (7 << 41) | (23 << 36) | ((1 << 36) - 1)
    si[10:18] = v.to_bytes(8, "big")
    data = b"fLaC" + bytes([0x80]) + (34).to_bytes(3, "big") + bytes(si)
    assert len(data) == 42
    path = os.path.join(str(tmp_path), "bomb.flac")
    open(path, "wb").write(data)
    with pytest.raises(IOError, match="EOF"):
        b200flac.verify_file(path)
    with pytest.raises(IOError, match="EOF"):
        b200flac.decode_to_wave(path, os.path.join(str(tmp_path), "bomb.wav"))
    with pytest.raises(IOError, match="EOF"):
        b200flac.decode(data)
    with pytest.raises(IOError):
        b200flac.verify_file(os.path.join(str(tmp_path), "missing.flac"))
