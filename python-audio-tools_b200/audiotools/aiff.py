"""audiotools.aiff -- the AIFF reader that feeds FlacAudio.from_pcm, under Python 3.

Reference: audiotools/aiff.py of widgital/python-audio-tools (Python 2 only):
  parse_ieee_extended  aiff.py:25-36
  parse_comm           aiff.py:327-347
  AiffReader           aiff.py:350-482
  AiffAudio            aiff.py:486-560 (only what `AiffAudio(filename).to_pcm()` needs; the container
                       object here takes its fields from AiffReader instead of its own chunk iterator)

Like audiotools.wav.WaveReader, AiffReader also publishes where its PCM lies in the file
(`b200_file_span`) so encode_flac can feed the engine from the file directly; AIFF samples are
big-endian and are byte-swapped into the pinned staging (b200flac_stream_write_file, flag 1).
The chunk walk exists in C as b200flac_aiff_probe.
"""
import struct

from .pcm import FrameList

ERR_AIFF_NOT_AIFF = "not an AIFF file"                    # audiotools/text.py:530-544
ERR_AIFF_INVALID_AIFF = "invalid AIFF file"
ERR_AIFF_INVALID_CHUNK_ID = "invalid AIFF chunk ID"
ERR_AIFF_INVALID_CHUNK = "invalid AIFF chunk"
ERR_AIFF_PREMATURE_SSND_CHUNK = "SSND chunk found before fmt"
ERR_AIFF_NO_SSND_CHUNK = "SSND chunk not found"
ERR_AIFF_TRUNCATED_SSND_CHUNK = "premature end of SSND chunk"
ERR_NEGATIVE_SEEK = "cannot seek to negative value"

PRINTABLE_ASCII = frozenset(range(0x20, 0x7E + 1))        # aiff.py:498


class InvalidAIFF(ValueError):
    pass


def parse_ieee_extended(data):
    """the 80-bit IEEE extended value AIFF stores its sample rate in (aiff.py:25-36)"""
    (sign_exponent, mantissa) = struct.unpack(">HQ", data)
    signed = sign_exponent >> 15
    exponent = sign_exponent & 0x7FFF
    if (exponent == 0) and (mantissa == 0):
        return 0
    elif exponent == 0x7FFF:
        return 1.79769313486231e+308
    else:
        f = mantissa * (2.0 ** (exponent - 16383 - 63))
        return f if not signed else -f


def parse_comm(f):
    """given a file positioned after the COMM chunk header, returns
    (channels, total_sample_frames, bits_per_sample, sample_rate, channel_mask) (aiff.py:327-347)"""
    data = f.read(18)
    if len(data) < 18:
        raise IOError("I/O error reading stream")
    (channels, total_sample_frames, bits_per_sample) = struct.unpack(">HIH", data[0:8])
    sample_rate = int(parse_ieee_extended(data[8:18]))
    if channels <= 2:
        # ChannelMask.from_channels, audiotools/__init__.py:2049-2060
        if channels == 2:
            channel_mask = 0x3
        elif channels == 1:
            channel_mask = 0x4
        else:
            raise ValueError("ambiguous channel assignment")
    else:
        channel_mask = 0
    return (channels, total_sample_frames, bits_per_sample, sample_rate, channel_mask)


class AiffReader(object):
    """a PCMReader object for reading AIFF file contents (aiff.py:350-482)"""

    def __init__(self, aiff_filename):
        self.file = open(aiff_filename, "rb")
        try:
            self.__walk__()
        except Exception:
            self.file.close()
            raise

    def __walk__(self):
        try:
            (form, total_size, aiff) = struct.unpack(">4sI4s", self.file.read(12))
        except struct.error:
            raise InvalidAIFF(ERR_AIFF_INVALID_AIFF)
        if form != b"FORM":
            raise ValueError(ERR_AIFF_NOT_AIFF)
        elif aiff != b"AIFF":
            raise ValueError(ERR_AIFF_INVALID_AIFF)
        total_size -= 4
        comm_chunk_read = False

        # walk through chunks until "SSND" chunk encountered
        while total_size > 0:
            try:
                (chunk_id, chunk_size) = struct.unpack(">4sI", self.file.read(8))
            except struct.error:
                raise ValueError(ERR_AIFF_INVALID_AIFF)
            if not frozenset(chunk_id).issubset(PRINTABLE_ASCII):
                raise ValueError(ERR_AIFF_INVALID_CHUNK_ID)
            total_size -= 8

            if chunk_id == b"COMM":
                (self.channels, self.total_pcm_frames, self.bits_per_sample, self.sample_rate,
                 self.channel_mask) = parse_comm(self.file)
                self.bytes_per_pcm_frame = (self.bits_per_sample // 8) * self.channels
                self.remaining_pcm_frames = self.total_pcm_frames
                comm_chunk_read = True
            elif chunk_id == b"SSND":
                if not comm_chunk_read:
                    raise ValueError(ERR_AIFF_PREMATURE_SSND_CHUNK)
                self.file.read(8)                   # the "offset" and "block_size" attributes
                self.ssnd_chunk_offset = self.file.tell()
                return
            else:
                self.file.read(chunk_size)          # all other chunks are ignored

            if chunk_size % 2:
                if len(self.file.read(1)) < 1:
                    raise ValueError(ERR_AIFF_INVALID_CHUNK)
                total_size -= (chunk_size + 1)
            else:
                total_size -= chunk_size
        raise ValueError(ERR_AIFF_NO_SSND_CHUNK)

    def read(self, pcm_frames):
        """try to read a pcm.FrameList with the given number of PCM frames (aiff.py:434-456)"""
        requested_pcm_frames = min(max(pcm_frames, 1), self.remaining_pcm_frames)
        requested_bytes = self.bytes_per_pcm_frame * requested_pcm_frames
        pcm_data = self.file.read(requested_bytes)
        if len(pcm_data) < requested_bytes:
            raise IOError(ERR_AIFF_TRUNCATED_SSND_CHUNK)
        self.remaining_pcm_frames -= requested_pcm_frames
        return FrameList(pcm_data, self.channels, self.bits_per_sample, True, True)

    def seek(self, pcm_frame_offset):
        """tries to seek to the given PCM frame offset (aiff.py:458-477)"""
        if pcm_frame_offset < 0:
            raise ValueError(ERR_NEGATIVE_SEEK)
        pcm_frame_offset = min(pcm_frame_offset, self.total_pcm_frames)
        self.file.seek(self.ssnd_chunk_offset + pcm_frame_offset * self.bytes_per_pcm_frame, 0)
        self.remaining_pcm_frames = self.total_pcm_frames - pcm_frame_offset
        return pcm_frame_offset

    def close(self):
        self.file.close()

    # ---- file-span protocol of the B200 engine (not in the reference) ----
    def b200_file_span(self):
        """(path, byte offset, PCM frames, flags, text of the truncation error); flag 1 = big-endian samples"""
        return (self.file.name, self.file.tell(), self.remaining_pcm_frames, 1, ERR_AIFF_TRUNCATED_SSND_CHUNK)

    def b200_file_span_consumed(self, pcm_frames):
        self.file.seek(pcm_frames * self.bytes_per_pcm_frame, 1)
        self.remaining_pcm_frames -= pcm_frames


class AiffAudio(object):
    """an AIFF file, as far as FlacAudio.from_pcm(filename, AiffAudio(path).to_pcm()) needs it"""
    SUFFIX = "aiff"
    NAME = SUFFIX
    PRINTABLE_ASCII = PRINTABLE_ASCII

    def __init__(self, filename):
        self.filename = filename
        r = AiffReader(filename)
        (self.__channels__, self.__sample_rate__, self.__bits_per_sample__, self.__channel_mask__,
         self.__total_sample_frames__) = (r.channels, r.sample_rate, r.bits_per_sample, r.channel_mask,
                                          r.total_pcm_frames)
        r.close()

    def lossless(self):
        return True

    def channel_mask(self):
        return self.__channel_mask__

    def to_pcm(self):
        """returns a PCMReader object containing the track's PCM data"""
        return AiffReader(self.filename)

    def total_frames(self):
        return self.__total_sample_frames__

    def sample_rate(self):
        return self.__sample_rate__

    def channels(self):
        return self.__channels__

    def bits_per_sample(self):
        return self.__bits_per_sample__
