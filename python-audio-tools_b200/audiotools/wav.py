"""audiotools.wav -- the RIFF WAVE reader that feeds FlacAudio.from_pcm, under Python 3.

Reference: audiotools/wav.py of widgital/python-audio-tools (Python 2 only):
  parse_fmt      wav.py:288-354
  WaveReader     wav.py:421-553
  wave_header    wav.py:357-418
  WaveAudio      wav.py:580-757 (what `WaveAudio(filename).to_pcm()` and `WaveAudio.from_pcm` need)

WaveReader is the reference's PCMReader, method for method, so any consumer can call read() as
before.  In addition it says where its PCM lies in the file (`b200_file_span`), which lets
audiotools.encoders.encode_flac pull the data chunk straight into the engine's pinned staging
(b200flac_stream_write_file) instead of one FrameList per read() call.  The same
chunk walk exists in C as b200flac_wave_probe; tests/test_pcm_sources.py holds the two together.
"""
import struct

from .pcm import FrameList

ERR_WAV_NOT_WAVE = "not a RIFF WAVE file"                 # audiotools/text.py:621-634
ERR_WAV_INVALID_WAVE = "invalid RIFF WAVE file"
ERR_WAV_NO_DATA_CHUNK = "data chunk not found"
ERR_WAV_INVALID_CHUNK = "invalid RIFF WAVE chunk ID"
ERR_WAV_PREMATURE_DATA = "data chunk found before fmt"
ERR_WAV_TRUNCATED_DATA_CHUNK = "premature end of data chunk"
ERR_NEGATIVE_SEEK = "cannot seek to negative value"

PRINTABLE_ASCII = frozenset(range(0x20, 0x7E + 1))        # wav.py:587

PCM_SUB_FORMAT = b"\x01\x00\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"

# ChannelMask.from_fields(...) of wav.py:304-331 as integers
# (front_left 0x1, front_right 0x2, front_center 0x4, low_frequency 0x8, back_left 0x10, back_right 0x20)
_MASK_BY_CHANNELS = {1: 0x4, 2: 0x3, 3: 0x7, 4: 0x33, 5: 0x37, 6: 0x3F}


def parse_fmt(f):
    """given a file positioned after the `fmt ` chunk header, returns
    (channels, sample_rate, bits_per_sample, channel_mask); reads 16 or 40 bytes (wav.py:288-354)"""
    head = f.read(16)
    if len(head) < 16:
        raise IOError("I/O error reading stream")
    (compression, channels, sample_rate, bytes_per_second, block_align, bits_per_sample) = struct.unpack("<HHIIHH", head)
    if compression == 1:
        return (channels, sample_rate, bits_per_sample, _MASK_BY_CHANNELS.get(channels, 0))
    elif compression == 0xFFFE:
        ext = f.read(24)
        if len(ext) < 24:
            raise IOError("I/O error reading stream")
        (cb_size, valid_bits_per_sample, channel_mask) = struct.unpack("<HHI", ext[0:8])
        if ext[8:24] != PCM_SUB_FORMAT:
            raise ValueError("invalid WAVE sub-format")
        return (channels, sample_rate, bits_per_sample, channel_mask)
    else:
        raise ValueError("unsupported WAVE compression")


def wave_header(sample_rate, channels, channel_mask, bits_per_sample, total_pcm_frames):
    """everything before a RIFF WAVE's PCM data (wav.py:357-418): a plain `fmt ` chunk for up to 2 channels at
    up to 16 bits, WAVEFORMATEXTENSIBLE otherwise; ValueError when the file would pass 4 GB"""
    avg_bytes_per_second = sample_rate * channels * (bits_per_sample // 8)
    block_align = channels * (bits_per_sample // 8)
    if (channels <= 2) and (bits_per_sample <= 16):
        fmt = struct.pack("<HHIIHH", 1, channels, sample_rate, avg_bytes_per_second, block_align, bits_per_sample)
    else:
        if channel_mask == 0:
            channel_mask = _MASK_BY_CHANNELS.get(channels, 0)
        fmt = struct.pack("<HHIIHH", 0xFFFE, channels, sample_rate, avg_bytes_per_second, block_align, bits_per_sample) + \
            struct.pack("<HHI", 22, bits_per_sample, int(channel_mask)) + PCM_SUB_FORMAT
    data_size = (bits_per_sample // 8) * channels * total_pcm_frames
    total_size = 4 + 8 + len(fmt) + 8 + data_size + (data_size % 2)
    if total_size >= 2 ** 32:
        raise ValueError("total size too large for wave file")
    return (b"RIFF" + struct.pack("<I", total_size) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt +
            b"data" + struct.pack("<I", data_size))


class WaveReader(object):
    """a PCMReader object for reading wave file contents (wav.py:421-553)"""

    def __init__(self, wave_filename):
        self.file = open(wave_filename, "rb")
        try:
            self.__walk__()
        except Exception:
            self.file.close()
            raise

    def __walk__(self):
        try:
            (riff, total_size, wave) = struct.unpack("<4sI4s", self.file.read(12))
        except struct.error:
            raise ValueError(ERR_WAV_INVALID_WAVE)
        if riff != b"RIFF":
            raise ValueError(ERR_WAV_NOT_WAVE)
        elif wave != b"WAVE":
            raise ValueError(ERR_WAV_INVALID_WAVE)
        total_size -= 4
        fmt_chunk_read = False

        # walk through chunks until "data" chunk encountered
        while total_size > 0:
            try:
                (chunk_id, chunk_size) = struct.unpack("<4sI", self.file.read(8))
            except struct.error:
                raise ValueError(ERR_WAV_INVALID_WAVE)
            if not frozenset(chunk_id).issubset(PRINTABLE_ASCII):
                raise ValueError(ERR_WAV_INVALID_CHUNK)
            total_size -= 8

            if chunk_id == b"fmt ":
                # (the reference leaves whatever follows the parsed fields of a longer chunk unread)
                (self.channels, self.sample_rate, self.bits_per_sample, self.channel_mask) = parse_fmt(self.file)
                self.bytes_per_pcm_frame = (self.bits_per_sample // 8) * self.channels
                fmt_chunk_read = True
            elif chunk_id == b"data":
                if not fmt_chunk_read:
                    raise ValueError(ERR_WAV_PREMATURE_DATA)
                self.total_pcm_frames = chunk_size // self.bytes_per_pcm_frame
                self.remaining_pcm_frames = self.total_pcm_frames
                self.data_chunk_offset = self.file.tell()
                return
            else:
                self.file.read(chunk_size)          # all other chunks are ignored

            if chunk_size % 2:
                if len(self.file.read(1)) < 1:
                    raise ValueError(ERR_WAV_INVALID_CHUNK)
                total_size -= (chunk_size + 1)
            else:
                total_size -= chunk_size
        raise ValueError(ERR_WAV_NO_DATA_CHUNK)

    def read(self, pcm_frames):
        """try to read a pcm.FrameList with the given number of PCM frames (wav.py:504-527)"""
        requested_pcm_frames = min(max(pcm_frames, 1), self.remaining_pcm_frames)
        requested_bytes = self.bytes_per_pcm_frame * requested_pcm_frames
        pcm_data = self.file.read(requested_bytes)
        if len(pcm_data) < requested_bytes:
            raise IOError(ERR_WAV_TRUNCATED_DATA_CHUNK)
        self.remaining_pcm_frames -= requested_pcm_frames
        return FrameList(pcm_data, self.channels, self.bits_per_sample, False, self.bits_per_sample != 8)

    def seek(self, pcm_frame_offset):
        """tries to seek to the given PCM frame offset; returns the frames actually seeked over (wav.py:529-548)"""
        if pcm_frame_offset < 0:
            raise ValueError(ERR_NEGATIVE_SEEK)
        pcm_frame_offset = min(pcm_frame_offset, self.total_pcm_frames)
        self.file.seek(self.data_chunk_offset + pcm_frame_offset * self.bytes_per_pcm_frame, 0)
        self.remaining_pcm_frames = self.total_pcm_frames - pcm_frame_offset
        return pcm_frame_offset

    def close(self):
        self.file.close()

    # ---- file-span protocol of the B200 engine (not in the reference) ----
    def b200_file_span(self):
        """(path, byte offset, PCM frames, flags, text of the truncation error) of the PCM this reader has not handed out yet;
        flags: 2 = unsigned samples (8-bit WAVE, wav.py:527)"""
        return (self.file.name, self.file.tell(), self.remaining_pcm_frames, 2 if self.bits_per_sample == 8 else 0,
                ERR_WAV_TRUNCATED_DATA_CHUNK)

    def b200_file_span_consumed(self, pcm_frames):
        """the engine took pcm_frames frames of the span by itself: same state as after read() calls"""
        self.file.seek(pcm_frames * self.bytes_per_pcm_frame, 1)
        self.remaining_pcm_frames -= pcm_frames


class WaveAudio(object):
    """a RIFF WAVE file, as far as FlacAudio.from_pcm(filename, WaveAudio(path).to_pcm()) needs it
    (wav.py:580-757)"""
    SUFFIX = "wav"
    NAME = SUFFIX
    PRINTABLE_ASCII = PRINTABLE_ASCII

    def __init__(self, filename):
        self.filename = filename
        r = WaveReader(filename)
        (self.__channels__, self.__sample_rate__, self.__bits_per_sample__, self.__channel_mask__,
         self.__total_frames__) = (r.channels, r.sample_rate, r.bits_per_sample, r.channel_mask, r.total_pcm_frames)
        r.close()

    def lossless(self):
        return True

    def channel_mask(self):
        return self.__channel_mask__

    def to_pcm(self):
        """returns a PCMReader object containing the track's PCM data (wav.py:652-657)"""
        return WaveReader(self.filename)

    @classmethod
    def from_pcm(cls, filename, pcmreader, compression=None, total_pcm_frames=None):
        """writes a new RIFF WAVE file from pcmreader's data (wav.py:660-729), the reference's quirk included:
        the pad byte after the data chunk follows the parity of the PCM *frame* count"""
        import os
        from . import EncodingError, FRAMELIST_SIZE
        try:
            header = wave_header(pcmreader.sample_rate, pcmreader.channels, pcmreader.channel_mask,
                                 pcmreader.bits_per_sample, total_pcm_frames if total_pcm_frames is not None else 0)
        except ValueError as err:
            raise EncodingError(str(err))
        try:
            f = open(filename, "wb")
        except IOError as err:
            raise EncodingError(str(err))
        frames_written = 0
        f.write(header)
        try:
            # transfer_framelist_data(counter, f.write, signed = bits_per_sample > 8, big_endian = False)
            while True:
                frame = pcmreader.read(FRAMELIST_SIZE)
                if len(frame) == 0:
                    break
                frames_written += frame.frames
                f.write(frame.to_bytes(False, pcmreader.bits_per_sample > 8))
            pcmreader.close()
        except (IOError, ValueError) as err:
            f.close()
            os.unlink(filename)
            raise EncodingError(str(err))
        except Exception:
            f.close()
            os.unlink(filename)
            raise
        if frames_written % 2:
            f.write(b"\x00")
        if total_pcm_frames is not None:
            if frames_written != total_pcm_frames:
                f.close()
                os.unlink(filename)
                raise EncodingError("total_pcm_frames mismatch")
        else:
            f.seek(0, 0)
            f.write(wave_header(pcmreader.sample_rate, pcmreader.channels, pcmreader.channel_mask,
                                pcmreader.bits_per_sample, frames_written))
        f.close()
        return WaveAudio(filename)

    def total_frames(self):
        return self.__total_frames__

    def sample_rate(self):
        return self.__sample_rate__

    def channels(self):
        return self.__channels__

    def bits_per_sample(self):
        return self.__bits_per_sample__
