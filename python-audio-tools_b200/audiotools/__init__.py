"""audiotools -- the slice of Python Audio Tools that sits on the FLAC-encode hot path,
loadable under Python 3 and backed by the B200 engine.

Only what `FlacAudio.from_pcm` -> `audiotools.encoders.encode_flac` needs is here
(reference: audiotools/__init__.py of widgital/python-audio-tools, Python 2 only):

  BufferedPCMReader      audiotools/__init__.py:2561-2603
  EncodingError & co.    audiotools/__init__.py (exception classes)
  FlacAudio              audiotools/flac.py:1251 (from_pcm, seektable, update_metadata)
  WaveReader/WaveAudio   audiotools/wav.py:421-757, AiffReader/AiffAudio audiotools/aiff.py:350-560
                         (the file-backed PCM sources in front of from_pcm)

`audiotools.pcm` and `audiotools.encoders` are C extension modules (pcm.c, encoders.c);
encoders.encode_flac keeps the reference's signature and returns the same
(byte offset, PCM frames) list.  Everything else in the reference package (other formats,
tagging, CLI tools, player) is out of scope for this repository.
"""
VERSION = "2.22alpha1"  # audiotools/__init__.py:134

FRAMELIST_SIZE = 0x100000 // 4


class EncodingError(IOError):
    """raised when a file cannot be encoded (audiotools/__init__.py)"""

    def __init__(self, error_message):
        IOError.__init__(self)
        self.error_message = error_message

    def __str__(self):
        return self.error_message if isinstance(self.error_message, str) else str(self.error_message)


class UnsupportedChannelCount(EncodingError):
    def __init__(self, filename, count):
        EncodingError.__init__(self, "unable to write \"%s\" with %d channel input" % (filename, count))


class UnsupportedChannelMask(EncodingError):
    def __init__(self, filename, mask):
        EncodingError.__init__(self, "unable to write \"%s\" with channel assignment 0x%X" % (filename, mask))


class InvalidFLAC(ValueError):
    pass


def __default_quality__(audio_type):
    """audiotools/__init__.py:5655; the FLAC default is "8" (flac.py:1260)"""
    return "8"


class BufferedPCMReader(object):
    """a PCMReader which reads exact counts of PCM frames (audiotools/__init__.py:2561-2603)"""

    def __init__(self, pcmreader):
        from . import pcm
        self.pcmreader = pcmreader
        self.sample_rate = pcmreader.sample_rate
        self.channels = pcmreader.channels
        self.channel_mask = pcmreader.channel_mask
        self.bits_per_sample = pcmreader.bits_per_sample
        self.buffer = pcm.from_list([], self.channels, self.bits_per_sample, True)

    def close(self):
        self.pcmreader.close()
        self.read = self.read_closed

    def read(self, pcm_frames):
        """returns exactly pcm_frames frames until the end of the stream, never more"""
        while self.buffer.frames < pcm_frames:
            frame = self.pcmreader.read(FRAMELIST_SIZE)
            if len(frame):
                self.buffer += frame
            else:
                break
        (output, self.buffer) = self.buffer.split(pcm_frames)
        return output

    def read_closed(self, pcm_frames):
        raise ValueError()


class PCMBytesReader(object):
    """a minimal PCMReader over packed signed little-endian PCM in memory
    (stands in for the reference's PCMReader(file, ...) in tests and examples)"""

    def __init__(self, data, sample_rate, channels, channel_mask, bits_per_sample):
        self.data = memoryview(data)
        self.sample_rate = sample_rate
        self.channels = channels
        self.channel_mask = channel_mask
        self.bits_per_sample = bits_per_sample
        self.pos = 0
        self.frame_bytes = channels * (bits_per_sample // 8)

    def read(self, pcm_frames):
        from . import pcm
        n = max(1, pcm_frames) * self.frame_bytes
        chunk = self.data[self.pos:self.pos + n]
        self.pos += len(chunk)
        return pcm.FrameList(bytes(chunk), self.channels, self.bits_per_sample, False, True)

    def close(self):
        pass


from .flac import FlacAudio  # noqa: E402,F401
from . import wav, aiff  # noqa: E402,F401
from .wav import WaveAudio, WaveReader  # noqa: E402,F401
from .aiff import AiffAudio, AiffReader  # noqa: E402,F401
