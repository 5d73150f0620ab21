"""audiotools.decoders -- FlacDecoder on the B200 engine's frame-parallel decoder.

Reference: the C type decoders.FlacDecoder, src/decoders/flac.c:28-500 (init :28-99, read :174-286,
offsets :365-443, close :101-110): built from an open file, it has the PCMReader attributes and hands
out one FLAC frame's PCM per read() as a pcm.FrameList, ends the stream with an empty FrameList once
STREAMINFO's total is reached, and raises ValueError("MD5 mismatch at end of stream") there if the
decoded PCM does not hash to STREAMINFO's MD5.

Here the whole stream is decoded by the GPU on the first read() (b200flac_decode_memory: every frame
at once) and read() then serves the frames from that PCM.  One consequence: a damaged frame raises
the reference's error on the first read() instead of after the frames that precede it.
seek() (FlacDecoder_seek, flac.c:288-356) follows the file's SEEKTABLE exactly as the reference does: the
latest seek point at or before the requested PCM frame (the stream's start without a table), MD5 validation
only when decoding restarts from frame 0.
There is no CPU fallback.
"""
import hashlib
import os
import sys

from . import pcm

try:
    import b200flac
except ImportError:  # the binding lives beside the package
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import b200flac


class FlacDecoder(object):
    def __init__(self, file):
        self.file = file
        self.__data__ = file.read()
        info = b200flac.read_streaminfo(self.__data__)   # ValueError("not a FLAC file") / IOError as flac.c:569-708
        self.__info__ = info
        self.sample_rate = info.sample_rate
        self.channels = info.channels
        self.bits_per_sample = info.bits_per_sample
        self.channel_mask = info.channel_mask
        self.__pcm__ = None
        self.__frames__ = None
        self.__next__ = 0
        self.__pos__ = 0
        self.__finalized__ = False
        self.__closed__ = False
        self.__validate__ = True
        self.__seektable__ = self.__read_seektable__()

    def __read_seektable__(self):
        """[(sample number, byte offset from the first frame), ...] of the SEEKTABLE block, if any
        (flacdec_read_metadata, flac.c:660-676)"""
        data, pos, points = self.__data__, 4, []
        while pos + 4 <= len(data):
            last, kind = data[pos] >> 7, data[pos] & 0x7F
            size = int.from_bytes(data[pos + 1:pos + 4], "big")
            if kind == 3:
                for i in range(size // 18):
                    p = pos + 4 + 18 * i
                    points.append((int.from_bytes(data[p:p + 8], "big"), int.from_bytes(data[p + 8:p + 16], "big")))
            pos += 4 + size
            if last:
                break
        return points

    def seek(self, pcm_frames_offset):
        """positions the stream at the latest seek point at or before pcm_frames_offset and returns that point's
        PCM frame number (FlacDecoder_seek, flac.c:288-356)"""
        if self.__closed__:
            raise ValueError("cannot seek closed stream")
        if pcm_frames_offset < 0:
            raise ValueError("cannot seek to negative value")
        self.__finalized__ = False
        sample, byte_offset = 0, 0
        for s, b in self.__seektable__:
            if s <= pcm_frames_offset:
                sample, byte_offset = s, b
            else:
                break
        if self.__pcm__ is None:
            self.__decode__()
        # the frame that starts at the seek point's byte offset
        index, pos = 0, 0
        for i, (off, n) in enumerate(self.__frames__):
            if off >= byte_offset:
                index = i
                break
            pos += n
        else:
            index = len(self.__frames__)
        self.__next__ = index
        self.__pos__ = pos * self.channels * (self.bits_per_sample // 8)
        self.__validate__ = sample == 0        # flac.c:345-352
        return sample

    def __decode__(self):
        info, pcm_bytes, frames, _ = b200flac.decode(self.__data__, check_md5=False, want_frames=True)
        self.__pcm__, self.__frames__ = memoryview(pcm_bytes), frames

    def read(self, pcm_frames):
        """the next FLAC frame as a FrameList (the argument is ignored, as in FlacDecoder_read)"""
        if self.__closed__:
            raise ValueError("cannot read closed stream")
        empty = pcm.FrameList(b"", self.channels, self.bits_per_sample, False, True)
        if self.__finalized__:
            return empty
        if self.__pcm__ is None:
            self.__decode__()
        if self.__next__ >= len(self.__frames__):
            self.__finalized__ = True
            # FlacDecoder_verify_okay, flac.c:479-490: a blank STREAMINFO MD5 always passes
            md5 = bytes(self.__info__.md5)
            if self.__validate__ and md5 != bytes(16) and hashlib.md5(self.__pcm__).digest() != md5:
                raise ValueError("MD5 mismatch at end of stream")
            return empty
        n = self.__frames__[self.__next__][1]
        size = n * self.channels * (self.bits_per_sample // 8)
        chunk = self.__pcm__[self.__pos__:self.__pos__ + size]
        self.__next__ += 1
        self.__pos__ += size
        return pcm.FrameList(bytes(chunk), self.channels, self.bits_per_sample, False, True)

    def offsets(self):
        """[(byte offset of the frame from the first frame, PCM frames in it), ...] (flac.c:365-443)"""
        if self.__closed__:
            raise ValueError("cannot read closed stream")
        if self.__pcm__ is None:
            self.__decode__()
        return list(self.__frames__)

    def close(self):
        self.__closed__ = True
        self.file.close()
