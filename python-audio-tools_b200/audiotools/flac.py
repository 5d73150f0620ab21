"""audiotools.flac -- FlacAudio.from_pcm and the metadata finalisation that follows it.

Python 3 restatement of the caller the B200 engine must be a drop-in for
(reference: audiotools/flac.py:1695-1845 from_pcm, 1847-1876 seektable,
1369-1462 update_metadata, 53-75 block ordering).  from_pcm's contract with the
encoder is unchanged: it passes the same keyword options per compression level,
wraps the reader in BufferedPCMReader, and turns the returned
(byte offset, PCM frames) list into a SEEKTABLE written into the PADDING block.
"""
import os
import struct
from bisect import bisect_right

BLOCK_STREAMINFO, BLOCK_PADDING, BLOCK_SEEKTABLE, BLOCK_VORBIS_COMMENT = 0, 1, 3, 4
# FlacMetaData.add_block's preferred order (flac.py:59-65)
PREFERRED_ORDER = [0, 3, 5, 4, 6, 2, 1]


def _native_finalize(filename, offsets, seekpoint_interval, channel_mask):
    from . import encoders
    encoders.finalize_flac_metadata(filename, offsets, seekpoint_interval, channel_mask)


def skip_id3v2_comment(f):
    """seeks past the ID3v2 tags at the start of the stream, if any; returns the bytes skipped
    (audiotools/id3.py:264-311)"""
    start = f.tell()
    head = f.read(10)
    if len(head) < 10 or head[0:3] != b"ID3" or head[3] not in (2, 3, 4) or any(b & 0x80 for b in head[6:10]):
        f.seek(start)
        return 0
    tag_size = (head[6] << 21) | (head[7] << 14) | (head[8] << 7) | head[9]     # decode_syncsafe32, id3.py:81-104
    f.seek(tag_size, 1)
    return 10 + tag_size + skip_id3v2_comment(f)


class FlacMetaData(object):
    """ordered list of (block_id, payload) metadata blocks"""

    def __init__(self, blocks):
        self.block_list = list(blocks)

    def get_blocks(self, block_id):
        return [b for b in self.block_list if b[0] == block_id]

    def add_block(self, block):
        stop = set(PREFERRED_ORDER[PREFERRED_ORDER.index(block[0]) + 1:])
        for i, old in enumerate(self.block_list):
            if old[0] in stop:
                self.block_list.insert(i, block)
                break
        else:
            self.block_list.append(block)

    def size(self):
        return sum(4 + len(b[1]) for b in self.block_list)

    def build(self):
        out = []
        for i, (bid, payload) in enumerate(self.block_list):
            last = 0x80 if i == len(self.block_list) - 1 else 0
            out.append(bytes([last | bid]) + len(payload).to_bytes(3, "big") + payload)
        return b"".join(out)


class FlacAudio(object):
    """a FLAC file on disk (the slice of audiotools.flac.FlacAudio used by from_pcm)"""

    SUFFIX = "flac"
    NAME = SUFFIX
    # True: SEEKTABLE / channel-mask tag / PADDING adjustment by the engine library (C); False: the
    # Python restatement of flac.py:1811-1832 below (kept as the checker of the native path)
    NATIVE_FINALIZE = True
    DEFAULT_COMPRESSION = "8"
    COMPRESSION_MODES = tuple(map(str, range(0, 9)))

    # compression level -> encode_flac options, flac.py:1719-1764
    ENCODING_OPTIONS = {
        "0": {"block_size": 1152, "max_lpc_order": 0, "min_residual_partition_order": 0,
              "max_residual_partition_order": 3},
        "1": {"block_size": 1152, "max_lpc_order": 0, "adaptive_mid_side": True,
              "min_residual_partition_order": 0, "max_residual_partition_order": 3},
        "2": {"block_size": 1152, "max_lpc_order": 0, "exhaustive_model_search": True,
              "min_residual_partition_order": 0, "max_residual_partition_order": 3},
        "3": {"block_size": 4096, "max_lpc_order": 6, "min_residual_partition_order": 0,
              "max_residual_partition_order": 4},
        "4": {"block_size": 4096, "max_lpc_order": 8, "adaptive_mid_side": True,
              "min_residual_partition_order": 0, "max_residual_partition_order": 4},
        "5": {"block_size": 4096, "max_lpc_order": 8, "mid_side": True,
              "min_residual_partition_order": 0, "max_residual_partition_order": 5},
        "6": {"block_size": 4096, "max_lpc_order": 8, "mid_side": True,
              "min_residual_partition_order": 0, "max_residual_partition_order": 6},
        "7": {"block_size": 4096, "max_lpc_order": 8, "mid_side": True, "exhaustive_model_search": True,
              "min_residual_partition_order": 0, "max_residual_partition_order": 6},
        "8": {"block_size": 4096, "max_lpc_order": 12, "mid_side": True, "exhaustive_model_search": True,
              "min_residual_partition_order": 0, "max_residual_partition_order": 6}}

    def __init__(self, filename):
        from . import InvalidFLAC
        self.filename = filename
        self.__samplerate__ = 0
        self.__channels__ = 0
        self.__bitspersample__ = 0
        self.__total_frames__ = 0
        self.__stream_offset__ = 0
        self.__md5__ = bytes(16)
        try:
            self.__read_streaminfo__()
        except IOError as msg:
            raise InvalidFLAC(str(msg))

    def __read_streaminfo__(self):
        """flac.py:2420-2462: skips ID3v2 tags in front of the stream (__stream_offset__), then walks the
        metadata blocks until STREAMINFO -- which need not be the first one"""
        from . import InvalidFLAC
        with open(self.filename, "rb") as f:
            self.__stream_offset__ = skip_id3v2_comment(f)
            f.read(4)
            while True:
                hdr = f.read(4)
                if len(hdr) < 4:
                    raise IOError("I/O error reading stream")
                stop, header_type, length = hdr[0] >> 7, hdr[0] & 0x7F, int.from_bytes(hdr[1:4], "big")
                if header_type > 6:
                    raise InvalidFLAC("invalid metadata block type")       # ERR_FLAC_INVALID_BLOCK, text.py:576
                if header_type == BLOCK_STREAMINFO:
                    si = f.read(34)
                    if len(si) < 34:
                        raise IOError("I/O error reading stream")
                    v = int.from_bytes(si[10:18], "big")
                    self.__samplerate__ = v >> 44
                    self.__channels__ = ((v >> 41) & 7) + 1
                    self.__bitspersample__ = ((v >> 36) & 31) + 1
                    self.__total_frames__ = v & ((1 << 36) - 1)
                    self.__md5__ = bytes(si[18:34])
                    break
                f.seek(length, 1)
                if stop:
                    break

    def sample_rate(self):
        return self.__samplerate__

    def channels(self):
        return self.__channels__

    def bits_per_sample(self):
        return self.__bitspersample__

    def total_frames(self):
        return self.__total_frames__

    def to_pcm(self):
        """returns a PCMReader object containing the track's PCM data (flac.py:1674-1693)"""
        from . import decoders
        flac = open(self.filename, "rb")
        if self.__stream_offset__ > 0:
            flac.seek(self.__stream_offset__)
        return decoders.FlacDecoder(flac)

    def verify(self, progress=None):
        """decodes the whole file, checking every frame CRC-16 and the STREAMINFO MD5 (AudioFile.verify reads
        to_pcm() to its end, audiotools/__init__.py); on the engine that is one call, b200flac_verify_file.
        Returns True or raises InvalidFLAC with the decoder's message"""
        from . import InvalidFLAC
        from . import decoders
        try:
            return decoders.b200flac.verify_file(self.filename)
        except (IOError, ValueError) as err:
            raise InvalidFLAC(str(err))

    def get_metadata(self):
        blocks = []
        with open(self.filename, "rb") as f:
            f.seek(self.__stream_offset__ + 4)
            while True:
                hdr = f.read(4)
                length = int.from_bytes(hdr[1:4], "big")
                blocks.append((hdr[0] & 0x7F, f.read(length)))
                if hdr[0] & 0x80:
                    break
        return FlacMetaData(blocks)

    def metadata_length(self):
        return self.get_metadata().size()

    def update_metadata(self, metadata):
        """flac.py:1369-1462: shrink/grow PADDING so the frames do not move when possible,
        else rewrite the file"""
        if metadata is None:
            return
        paddings = [i for i, b in enumerate(metadata.block_list) if b[0] == BLOCK_PADDING]
        total_padding = sum(len(metadata.block_list[i][1]) for i in paddings)
        delta = metadata.size() - self.metadata_length()
        if paddings and delta <= total_padding:
            for i in paddings:
                length = len(metadata.block_list[i][1])
                if delta > 0:
                    take = min(delta, length)
                    length -= take
                    delta -= take
                elif delta < 0:
                    length -= delta
                    delta = 0
                else:
                    break
                metadata.block_list[i] = (BLOCK_PADDING, b"\x00" * length)
            with open(self.filename, "r+b") as f:
                f.seek(self.__stream_offset__)
                f.write(b"fLaC" + metadata.build())
        else:
            with open(self.filename, "rb") as f:
                prefix = f.read(self.__stream_offset__)
                f.seek(self.__stream_offset__ + 4 + self.metadata_length())
                frames = f.read()
            tmp = self.filename + ".tmp"
            with open(tmp, "wb") as f:
                f.write(prefix + b"fLaC" + metadata.build() + frames)
            os.replace(tmp, self.filename)

    def seektable(self, offsets, seekpoint_interval=None):
        """SEEKTABLE block from the encoder's (byte offset, PCM frames) list, flac.py:1847-1876"""
        if seekpoint_interval is None:
            seekpoint_interval = self.sample_rate() * 10
        total = 0
        all_frames = {}
        sample_offsets = []
        for byte_offset, pcm_frames in offsets:
            all_frames[total] = (byte_offset, pcm_frames)
            sample_offsets.append(total)
            total += pcm_frames
        points = []
        for pcm_frame in range(0, self.total_frames(), seekpoint_interval):
            i = bisect_right(sample_offsets, pcm_frame) - 1
            so = sample_offsets[i]
            points.append((so, all_frames[so][0], all_frames[so][1]))
        return (BLOCK_SEEKTABLE, b"".join(struct.pack(">QQH", a, b, c) for a, b, c in points))

    @classmethod
    def from_pcm(cls, filename, pcmreader, compression=None, total_pcm_frames=None, encoding_function=None):
        """encodes a new file from PCM data (flac.py:1695-1845); the encoder call is unchanged"""
        from .encoders import encode_flac
        from . import (EncodingError, UnsupportedChannelCount, UnsupportedChannelMask, BufferedPCMReader,
                       __default_quality__)

        if compression is None or compression not in cls.COMPRESSION_MODES:
            compression = __default_quality__(cls.NAME)
        encoding_options = cls.ENCODING_OPTIONS[compression]

        if pcmreader.channels > 8:
            raise UnsupportedChannelCount(filename, pcmreader.channels)
        if int(pcmreader.channel_mask) == 0:
            if pcmreader.channels <= 6:
                channel_mask = {1: 0x0004, 2: 0x0003, 3: 0x0007, 4: 0x0033, 5: 0x0037, 6: 0x003F}[pcmreader.channels]
            else:
                channel_mask = 0
        elif int(pcmreader.channel_mask) not in (0x0001, 0x0004, 0x0003, 0x0007, 0x0033, 0x0603, 0x0037, 0x0607,
                                                 0x003F, 0x060F):
            raise UnsupportedChannelMask(filename, int(pcmreader.channel_mask))
        else:
            channel_mask = int(pcmreader.channel_mask)

        if total_pcm_frames is not None:
            interval = pcmreader.sample_rate * 10
            expected_seekpoints = (total_pcm_frames // interval) + (1 if (total_pcm_frames % interval) else 0)
            padding_size = 4096 + 4 + (expected_seekpoints * 18)
        else:
            padding_size = 4096

        try:
            offsets = (encode_flac if encoding_function is None else encoding_function)(
                filename, pcmreader=BufferedPCMReader(pcmreader), padding_size=padding_size, **encoding_options)
            flac = FlacAudio(filename)
            mask_tag = channel_mask if (((pcmreader.channels > 2) or (pcmreader.bits_per_sample > 16)) and
                                        channel_mask != 0) else 0
            if cls.NATIVE_FINALIZE:
                # the same three steps in C (b200flac_finalize_metadata): no Python pass over the file
                _native_finalize(filename, list(offsets), pcmreader.sample_rate * 10, mask_tag)
                return FlacAudio(filename)
            metadata = flac.get_metadata()
            assert metadata is not None
            metadata.add_block(flac.seektable(list(offsets), pcmreader.sample_rate * 10))
            if ((pcmreader.channels > 2) or (pcmreader.bits_per_sample > 16)) and channel_mask != 0:
                for i, (bid, payload) in enumerate(metadata.block_list):
                    if bid == BLOCK_VORBIS_COMMENT:
                        vlen = int.from_bytes(payload[0:4], "little")
                        vendor = payload[4:4 + vlen]
                        count = int.from_bytes(payload[4 + vlen:8 + vlen], "little")
                        rest = payload[8 + vlen:]
                        comment = ("WAVEFORMATEXTENSIBLE_CHANNEL_MASK=0x%.4X" % channel_mask).encode("utf-8")
                        payload = (vlen.to_bytes(4, "little") + vendor + (count + 1).to_bytes(4, "little") + rest +
                                   len(comment).to_bytes(4, "little") + comment)
                        metadata.block_list[i] = (bid, payload)
                        break
            flac.update_metadata(metadata)
            return flac
        except (IOError, ValueError) as err:
            cls.__unlink__(filename)
            raise EncodingError(str(err))
        except Exception:
            cls.__unlink__(filename)
            raise

    @classmethod
    def __unlink__(cls, filename):
        try:
            os.unlink(filename)
        except OSError:
            pass
