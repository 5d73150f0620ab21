#!/usr/bin/env python
"""Timing tier (A) of SURVEY.md 8(d): through the drop-in API end to end (reader calls, host MD5, file write).
    python tools/tier_a.py [seconds]"""
import hashlib
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import audiotools  # noqa: E402
import b200flac  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    n = int(seconds * 44100)
    pcm = b200flac.synth_pcm(1235, 2, 16, n)
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    t0 = time.perf_counter(); hashlib.md5(pcm).digest(); t_md5 = time.perf_counter() - t0
    p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6,
                             adaptive_mid_side=True)
    # the same PCM as a RIFF WAVE file on tmpfs: the file-backed sources of SURVEY 8f-2
    import struct
    wav = os.path.join(d, "in.wav")
    with open(wav, "wb") as fh:
        fmt = struct.pack("<HHIIHH", 1, 2, 44100, 44100 * 4, 4, 16)
        fh.write(b"RIFF" + struct.pack("<I", 4 + 8 + len(fmt) + 8 + len(pcm)) + b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt +
                 b"data" + struct.pack("<I", len(pcm)))
        fh.write(pcm)

    def from_wave(path, level, feed):
        os.environ["B200FLAC_FILE_FEED"] = feed
        audiotools.FlacAudio.from_pcm(path, audiotools.WaveAudio(wav).to_pcm(), level)

    for name, fn in (
        ("b200flac_encode_file (C, PCM in memory)", lambda path: b200flac.encode_file(path, p, pcm, n)),
        ("b200flac_encode_wave (C, WAVE file -> FLAC file)", lambda path: b200flac.encode_container(path, wav, p, "wave")),
        ("from_pcm(WaveAudio.to_pcm()) level 8, file feed", lambda path: from_wave(path, "8", "1")),
        ("from_pcm(WaveAudio.to_pcm()) level 8, FrameLists", lambda path: from_wave(path, "8", "0")),
        ("from_pcm(WaveAudio.to_pcm()) level 4, file feed", lambda path: from_wave(path, "4", "1")),
        ("from_pcm(WaveAudio.to_pcm()) level 4, FrameLists", lambda path: from_wave(path, "4", "0")),
        ("FlacAudio.from_pcm level 4 (Python reader)", lambda path: audiotools.FlacAudio.from_pcm(
            path, audiotools.PCMBytesReader(pcm, 44100, 2, 0x3, 16), "4", total_pcm_frames=n)),
        ("FlacAudio.from_pcm level 8 (Python reader)", lambda path: audiotools.FlacAudio.from_pcm(
            path, audiotools.PCMBytesReader(pcm, 44100, 2, 0x3, 16), "8", total_pcm_frames=n)),
    ):
        path = os.path.join(d, "a.flac")
        fn(path)
        t0 = time.perf_counter()
        fn(path)
        dt = time.perf_counter() - t0
        print("%-50s %7.1f ms for %.0f s of audio  = %8.1f Msamples/s  (x%.0f real time)" % (
            name, dt * 1e3, seconds, 2 * n / dt / 1e6, seconds / dt))
    print("hashlib.md5 of the same PCM alone: %.1f ms (%.0f MB/s) -- the serial floor of one stream" % (
        t_md5 * 1e3, len(pcm) / t_md5 / 1e6))


if __name__ == "__main__":
    main()
