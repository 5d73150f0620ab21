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
sys.path.insert(0, os.path.join(ROOT, "tests"))
import audiotools  # noqa: E402
import b200flac  # noqa: E402
import helpers  # noqa: E402


def main():
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 600.0
    n = int(seconds * 44100)
    pcm = helpers.synth_pcm(1235, 2, 16, n)
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    t0 = time.perf_counter(); hashlib.md5(pcm).digest(); t_md5 = time.perf_counter() - t0
    p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6,
                             adaptive_mid_side=True)
    for name, fn in (
        ("b200flac_encode_file (C, PCM in memory)", lambda path: b200flac.encode_file(path, p, pcm, n)),
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
        print("%-46s %7.1f ms for %.0f s of audio  = %8.1f Msamples/s  (x%.0f real time)" % (
            name, dt * 1e3, seconds, 2 * n / dt / 1e6, seconds / dt))
    print("hashlib.md5 of the same PCM alone: %.1f ms (%.0f MB/s) -- the serial floor of one stream" % (
        t_md5 * 1e3, len(pcm) / t_md5 / 1e6))


if __name__ == "__main__":
    main()
