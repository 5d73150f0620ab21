#!/usr/bin/env python
"""BASELINE.json config #5 at the API tier: many 3-minute tracks encoded concurrently through
b200flac_encode_file from a pool of host threads (the MD5 of every track is serial, so the host cores
are what scales).   python tools/many_tracks.py [tracks] [threads]"""
import os
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402


def main():
    tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
    n = 7938000                                   # 3 minutes at 44.1 kHz
    pcms = [b200flac.synth_pcm(2000 + i, 2, 16, n) for i in range(4)]
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    for label, kw in (("level 8 (from_pcm default: -m -e, lpc 12)", dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6,
                                                                        mid_side=True, exhaustive_model_search=True)),
                      ("level 5 (-m, lpc 8)", dict(block_size=4096, max_lpc_order=8, max_residual_partition_order=5, mid_side=True))):
        p = b200flac.make_params(44100, 2, 16, **kw)
        nxt = [0]
        lock = threading.Lock()

        def worker(tid):
            while True:
                with lock:
                    i = nxt[0]
                    nxt[0] += 1
                if i >= tracks:
                    return
                b200flac.encode_file(os.path.join(d, "t%d_%d.flac" % (tid, i % 4)), p, pcms[i % 4], n)

        for rep in range(2):                      # the first pass fills the encoder pool
            nxt[0] = 0
            th = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            dt = time.perf_counter() - t0
        print("%-44s %d tracks, %d threads: %.2f s  = %.1f tracks/s = %.0f Msamples/s  (10,000 tracks: %.0f s)" % (
            label, tracks, threads, dt, tracks / dt, tracks * n * 2 / dt / 1e6, 10000 * dt / tracks))


if __name__ == "__main__":
    main()
