#!/usr/bin/env python
"""Developer probe: b200flac_encode_files on N three-minute tracks (level 8), PCM in page-locked memory, files on tmpfs.
   B200FLAC_FILES_TRACE=1 python tools/many_files_probe.py [tracks] [threads] [jobs]"""
import ctypes as C
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402


def main():
    tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
    L = b200flac.lib()
    tn, distinct = 7938000, 20
    nbytes = tn * distinct * 4
    h = L.b200flac_host_alloc(nbytes)
    d = L.b200flac_device_alloc(0, nbytes)
    L.b200flac_device_synth_pcm(0, d, 1234, 2, 16, 0, tn * distinct)
    L.b200flac_device_download(0, h, d, nbytes)
    L.b200flac_device_free(0, d)
    p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True,
                             exhaustive_model_search=True)
    tmp = tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    names = [os.path.join(tmp.name, "t%d.flac" % i) for i in range(tracks)]
    c_names = (C.c_char_p * tracks)(*[os.fsencode(x) for x in names])
    c_ptrs = (C.c_void_p * tracks)(*[h + (i % distinct) * tn * 4 for i in range(tracks)])
    c_lens = (C.c_uint64 * tracks)(*([tn] * tracks))
    jobs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    total = 0.0
    for n, k in enumerate([min(64, tracks)] + [tracks] * jobs):
        t0 = time.perf_counter()
        if L.b200flac_encode_files(k, c_names, C.byref(p), 4096, None, c_ptrs, c_lens, 0, threads):
            raise SystemExit(L.b200flac_last_error().decode())
        dt = time.perf_counter() - t0
        print("%d tracks, %d threads: %.3f s = %.1f tracks/s" % (k, threads, dt, k / dt))
        if n:
            total += dt
        for x in names[:k]:
            os.unlink(x)
    if jobs > 2:
        # (BASELINE config 5 as written -- 10,000 tracks -- does not fit the box's tmpfs as one job's output: several jobs
        # back to back, the files deleted in between)
        print("%d jobs of %d tracks: %.3f s of encoding = %.1f tracks/s" % (jobs, tracks, total, jobs * tracks / total))


if __name__ == "__main__":
    main()
