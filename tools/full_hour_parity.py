#!/usr/bin/env python
"""One-off: the WHOLE bench hour through the stream layer against the compiled reference encoder
(oracle/_ref/flacenc, single process, ~50 s) -- file bytes must be identical.  Developer tool; the test
suite does the same on 10 minutes."""
import hashlib
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402


def main():
    L = b200flac.lib()
    n = 158760000
    p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)
    d_pcm = L.b200flac_device_alloc(0, n * 4)
    L.b200flac_device_synth_pcm(0, d_pcm, 1235, 2, 16, 0, n)
    host = np.empty(n * 4, dtype=np.uint8)
    L.b200flac_device_download(0, host.ctypes.data, d_pcm, n * 4)
    d = tempfile.mkdtemp(dir="/dev/shm")
    mine, ref, raw = os.path.join(d, "b200.flac"), os.path.join(d, "ref.flac"), os.path.join(d, "in.pcm")
    t0 = time.perf_counter()
    b200flac.encode_file(mine, p, host, n)
    t1 = time.perf_counter()
    host.tofile(raw)
    t2 = time.perf_counter()
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", "flacenc"), "-c", "2", "-r", "44100", "-b", "16", "-B", "4096", "-l", "12",
                    "-R", "6", "-M", ref], stdin=open(raw, "rb"), check=True)
    t3 = time.perf_counter()
    a = hashlib.sha256(open(mine, "rb").read()).hexdigest()
    b = hashlib.sha256(open(ref, "rb").read()).hexdigest()
    print("B200 stream layer: %.2f s, reference encoder: %.1f s, file %d bytes" % (t1 - t0, t3 - t2, os.path.getsize(mine)))
    print("sha256 b200 %s\nsha256 ref  %s\n%s" % (a, b, "IDENTICAL" if a == b else "DIFFERENT"))
    assert a == b


if __name__ == "__main__":
    main()
