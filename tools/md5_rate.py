#!/usr/bin/env python
"""Developer probe: what one device thread hashes per second (k_md5_tracks through the test hook, one string)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402

b200flac.device_md5(b"warm")
rng = np.random.RandomState(3)
prev = None
for mb in (8, 16, 32, 64):
    data = rng.randint(0, 256, size=mb << 20).astype(np.uint8)
    t0 = time.perf_counter()
    b200flac.device_md5(data)
    dt = time.perf_counter() - t0
    print("%d MB: %.1f ms" % (mb, dt * 1e3), "" if prev is None else "-> %.1f MB/s for the extra bytes" % ((mb - prev[0]) / (dt - prev[1])))
    prev = (mb, dt)
