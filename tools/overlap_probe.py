#!/usr/bin/env python
"""Developer probe: does running the hour as P concurrent parts (separate streams) beat one batch?
The FP64-bound autocorrelation of one part can then overlap the integer kernels of another."""
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402

L = b200flac.lib()
N = 158760000
p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)
d_pcm = L.b200flac_device_alloc(0, N * 4)
L.b200flac_device_synth_pcm(0, d_pcm, 1235, 2, 16, 0, N)
for parts in (1, 2, 3, 4, 6, 8):
    blocks = (N + 4095) // 4096
    per = (blocks + parts - 1) // parts
    segs = []
    for k in range(parts):
        f0 = k * per * 4096
        n = min(per * 4096, N - f0)
        if n > 0:
            segs.append((f0, n))
    encs = [b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=n, n_slots=1) for _, n in segs]
    caps = [e.output_bound(n, 1) for e, (_, n) in zip(encs, segs)]
    outs = [L.b200flac_device_alloc(0, c) for c in caps]

    def work(i):
        f0, n = segs[i]
        encs[i].encode_device(d_pcm + f0 * 4, [(0, n, f0 // 4096)], outs[i], caps[i])

    def step():
        th = [threading.Thread(target=work, args=(i,)) for i in range(len(segs))]
        for t in th:
            t.start()
        for t in th:
            t.join()
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    K = 5
    for _ in range(K):
        step()
    dt = (time.perf_counter() - t0) / K
    print("parts %d: %.3f ms per hour  (%.1f Gsamples/s)" % (parts, dt * 1e3, N * 2 / dt / 1e9))
    for o in outs:
        L.b200flac_device_free(0, o)
    for e in encs:
        e.close()
