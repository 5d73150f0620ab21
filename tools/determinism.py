#!/usr/bin/env python
"""Encodes the bench hour repeatedly and compares the SHA-256 of all frame bytes (a data race in a
kernel would show up as a run that differs).   python tools/determinism.py [runs]"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402


def main():
    runs = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    L = b200flac.lib()
    n = 158760000
    for name, kw in (("-M lpc 12", dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, adaptive_mid_side=True)),
                     ("-m -e lpc 12", dict(block_size=4096, max_lpc_order=12, max_residual_partition_order=6, mid_side=True,
                                           exhaustive_model_search=True))):
        nn = n if "-e" not in name else n // 6
        p = b200flac.make_params(44100, 2, 16, **kw)
        enc = b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=nn, n_slots=1)
        cap = enc.output_bound(nn, 1)
        d_pcm = L.b200flac_device_alloc(0, nn * 4)
        d_out = L.b200flac_device_alloc(0, cap)
        L.b200flac_device_synth_pcm(0, d_pcm, 1235, 2, 16, 0, nn)
        digests = set()
        for _ in range(runs):
            nbytes, nfr, ms = enc.encode_device(d_pcm, [(0, nn, 0)], d_out, cap)
            host = np.empty(nbytes, dtype=np.uint8)
            L.b200flac_device_download(0, host.ctypes.data, d_out, nbytes)
            digests.add(hashlib.sha256(host).hexdigest())
        print("%-14s %d runs of %d frames: %d distinct output(s) %s" % (name, runs, nfr, len(digests), sorted(digests)[0][:16]))
        assert len(digests) == 1
        L.b200flac_device_free(0, d_pcm); L.b200flac_device_free(0, d_out); enc.close()


if __name__ == "__main__":
    main()
