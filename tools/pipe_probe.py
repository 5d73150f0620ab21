#!/usr/bin/env python
"""Developer probe: the chunk pipeline (b200flac_encoder_set_chunking) on the bench hour, device resident.
For every (chunk_frames, lookahead): ms per hour (CUDA events inside the library), and the sha256 of the
frames left in HBM -- which must not depend on the setting.  python tools/pipe_probe.py [seconds]"""
import ctypes as C
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-audio-tools_b200"))
import b200flac  # noqa: E402


def main():
    import numpy as np
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3600.0
    L = b200flac.lib()
    n = int(seconds * 44100)
    p = b200flac.make_params(44100, 2, 16, block_size=4096, max_lpc_order=12, max_residual_partition_order=6,
                             adaptive_mid_side=True)
    enc = b200flac.Encoder(p, device=0, max_pcm_frames_per_batch=n, n_slots=1)
    cap = enc.output_bound(n, 1)
    d_pcm = L.b200flac_device_alloc(0, n * 4)
    d_out = L.b200flac_device_alloc(0, cap)
    L.b200flac_device_synth_pcm(0, d_pcm, 1235, 2, 16, 0, n)
    ref = None
    # (chunk_frames, lookahead, model streams, high priority, CTA cap per model launch, lag split G)
    settings = [(0, 3, 1, 1, 0, 0)]
    for chunk in (2048, 4096):
        for nh in (2, 4, 8):
            for prio in (1, 0):
                for gc in (0, 296, 148):
                    settings.append((chunk, nh + 1, nh, prio, gc, 0))
    settings += [(2048, 5, 4, 1, 0, 2), (2048, 9, 8, 1, 0, 2), (1024, 9, 8, 1, 0, 2), (1024, 9, 8, 1, 0, 1),
                 (2048, 5, 4, 0, 0, 2), (8192, 3, 2, 1, 0, 1), (8192, 3, 2, 0, 0, 1), (8192, 3, 2, 1, 592, 1)]
    if len(sys.argv) > 2:
        settings = [(0, 3, 1, 1, 0, 0)] + [tuple(int(x) for x in a.split(":")) for a in sys.argv[2:]]
    for chunk, la, nh, prio, cap_, g in settings:
        os.environ["B200FLAC_NH"] = str(nh)
        os.environ["B200FLAC_LPC_PRIO"] = str(prio)
        os.environ["B200FLAC_LPC_GRID"] = str(cap_)
        if g:
            os.environ["B200FLAC_LPC_G"] = str(g)
        else:
            os.environ.pop("B200FLAC_LPC_G", None)
        enc.set_chunking(chunk, la)
        for _ in range(3):
            nb, nf, ms = enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
        host = np.empty(nb, dtype=np.uint8)
        L.b200flac_device_download(0, host.ctypes.data, d_out, nb)
        h = hashlib.sha256(host.tobytes()).hexdigest()[:16]
        if ref is None:
            ref = h
        K = 5
        tot = 0.0
        t0 = time.perf_counter()
        for _ in range(K):
            nb, nf, ms = enc.encode_device(d_pcm, [(0, n, 0)], d_out, cap)
            tot += ms
        wall = (time.perf_counter() - t0) / K * 1e3
        print("chunk %5d la %d nh %d prio %d cap %3d G %d: device %.3f ms  wall %.3f ms  %.1f Gsamples/s  sha %s %s  kernels %s" % (
            chunk, la, nh, prio, cap_, g, tot / K, wall, n * 2 / (wall * 1e-3) / 1e9, h, "OK" if h == ref else "DIFFERS",
            " ".join("%.2f" % v for v in enc.kernel_ms(0))), flush=True)
    enc.close()


if __name__ == "__main__":
    main()
