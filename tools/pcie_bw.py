#!/usr/bin/env python
"""Pinned host <-> device copy bandwidth of the box (developer tool; explains the e2e ceiling)."""
import torch

def main():
    n = 512 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def t(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms = t(lambda: d.copy_(h, non_blocking=True)); print("H2D  %.1f GB/s" % (n / ms / 1e6))
    ms = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H  %.1f GB/s" % (n / ms / 1e6))
    def both():
        with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    import time
    both(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): both()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 5
    print("both directions at once: %.1f GB/s each (%.2f ms for %d MiB each way)" % (n / ms / 1e6, ms, n >> 20))

if __name__ == "__main__":
    main()
