#!/usr/bin/env python
"""c2e backward (cp360_c2e_bwd) over batch sizes / channel counts, cold L2 (kernel time = prologue + groups)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cp360_b200  # noqa: E402

dev = torch.device("cuda", 0)
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for w in (8, 7, 16, 14):
    c2e = cp360_b200.Cube2Equi(w)
    for B, C in ((16, 16), (16, 1000), (32, 1000), (64, 1000), (1, 1000), (32, 2048)):
        if w > 8 and C > 1000:
            continue
        g = torch.randn(B, C, 2 * w, 4 * w, device=dev)
        us = timeit(lambda: c2e._backward(g))
        nbytes = B * C * 14 * w * w * 4
        print("c2e bwd w=%-2d B=%-3d C=%-5d %8.2f MB %8.1f us %8.1f GB/s" % (w, B, C, nbytes / 1e6, us, nbytes / us / 1e3), flush=True)
