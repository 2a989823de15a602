#!/usr/bin/env python
"""CubePad backward over batch sizes (cold L2): the intercept of time vs. chunks is the per-CTA prologue."""
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


shapes = [(256, 32), (512, 16), (2000, 7)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for C, H in shapes:
    for B in (1, 2, 4, 8, 16, 32):
        gy = torch.randn(6 * B, C, H + 2, H + 2, device=dev)
        us = timeit(lambda: cp360_b200.cube_pad.cubepad_backward(gy, (1, 1, 1, 1), (H, H)))
        nbytes = 6 * B * C * (H * H + (H + 2) ** 2) * 4
        print("cubepad bwd [%d,%d,%d,%d] %8.2f MB %8.1f us %8.1f GB/s" % (6 * B, C, H, H, nbytes / 1e6, us, nbytes / us / 1e3), flush=True)
        del gy
