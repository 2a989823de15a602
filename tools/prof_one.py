#!/usr/bin/env python
"""Run one kernel a few times (for ncu): python tools/prof_one.py cubepad C H p [algo] [B]
                                          python tools/prof_one.py e2c w [B] | c2e w C [B] | c2emax w C [B]
                                          python tools/prof_one.py cubepadbwd C H p [B] | c2ebwd w C [B] | c2ecubic w C [B]
The last call sits inside cudaProfilerStart/Stop: run ncu with --profile-from-start off to capture
exactly that launch (the warm-up calls include the first-call autotuning of CubePad)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cp360_b200

kind = sys.argv[1]
dev = torch.device("cuda", 0)
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if os.environ.get("CP360_PROF_FLUSH") else None


def flush():
    if FLUSH is not None:
        FLUSH.zero_()


if kind == "cubepad":
    C, H, p = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    algo = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    B = int(sys.argv[6]) if len(sys.argv) > 6 else 16
    x = torch.randn(6 * B, C, H, H, device=dev)
    for _ in range(3 if FLUSH is None else 5):
        flush()
        y = cp360_b200.cubepad_forward(x, (p, p, p, p), algo=algo)
    torch.cuda.synchronize()
    flush()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    y = cp360_b200.cubepad_forward(x, (p, p, p, p), algo=algo)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
elif kind == "cubepadbwd":
    C, H, p = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    B = int(sys.argv[5]) if len(sys.argv) > 5 else 16
    gy = torch.randn(6 * B, C, H + 2 * p, H + 2 * p, device=dev)
    for _ in range(3):
        cp360_b200.cube_pad.cubepad_backward(gy, (p, p, p, p), (H, H))
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    cp360_b200.cube_pad.cubepad_backward(gy, (p, p, p, p), (H, H))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
elif kind in ("c2ebwd", "c2ecubic"):
    w, C = int(sys.argv[2]), int(sys.argv[3]); B = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    c2e = cp360_b200.Cube2Equi(w)
    t = torch.randn(B, C, 2 * w, 4 * w, device=dev) if kind == "c2ebwd" else torch.randn(6 * B, C, w, w, device=dev)
    fn = c2e._backward if kind == "c2ebwd" else c2e.to_equi_cv2
    for _ in range(3):
        fn(t)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn(t)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
elif kind == "e2c":
    w = int(sys.argv[2]); B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    e2c = cp360_b200.Equi2Cube(w, np.empty((960, 1920, 3), np.float32))
    fr = torch.rand(B, 960, 1920, 3, device=dev)
    for _ in range(3):
        e2c.to_cube_tensor(fr)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    e2c.to_cube_tensor(fr)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    w, C = int(sys.argv[2]), int(sys.argv[3]); B = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    c2e = cp360_b200.Cube2Equi(w)
    x = torch.randn(6 * B, C, w, w, device=dev)
    for _ in range(3):
        (c2e.to_equi_max if kind == "c2emax" else c2e.to_equi_nn)(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    (c2e.to_equi_max if kind == "c2emax" else c2e.to_equi_nn)(x)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
torch.cuda.synchronize()
