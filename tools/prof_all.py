#!/usr/bin/env python
"""Every distinct kernel/shape of the benchmark chain in ONE process, for one ncu invocation:

    ncu --set full --clock-control none --profile-from-start off -k regex:'cubepad|e2c_kernel|c2e_' -o full_all \
        python tools/prof_all.py <frames per launch> <order file>

Each spec is warmed (the first CubePad call of a shape runs the autotuner), then launched once inside
cudaProfilerStart/Stop, so the report holds exactly one record per spec, in the order written to <order file>
(read by tools/make_profiles.py to name the records)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import cp360_b200

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
order_path = sys.argv[2] if len(sys.argv) > 2 else ""
dev = torch.device("cuda", 0)
SPECS = [("cubepad", 64, 128, 1), ("cubepad", 128, 64, 1), ("cubepad", 64, 64, 1), ("cubepad", 3, 256, 3), ("e2c", 256),
         ("c2emax", 8, 1000), ("cubepad", 256, 32, 1), ("cubepad", 128, 32, 1), ("cubepad", 256, 16, 1),
         ("cubepad", 512, 16, 1), ("cubepad", 2048, 8, 1), ("cubepad", 512, 8, 1)]
only = os.environ.get("CP360_PROF_ONLY", "")
if only:
    SPECS = [s for s in SPECS if "_".join(str(v) for v in s) in only.split(",")]


def captured(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


names = []
for spec in SPECS:
    if spec[0] == "cubepad":
        _, C, H, p = spec
        x = torch.randn(6 * B, C, H, H, device=dev)
        captured(lambda: cp360_b200.cubepad_forward(x, (p, p, p, p)))
        del x
    elif spec[0] == "e2c":
        e2c = cp360_b200.Equi2Cube(spec[1], np.empty((960, 1920, 3), np.float32))
        fr = torch.rand(B, 960, 1920, 3, device=dev)
        captured(lambda: e2c.to_cube_tensor(fr))
        del fr
    else:
        c2e = cp360_b200.Cube2Equi(spec[1])
        x = torch.randn(6 * B, spec[2], spec[1], spec[1], device=dev)
        captured(lambda: c2e.to_equi_max(x))
        del x
    names.append("_".join(str(v) for v in spec))
    torch.cuda.empty_cache()
if order_path:
    with open(order_path, "w") as f:
        f.write("\n".join(names) + "\n")
print("captured:", " ".join(names))
