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
CUBE = int(os.environ.get("CP360_PROF_CUBE", "256"))
if CUBE == 224:
    SPECS = [("cubepad", 64, 112, 1), ("cubepad", 128, 56, 1), ("cubepad", 64, 56, 1), ("cubepad", 3, 224, 3), ("e2c", 224),
             ("c2emax", 7, 1000), ("cubepad", 256, 28, 1), ("cubepad", 128, 28, 1), ("cubepad", 256, 14, 1),
             ("cubepad", 512, 14, 1), ("cubepad", 2048, 7, 1), ("cubepad", 512, 7, 1), ("cubepad", 2000, 7, 1),
             ("cubepad", 4000, 7, 1)]
elif CUBE == 0:        # the rows of SURVEY.md section 8(f): backward passes, bicubic, fused producers, ConvLSTM widths
    SPECS = [("cubepadbwd", 256, 32, 1), ("cubepadbwd", 512, 16, 1), ("cubepadbwd", 4000, 7, 1), ("c2ebwd", 8, 1000),
             ("c2emaxarg", 8, 1000), ("c2ecubic", 8, 1000), ("c2e", 8, 1000), ("e2cpad", 256), ("bnrelu", 64, 128, 1),
             ("bnrelu", 256, 16, 1), ("cubepad", 4096, 8, 1), ("cubepad", 8192, 8, 1)]
else:
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
    elif spec[0] in ("e2c", "e2cpad"):
        e2c = cp360_b200.Equi2Cube(spec[1], np.empty((960, 1920, 3), np.float32))
        fr = torch.rand(B, 960, 1920, 3, device=dev)
        captured((lambda: e2c.to_cube_tensor(fr)) if spec[0] == "e2c" else (lambda: e2c.to_padded_cube_tensor(fr, 3)))
        del fr
    elif spec[0] == "cubepadbwd":
        _, C, H, p = spec
        gy = torch.randn(6 * B, C, H + 2 * p, H + 2 * p, device=dev)
        captured(lambda: cp360_b200.cube_pad.cubepad_backward(gy, (p, p, p, p), (H, H)))
        del gy
    elif spec[0] == "bnrelu":
        _, C, H, p = spec
        x = torch.randn(6 * B, C, H, H, device=dev)
        sc, sh = torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev)
        with torch.no_grad():
            captured(lambda: cp360_b200.cubepad_fused(x, (p, p, p, p), scale=sc, shift=sh, relu=True))
        del x
    elif spec[0] in ("c2ebwd", "c2ecubic", "c2e", "c2emaxarg"):
        c2e = cp360_b200.Cube2Equi(spec[1])
        w, C = spec[1], spec[2]
        if spec[0] == "c2ebwd":
            t = torch.randn(B, C, 2 * w, 4 * w, device=dev)
            captured(lambda: c2e._backward(t))
        else:
            t = torch.randn(6 * B, C, w, w, device=dev)
            fn = {"c2ecubic": c2e.to_equi_cv2, "c2e": c2e.to_equi_nn, "c2emaxarg": c2e.to_equi_max_with_indices}[spec[0]]
            captured(lambda: fn(t))
        del t
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
