#!/usr/bin/env python
"""Per-site CubePad timing through the AUTO path (first-call autotuner / tuning table), cold L2, at one
or more batch sizes: GB/s of every distinct cubic-ResNet-50 / ConvLSTM site of a face width.

    python tools/site_sweep.py --cube 224 --batch 32,16,8,4,2,1 [--iters 20]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import cp360_b200  # noqa: E402
from cp360_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cube", default="224")
    ap.add_argument("--batch", default="32")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--extra", default="", help="extra sites CxHxp,...")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = 6546.6
    for cube in [int(v) for v in args.cube.split(",")]:
        sites = list(dict.fromkeys(cp360_b200.resnet50_cubepad_sites(cube)))
        if cube == 224:
            sites += [(2000, 7, 1), (4000, 7, 1)]
        else:
            sites += [(2048, cube // 32, 1), (4096, cube // 32, 1), (8192, cube // 32, 1)]
        for t in [t for t in args.extra.split(",") if t]:
            sites.append(tuple(int(v) for v in t.split("x")))
        for B in [int(v) for v in args.batch.split(",")]:
            tot_us, tot_bytes = 0.0, 0
            mult = {s: cp360_b200.resnet50_cubepad_sites(cube).count(s) or 1 for s in sites}
            for (C, H, p) in sites:
                n = 6 * B
                x = torch.randn(n, C, H, H, device=dev)
                y = torch.empty(n, C, H + 2 * p, H + 2 * p, device=dev)
                st = torch.cuda.current_stream().cuda_stream

                def fn():
                    _lib.check(lib.cp360_cubepad_fwd(x.data_ptr(), y.data_ptr(), n, C, H, H, p, p, p, p, 4, st))
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                evs = []
                for _ in range(args.iters):
                    flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    fn()
                    b.record()
                    evs.append((a, b))
                torch.cuda.synchronize()
                ts = sorted(a.elapsed_time(b) for a, b in evs)
                ms = ts[len(ts) // 2]
                nbytes = n * C * (H * H + (H + 2 * p) ** 2) * 4
                buf = ctypes.create_string_buffer(200)
                lib.cp360_cubepad_tune_info(n, C, H, H, p, p, p, p, buf, 200)
                algo = lib.cp360_cubepad_pick_algo(n, C, H, H, p, p, p, p, 4, 1)
                gbs = nbytes / (ms * 1e-3) / 1e9
                print("cube %3d B %2d site %4dx%3d p%d x%d  %8.2f MB %8.1f us %7.1f GB/s %.2f  algo %d  %s"
                      % (cube, B, C, H, p, mult[(C, H, p)], nbytes / 1e6, ms * 1e3, gbs, gbs / peak, algo,
                         buf.value.decode() or "heuristic"), flush=True)
                if (C, H, p) in cp360_b200.resnet50_cubepad_sites(cube):
                    tot_us += ms * 1e3 * mult[(C, H, p)]
                    tot_bytes += nbytes * mult[(C, H, p)]
                del x, y
            print("cube %3d B %2d 18-site total %.1f us, %.1f GB/s (%.2f of %.1f)"
                  % (cube, B, tot_us, tot_bytes / (tot_us * 1e-6) / 1e9, tot_bytes / (tot_us * 1e-6) / 1e9 / peak, peak),
                  flush=True)


if __name__ == "__main__":
    main()
