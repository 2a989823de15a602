#!/usr/bin/env python
"""Per-kernel micro-benchmark (GPU box): GB/s of every kernel variant at the BASELINE shapes,
beside a plain device copy of comparable size. Development tool; bench.py is the contract.

    python tools/kbench.py [--batch 16] [--iters 20] [--only cubepad|e2c|c2e]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import cp360_b200  # noqa: E402
from cp360_b200 import _lib  # noqa: E402


def timeit(fn, iters, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    B = args.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2 (126 MB)
    rows = []

    def report(name, nbytes, ms):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append({"name": name, "MB": round(nbytes / 1e6, 2), "us": round(ms * 1e3, 1), "GB/s": round(gbs, 1)})
        print("%-58s %9.2f MB %9.1f us %8.1f GB/s" % (name, nbytes / 1e6, ms * 1e3, gbs), flush=True)

    # plain copies for scale
    for mb in (8, 64, 512):
        a = torch.empty(mb << 18, dtype=torch.float32, device=dev)
        b = torch.empty_like(a)
        report("torch copy %d MB" % mb, 2 * a.numel() * 4, timeit(lambda: b.copy_(a), args.iters, flush))

    if args.only in ("", "cubepad"):
        sites = []
        for s in cp360_b200.resnet50_cubepad_sites(256) + [(2048, 8, 1), (64, 256, 2)] + \
                cp360_b200.resnet50_cubepad_sites(224) + [(2000, 7, 1), (4000, 7, 1)]:
            if s not in sites:
                sites.append(s)
        only_sites = os.environ.get("CP360_KB_SITES", "")
        if only_sites:
            keep = {tuple(int(v) for v in t.split("x")) for t in only_sites.split(",")}
            sites = [s for s in sites if (s[0], s[1]) in keep]
        for C, H, p in sites:
            n = 6 * B if (C, H) != (64, 256) else 12
            x = torch.randn(n, C, H, H, device=dev)
            nbytes = n * C * (H * H + (H + 2 * p) ** 2) * 4
            # the practical ceiling at this size: a plain device copy moving the same number of bytes
            ca = torch.empty(nbytes // 8, dtype=torch.float32, device=dev)
            cb = torch.empty_like(ca)
            report("  d2d copy of the same bytes [%d,%d,%d,%d] p%d" % (n, C, H, H, p), nbytes,
                   timeit(lambda: cb.copy_(ca), args.iters, flush))
            del ca, cb
            for algo, an in ((1, "generic"), (4, "cube"), (5, "row"), (6, "cube2")):
                try:
                    cp360_b200.cubepad_forward(x, (p, p, p, p), algo=algo)
                except _lib.CP360Error:
                    continue
                auto = _lib.lib().cp360_cubepad_pick_algo(n, C, H, H, p, p, p, p, 4, 1) == algo
                yoff = int(os.environ.get("CP360_KB_YOFF", "0")) // 4
                ybuf = torch.empty(n * C * (H + 2 * p) * (H + 2 * p) + yoff, device=dev)
                y = ybuf[yoff:].view(n, C, H + 2 * p, H + 2 * p)
                st = torch.cuda.current_stream().cuda_stream

                def fn():
                    _lib.lib().cp360_cubepad_fwd_algo(x.data_ptr(), y.data_ptr(), n, C, H, H, p, p, p, p, 4, algo, st)
                report("cubepad [%d,%d,%d,%d] p%d %s%s" % (n, C, H, H, p, an, " *" if auto else ""), nbytes,
                       timeit(fn, args.iters, flush))
            del x

    if args.only in ("", "fused"):
        # rows (f2) of SURVEY.md section 8: producer ops folded into the pad kernel, next to the unfused sequence
        for C, H in ((64, 128), (256, 32), (512, 16)):
            n = 6 * B
            x = torch.randn(n, C, H, H, device=dev)
            bn = torch.nn.BatchNorm2d(C).to(dev).eval()
            pad = cp360_b200.CubePad(1)
            nbytes = n * C * (H * H + (H + 2) ** 2) * 4
            with torch.no_grad():
                report("bn+relu+CubePad [%d,%d,%d,%d] torch bn, relu + cp360 pad" % (n, C, H, H), nbytes,
                       timeit(lambda: pad(torch.relu(bn(x))), args.iters, flush))
                report("bn+relu+CubePad [%d,%d,%d,%d] fused (cp360_cubepad_fused_fwd)" % (n, C, H, H), nbytes,
                       timeit(lambda: cp360_b200.cubepad_bn_relu(x, bn, 1), args.iters, flush))
                xs = [x, torch.randn(n, C, H, H, device=dev)]
                nb2 = 2 * nbytes
                report("cat+CubePad     2x[%d,%d,%d,%d] torch.cat + cp360 pad" % (n, C, H, H), nb2,
                       timeit(lambda: pad(torch.cat(xs, 1)), args.iters, flush))
                report("cat+CubePad     2x[%d,%d,%d,%d] fused (2 windowed launches)" % (n, C, H, H), nb2,
                       timeit(lambda: cp360_b200.cubepad_cat(xs, 1), args.iters, flush))
            del x, xs

    if args.only in ("", "e2c"):
        pipe = cp360_b200.SphericalPipeline(device=dev)
        for w in (256, 224):
            import numpy as np
            e2c = cp360_b200.Equi2Cube(w, np.empty((960, 1920, 3), np.float32))
            frames = torch.rand(B, 960, 1920, 3, device=dev)
            out = torch.empty(6 * B, 3, w, w, device=dev)
            pipe.e2c, pipe.cube = e2c, w
            nb = pipe.e2c_bytes_per_frame() * B
            report("e2c 1920x960 -> %d NCHW B=%d" % (w, B), nb,
                   timeit(lambda: e2c.to_cube_tensor(frames, out=out), args.iters, flush))
            report("e2c 1920x960 -> %d NCHW+norm B=%d" % (w, B), nb,
                   timeit(lambda: e2c.to_cube_tensor(frames, out=out, mean=[.485, .456, .406], std=[.229, .224, .225]),
                          args.iters, flush))
            # row f2: e2c + norm + CubePad(3) — two kernels vs the fused one (bytes = e2c + CubePad(3) algorithmic)
            pad3 = cp360_b200.CubePad(3)
            padded = torch.empty(6 * B, 3, w + 6, w + 6, device=dev)
            nb2 = nb + 6 * B * 3 * (w * w + (w + 6) ** 2) * 4
            u8 = torch.randint(0, 256, (B, 960, 1920, 3), dtype=torch.uint8, device=dev)
            report("e2c+norm -> CubePad(3) %d two kernels B=%d" % (w, B), nb2,
                   timeit(lambda: pad3(e2c.to_cube_tensor(frames, out=out, mean=[.485, .456, .406], std=[.229, .224, .225])),
                          args.iters, flush))
            report("e2c+norm+CubePad(3) %d fused B=%d" % (w, B), nb2,
                   timeit(lambda: e2c.to_padded_cube_tensor(frames, 3, mean=[.485, .456, .406], std=[.229, .224, .225], out=padded),
                          args.iters, flush))
            report("e2c+norm+CubePad(3) %d fused, uint8 frames B=%d" % (w, B), nb2,
                   timeit(lambda: e2c.to_padded_cube_tensor(u8, 3, mean=[.485, .456, .406], std=[.229, .224, .225], out=padded),
                          args.iters, flush))

    if args.only in ("", "c2e"):
        for w, C in ((8, 1000), (8, 2048), (7, 1000), (16, 256), (64, 64), (256, 8)):
            bb = B if w <= 16 else 2
            c2e = cp360_b200.Cube2Equi(w)
            x = torch.randn(6 * bb, C, w, w, device=dev)
            report("c2e      [%d,%d,%d,%d]" % (6 * bb, C, w, w), bb * C * 14 * w * w * 4,
                   timeit(lambda: c2e.to_equi_nn(x), args.iters, flush))
            report("c2e+max  [%d,%d,%d,%d]" % (6 * bb, C, w, w), bb * (C * 6 * w * w * 4 + 8 * w * w * 4),
                   timeit(lambda: c2e.to_equi_max(x), args.iters, flush))
            report("c2e cubic [%d,%d,%d,%d]" % (6 * bb, C, w, w), bb * C * 14 * w * w * 4,
                   timeit(lambda: c2e.to_equi_cv2(x), args.iters, flush))
    if args.only in ("", "bwd"):
        # row f1: the training path's backward kernels (train_temporal.py:105-107,167-170)
        for C, H, p in ((64, 128, 1), (128, 64, 1), (256, 32, 1), (512, 16, 1), (2048, 8, 1), (2000, 7, 1), (4000, 7, 1)):
            n = 6 * B
            gy = torch.randn(n, C, H + 2 * p, H + 2 * p, device=dev)
            nbytes = n * C * (H * H + (H + 2 * p) ** 2) * 4
            report("cubepad bwd [%d,%d,%d,%d] p%d" % (n, C, H, H, p), nbytes,
                   timeit(lambda: cp360_b200.cube_pad.cubepad_backward(gy, (p, p, p, p), (H, H)), args.iters, flush))
            del gy
        for w, C in ((7, 1000), (8, 1000)):
            c2e = cp360_b200.Cube2Equi(w)
            g = torch.randn(B, C, 2 * w, 4 * w, device=dev)
            report("c2e bwd  [%d,%d,%d,%d]" % (B, C, 2 * w, 4 * w), B * C * 14 * w * w * 4,
                   timeit(lambda: c2e._backward(g), args.iters, flush))
            # training head (train_temporal.py:105-107): map + channel max, forward and backward
            x = torch.randn(6 * B, C, w, w, device=dev, requires_grad=True)
            gs = torch.randn(B, 2 * w, 4 * w, device=dev)

            def unfused():
                x.grad = None
                c2e.to_equi_nn(x).max(1)[0].backward(gs)

            def fused():
                x.grad = None
                c2e.to_equi_max(x).backward(gs)
            report("c2e+max fwd+bwd [%d,%d,%d,%d] to_equi_nn + torch.max + autograd" % (6 * B, C, w, w),
                   2 * B * C * 6 * w * w * 4, timeit(unfused, args.iters, flush))
            report("c2e+max fwd+bwd [%d,%d,%d,%d] fused (max_arg_fwd + max_bwd)" % (6 * B, C, w, w),
                   2 * B * C * 6 * w * w * 4, timeit(fused, args.iters, flush))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
