#!/usr/bin/env python
"""Shapes for compute-sanitizer racecheck / memcheck that make every hand-rolled shared-memory pipeline WRAP its
ring (the parity suite's small shapes give each CTA / warp fewer tiles than ring slots, so slot re-use — the only
place a write-after-read hazard can exist — is never exercised there). Every result is checked against a torch
index_select application of the host index map, so a real race would also show up as a wrong value.

    compute-sanitizer --tool racecheck python tools/racecheck_probe.py [--only row,cube,bwd,c2e,e2c]
    CP360_LIB=.../libcp360_arriveall.so ... (the -DCP360_ARRIVE_ALL diagnostic build, profiles/README.md)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cp360_b200  # noqa: E402
from cp360_b200 import _lib  # noqa: E402


def gather_reference(x, imap):
    n6, C, H, W = x.shape
    _, Ho, Wo = imap.shape
    g = x.reshape(n6 // 6, 6, C, H * W).permute(0, 2, 1, 3).reshape(n6 // 6, C, 6 * H * W)
    idx = torch.from_numpy(imap.reshape(-1).astype(np.int64)).to(x.device)
    out = g.index_select(2, idx).reshape(n6 // 6, C, 6, Ho, Wo).permute(0, 2, 1, 3, 4)
    return out.reshape(n6, C, Ho, Wo).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="row,cube,bwd,c2e,e2c")
    args = ap.parse_args()
    only = set(args.only.split(","))
    dev = torch.device("cuda", 0)
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    print("library:", cp360_b200.LIB_PATH, "SMs:", sm, flush=True)
    g = torch.Generator(device=dev).manual_seed(3)
    ok = True

    def check(tag, got, want):
        nonlocal ok
        same = torch.equal(got, want)
        ok &= same
        print("%-58s %s" % (tag, "bit-exact" if same else "MISMATCH"), flush=True)

    fwd_cases = []
    if "row" in only:       # (n, C, H, p, algo): bands of rows (many tiles per warp), whole-plane tiles, p=3 stem
        fwd_cases += [(12, 64, 128, 1, _lib.ALGO_ROW), (12, 3, 224, 3, _lib.ALGO_ROW), (48, 256, 28, 1, _lib.ALGO_ROW),
                      (24, 128, 56, 1, _lib.ALGO_ROW)]
    if "cube" in only:      # chunks per CTA > ring depth
        fwd_cases += [(48, 2048, 8, 1, _lib.ALGO_CUBE2), (48, 512, 16, 1, _lib.ALGO_CUBE2), (24, 256, 32, 1, _lib.ALGO_CUBE2),
                      (48, 2000, 7, 1, _lib.ALGO_CUBE2), (24, 512, 14, 1, _lib.ALGO_CUBE2)]
    for n, C, H, p, algo in fwd_cases:
        x = torch.randn((n, C, H, H), device=dev, generator=g)
        y = cp360_b200.cubepad_forward(x, (p, p, p, p), algo=algo)
        want = gather_reference(x, cp360_b200.cubepad_index_map(H, H, p))
        check("cubepad fwd algo %d [%d,%d,%d,%d] p%d" % (algo, n, C, H, H, p), y, want)
        if algo == _lib.ALGO_CUBE2:
            sc, sh = torch.rand(C, device=dev, generator=g) + 0.5, torch.randn(C, device=dev, generator=g)
            yf = cp360_b200.cubepad_fused(x, (p, p, p, p), scale=sc, shift=sh, relu=True)
            wf = gather_reference(torch.relu(x * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)),
                                  cp360_b200.cubepad_index_map(H, H, p))
            check("cubepad fused bn+relu [%d,%d,%d,%d]" % (n, C, H, H), yf, wf)
        del x, y, want
    if "bwd" in only:
        for n, C, H, p in [(48, 2048, 8, 1), (48, 512, 16, 1), (24, 256, 32, 1), (48, 2000, 7, 1), (12, 64, 64, 1)]:
            gy = torch.randn((n, C, H + 2 * p, H + 2 * p), device=dev, generator=g)
            gx = cp360_b200.cube_pad.cubepad_backward(gy, (p, p, p, p), (H, H))
            os.environ["CP360_BWD_ALGO"] = "1"                  # the two-kernel path (no shared-memory ring)
            gx2 = cp360_b200.cube_pad.cubepad_backward(gy, (p, p, p, p), (H, H))
            del os.environ["CP360_BWD_ALGO"]
            check("cubepad bwd [%d,%d,%d,%d] p%d cube-tile == two-kernel" % (n, C, H, H, p), gx, gx2)
            del gy, gx, gx2
    if "c2e" in only:
        for w, C, B in [(8, 1000, 40), (7, 1000, 40), (16, 256, 8)]:
            c2e = cp360_b200.Cube2Equi(w)
            x = torch.randn((6 * B, C, w, w), device=dev, generator=g)
            full = c2e.to_equi_nn(x)
            mx = c2e.to_equi_max(x)
            same = float((mx - full.max(1)[0]).abs().max()) == 0.0
            ok &= same
            print("%-58s %s" % ("c2e max == max(c2e) [%d,%d,%d,%d]" % (6 * B, C, w, w), "bit-exact" if same else "MISMATCH"), flush=True)
    if "c2ebwd" in only or "c2e" in only:
        # shared-memory backward: two-stage ring re-used every second channel group (B = 40: each CTA walks ~9-20 groups)
        for w, C, B in [(8, 1000, 40), (16, 300, 30)]:
            c2e = cp360_b200.Cube2Equi(w)
            gy = torch.randn((B, C, 2 * w, 4 * w), device=dev, generator=g)
            x = torch.zeros((6 * B, C, w, w), device=dev, requires_grad=True)
            c2e.to_equi_nn(x).backward(gy)                       # same plan, but x.grad from our kernel
            got = c2e._backward(gy)
            os.environ["CP360_C2E_BWD_SMALL"] = "0"
            want = c2e._backward(gy)
            del os.environ["CP360_C2E_BWD_SMALL"]
            diff = float((got - want).abs().max())
            good = diff <= 1e-4 and torch.equal(got, x.grad)
            ok &= good
            print("%-58s %s (max diff vs read-only-path gather %.2e)" % ("c2e bwd [%d,%d,%d,%d]" % (B, C, 2 * w, 4 * w),
                                                                        "ok" if good else "MISMATCH", diff), flush=True)
    if "e2c" in only:
        rng = np.random.default_rng(0)
        H, W, w, B = 240, 480, 64, 5
        img = rng.random((H, W, 3), dtype=np.float32)
        e2c = cp360_b200.Equi2Cube(w, img)
        fr = torch.rand((B, H, W, 3), device=dev, generator=g)
        a = e2c.to_cube_tensor(fr)
        b = e2c.to_cube_tensor((fr * 255).to(torch.uint8))
        pa = e2c.to_padded_cube_tensor(fr, 3)
        check("e2c+CubePad(3) fused == CubePad(e2c) [%d frames]" % B, pa, cp360_b200.CubePad(3)(a))
        print("e2c u8 path ran:", tuple(b.shape), flush=True)
    torch.cuda.synchronize()
    print("racecheck_probe:", "ALL OK" if ok else "FAILURES", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
