#!/usr/bin/env python
"""Device-side timeline of one chain step replayed as a CUDA graph (GPU box).

Needs the trace build of the library:
    CP360_LIB=$PWD/cp-360-weakly-supervised-saliency_b200/lib/libcp360_trace.so CP360_NVCC_EXTRA=-DCP360_TRACE \
        python -c "import cp360_b200; cp360_b200.build_library(force=True)"
    CP360_LIB=.../libcp360_trace.so python tools/trace_chain.py [--batch 16] [--json out.json]

Every CTA records %globaltimer at entry (t0), after griddepcontrol.wait (t1), when its first tile
landed (t2) and at exit (t3). Per launch this prints: gap to the previous launch's last exit,
prologue, pipeline fill, body, and the spread of CTA exit times (tail). Numbers are taken from a
graph replay after warm-up; the trace build is never the one benchmarked.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import cp360_b200  # noqa: E402
from cp360_b200 import _lib  # noqa: E402

KID = {1: "row", 2: "cube2", 3: "e2c", 4: "c2e_small"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--json", default="")
    ap.add_argument("--eager", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.lib()
    cap = 1 << 20
    rec = torch.zeros(cap * 12, dtype=torch.int32, device=dev)      # 48 B records
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    for name in ("cubepad", "e2c", "c2e"):
        fn = getattr(lib, "cp360_trace_bind_" + name)
        fn.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]
        fn.restype = ctypes.c_int
        assert fn(rec.data_ptr(), cap, cnt.data_ptr()) == 0

    pipe = cp360_b200.SphericalPipeline(device=dev)
    pipe.allocate(args.batch)
    frames = pipe.synthetic_frames(args.batch)
    graph = None if args.eager else pipe.capture(frames)
    for _ in range(5):
        graph.replay() if graph else pipe.step(frames)
    torch.cuda.synchronize()
    cnt.zero_()
    torch.cuda.synchronize()
    graph.replay() if graph else pipe.step(frames)
    torch.cuda.synchronize()
    n = int(cnt.item())
    raw = rec[: n * 12].cpu().numpy().view(np.uint8).reshape(n, 48)
    t = raw[:, :32].copy().view(np.uint64).reshape(n, 4).astype(np.int64)
    meta = raw[:, 32:].copy().view(np.uint32).reshape(n, 4)
    kid, cta, smid, nctas = meta[:, 0], meta[:, 1], meta[:, 2], meta[:, 3]
    # split into launches: records of one launch share (kid, nctas) and are contiguous in time;
    # sort by entry time and cut when cta ids restart for the same kid or kid changes
    order = np.argsort(t[:, 0], kind="stable")
    launches, cur, seen = [], [], set()
    for i in order:
        key = (int(kid[i]), int(nctas[i]))
        if cur and (key != cur_key or int(cta[i]) in seen):
            launches.append(cur)
            cur, seen = [], set()
        cur_key = key
        cur.append(i)
        seen.add(int(cta[i]))
    if cur:
        launches.append(cur)
    t_base = int(t[order[0], 0])
    rows, prev_end = [], None
    print("%-3s %-10s %6s %9s %8s %8s %8s %8s %8s %8s %6s" % ("#", "kernel", "ctas", "start_us", "gap", "entryspr", "prolog",
                                                             "fill", "body", "tailspr", "sms"))
    for li, idx in enumerate(launches):
        idx = np.array(idx)
        t0, t1, t2, t3 = (t[idx, k] for k in range(4))
        start, end = int(t0.min()), int(t3.max())
        ok2 = (t2 > 0) & (t2 < (1 << 62))
        row = {"launch": li, "kernel": KID.get(int(kid[idx[0]]), str(kid[idx[0]])), "ctas": len(idx),
               "start_us": (start - t_base) / 1e3, "gap_us": None if prev_end is None else (start - prev_end) / 1e3,
               "entry_spread_us": (int(t0.max()) - start) / 1e3,
               "prologue_us": float(np.median(t1 - t0)) / 1e3,
               "wait_done_us": (int(t1.max()) - start) / 1e3,
               "fill_us": float(np.median((t2 - t1)[ok2])) / 1e3 if ok2.any() else None,
               "body_us": (end - int(t1.min())) / 1e3,
               "tail_spread_us": (end - int(np.percentile(t3, 10))) / 1e3,
               "total_us": (end - start) / 1e3, "sms": int(len(set(smid[idx].tolist()))),
               "cta_end_pcts_us": [round((float(np.percentile(t3, q)) - int(t1.min())) / 1e3, 1) for q in (0, 10, 50, 90, 100)]}
        rows.append(row)
        prev_end = end
        print("%-3d %-10s %6d %9.1f %8s %8.1f %8.1f %8s %8.1f %8.1f %6d" % (
            li, row["kernel"], row["ctas"], row["start_us"], "-" if row["gap_us"] is None else "%.1f" % row["gap_us"],
            row["entry_spread_us"], row["prologue_us"], "-" if row["fill_us"] is None else "%.1f" % row["fill_us"],
            row["body_us"], row["tail_spread_us"], row["sms"]), row["cta_end_pcts_us"])
    tot = (max(int(t[i, 3]) for i in order) - t_base) / 1e3
    print("records %d, launches %d, step span %.1f us; sum gaps %.1f us, sum tail spread %.1f us" % (
        n, len(launches), tot, sum(r["gap_us"] or 0 for r in rows), sum(r["tail_spread_us"] for r in rows)))
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
