#!/usr/bin/env python
"""Host->device ceiling of this box, per GPU count (run under torchrun for N > 1): every rank copies the e2e
path's per-step payload (B uint8 frames, 176.9 MB at B=32) from page-locked host memory to its GPU, all ranks
at once, no kernels. Variants: how the host buffer was allocated (torch pin_memory / cp360_host_alloc modes)
and how many copy streams feed the GPU. Prints one JSON line per variant on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/h2d_probe.py
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cp360_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--variants", default="torch_pin,pinned,write_combined,hugepage")
    ap.add_argument("--streams", default="1,2")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shape = (args.batch, 960, 1920, 3)
    nbytes = args.batch * 960 * 1920 * 3
    dst = [torch.empty(shape, dtype=torch.uint8, device=dev) for _ in range(2)]
    if rank == 0:
        try:
            thp = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
        except OSError:
            thp = "?"
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] if os.path.isdir("/sys/devices/system/node") else []
        print(json.dumps({"probe": "env", "world": world, "cpus": len(os.sched_getaffinity(0)), "numa_nodes": len(nodes),
                          "thp": thp, "bytes_per_copy": nbytes}), flush=True)
    for variant in args.variants.split(","):
        try:
            if variant == "torch_pin":
                host = [torch.empty(shape, dtype=torch.uint8).pin_memory() for _ in range(2)]
            else:
                host = [cp360_b200.pinned_empty(shape, torch.uint8, mode=variant) for _ in range(2)]
            for h in host:
                h.random_(0, 256) if variant != "write_combined" else h.fill_(7)
        except Exception as e:                      # noqa: BLE001
            if rank == 0:
                print(json.dumps({"probe": "h2d", "variant": variant, "error": str(e)[:200]}), flush=True)
            continue
        for ns in [int(v) for v in args.streams.split(",")]:
            streams = [torch.cuda.Stream(device=dev) for _ in range(ns)]

            def burst(n):
                part = (args.batch + ns - 1) // ns
                for i in range(n):                     # every payload split across the streams
                    for k in range(ns):
                        with torch.cuda.stream(streams[k]):
                            dst[i & 1][k * part:(k + 1) * part].copy_(host[i & 1][k * part:(k + 1) * part], non_blocking=True)
            burst(2)
            torch.cuda.synchronize()
            # calibrate the number of copies for ~args.seconds
            t0 = time.perf_counter()
            burst(4)
            torch.cuda.synchronize()
            per = (time.perf_counter() - t0) / 4
            n = max(4, min(400, int(args.seconds / max(per, 1e-4))))
            if world > 1:
                t = torch.tensor([n], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                n = int(t.item())
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for s in streams:
                s.wait_stream(torch.cuda.current_stream(dev))
            e0.record()
            for s in streams:
                s.wait_event(e0)
            burst(n)
            for s in streams:
                torch.cuda.current_stream(dev).wait_stream(s)
            e1.record()
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) / 1e3
            gbs = n * nbytes / dt / 1e9
            vals = [gbs]
            if world > 1:
                t = torch.tensor([gbs, dt], dtype=torch.float64, device=dev)
                all_t = [torch.zeros_like(t) for _ in range(world)]
                dist.all_gather(all_t, t)
                vals = [float(a[0]) for a in all_t]
                tmax = max(float(a[1]) for a in all_t)
                agg = world * n * nbytes / tmax / 1e9
            else:
                agg = gbs
            if rank == 0:
                print(json.dumps({"probe": "h2d", "variant": variant, "streams": ns, "world": world, "copies": n,
                                  "per_gpu_gbs": [round(v, 1) for v in vals], "aggregate_gbs": round(agg, 1)}), flush=True)
        del host
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
