#!/usr/bin/env python
"""bench.py — frames/sec of the spherical-projection chain equi->cube->CubePad->cube->equi at
1920x960 (BASELINE.json metric), on N B200s, frames sharded across ranks (weak scaling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step = one pass of the chain (SphericalPipeline.step: e2c, the 18 ResNet-50 CubePad sites, the
2048-channel CubePad site, c2e + channel max) over one batch of B synthetic frames per GPU.
One JSON line on stdout (rank 0). Keys beyond the base contract:
  roofline      dominant kernel: algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  kernels       per-kernel-class share of the step, GB/s (same per-launch event pass)
  e2e           same metric through SphericalPipeline.process_host with pinned HOST frame
                buffers: H2D of every frame and D2H of every saliency map inside the timed region
  cpu_baseline  the reference's CPU path (oracle.ref_port: cv2.remap / torch.cat / grid_sample, the
                library calls the reference makes) on a bounded sample, host cores stated
--impl reference times only that CPU path (rank 0; other ranks exit 0).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frames/sec equi->cube->CubePad->cube->equi @1920x960"
EQUI_H, EQUI_W, CUBE, CAM_C, FEAT_C = 960, 1920, 256, 1000, 2048
WORKLOAD = ("chain per frame: e2c 960x1920x3 -> 6x3x256x256; CubePad at the 18 cubic-ResNet-50 sites "
            "(cube 256) + [6,2048,8,8] p1; c2e+channel-max [6,1000,8,8] -> [16,32]; fp32")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32,
                    help="frames per step per GPU (measured: 16 -> 23.4k, 32 -> 25.0k, 64 -> 25.5k frames/s; every "
                         "launch pays a fixed ~5 us of ramp/tail inside the chain, profiles/README.md)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-frames", type=int, default=2, help="frames per step of --impl reference")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# CPU baseline: the reference's CPU path via oracle.ref_port (the one place bench.py runs oracle/)
# ---------------------------------------------------------------------------------------------
class CpuChain:
    def __init__(self):
        import numpy as np
        import torch
        from oracle import ref_port
        from cp360_b200.pipeline import resnet50_cubepad_sites
        # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1 to its workers, which
        # would otherwise pin the reference's torch ops to one core
        try:
            n_cpu = len(os.sched_getaffinity(0))
        except Exception:
            n_cpu = os.cpu_count() or 1
        if torch.get_num_threads() < n_cpu:
            torch.set_num_threads(n_cpu)
        self.np, self.torch = np, torch
        self.sites = resnet50_cubepad_sites(CUBE) + [(FEAT_C, CUBE // 32, 1)]
        self.e2c = ref_port.Equi2CubePort(CUBE, EQUI_H, EQUI_W)
        self.pads = {p: ref_port.CubePadPort(p) for p in {s[2] for s in self.sites}}
        self.c2e = ref_port.Cube2EquiPort(CUBE // 32)
        rng = np.random.default_rng(0)
        self.frame = rng.random((EQUI_H, EQUI_W, 3), dtype=np.float32)
        g = torch.Generator().manual_seed(0)
        self.feats = [torch.randn((6, C, H, H), generator=g) for C, H, _ in self.sites[1:]]
        self.cam = torch.randn((6, CAM_C, CUBE // 32, CUBE // 32), generator=g)

    def one_frame(self):
        np, torch = self.np, self.torch
        faces = self.e2c.to_cube(self.frame)
        x0 = torch.from_numpy(np.stack([faces[i] for i in range(6)])).permute(0, 3, 1, 2).contiguous()
        self.pads[self.sites[0][2]](x0)
        for (C, H, p), x in zip(self.sites[1:], self.feats):
            self.pads[p](x)
        return self.c2e.to_equi_max(self.cam)

    def threads(self):
        try:
            import cv2
            cvt = cv2.getNumThreads()
        except Exception:
            cvt = 0
        return max(self.torch.get_num_threads(), cvt)


def cpu_baseline(budget_s):
    chain = CpuChain()
    chain.one_frame()                         # warm-up (allocator, cv2 thread pool)
    n, t0 = 0, time.perf_counter()
    while True:
        chain.one_frame()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 512:
            break
    return {"value": round(n / dt, 3), "unit": "frames/s", "cores": chain.threads(), "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": "%d frames of the same chain, one at a time (reference batch_size 1), %.1f s; "
                      "oracle.ref_port = cv2.remap x18 + torch slice/flip/cat CubePad x19 + 6x grid_sample + max"
                      % (n, dt)}


def run_reference(args, rank):
    if rank != 0:
        return
    chain = CpuChain()
    S = max(1, args.ref_frames)
    for _ in range(max(1, min(args.warmup, 3))):
        chain.one_frame()
    steps = max(1, min(args.steps, 200))
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        for _ in range(S):
            chain.one_frame()
        done += 1
        if time.perf_counter() - t0 > 150:     # keep the whole run within a few minutes
            break
    dt = time.perf_counter() - t0
    v = done * S / dt
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": "frames/s",
            "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": round(1e3 * dt / done, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "frames_per_step": S, "device": "host CPU"},
            "cpu_baseline": {"value": round(v, 3), "unit": "frames/s", "cores": chain.threads(), "kind": "port",
                             "host_cpus": os.cpu_count(),
                             "sample": "%d steps x %d frames, %.1f s, oracle.ref_port (the reference is pure "
                                       "Python and cannot travel to the GPU box)" % (done, S, dt)},
            "e2e": {"value": round(v, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------
# clocks during the timed region (NVML, nvidia-smi fallback)
# ---------------------------------------------------------------------------------------------
REASON_BITS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}


class ClockSampler(threading.Thread):
    def __init__(self, torch_device, period=0.02):
        super().__init__(daemon=True)
        self.period, self.samples, self._stop_evt = period, [], threading.Event()
        self.tag = "idle"
        self.handle = self.nv = None
        self.sm_max = None
        try:
            import pynvml as nv
            import torch
            nv.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(torch_device).uuid)
                self.handle = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.handle = nv.nvmlDeviceGetHandleByIndex(torch_device.index or 0)
            self.nv = nv
            self.sm_max = int(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _read(self):
        nv = self.nv
        sm = int(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        try:
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        return sm, r

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            try:
                sm, r = self._read()
                self.samples.append((self.tag, sm, r))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)

    def summary(self):
        if self.nv is None:
            return self._smi_fallback()
        timed = [s for s in self.samples if s[0] == "timed"] or [s for s in self.samples if s[0] != "idle"]
        if not timed:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "source": "nvml"}
        bits = 0
        for _, _, r in timed:
            bits |= r
        reasons = sorted(name for b, name in REASON_BITS.items() if bits & b and name != "gpu_idle")
        return {"sm_mhz": statistics.median(s[1] for s in timed), "sm_max_mhz": self.sm_max,
                "reasons": reasons, "samples": len(timed), "source": "nvml"}

    @staticmethod
    def _smi_fallback():
        import subprocess
        try:
            out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            sm, mx = [float(v) for v in out.split(",")]
            return {"sm_mhz": sm, "sm_max_mhz": mx, "reasons": [], "samples": 1, "source": "nvidia-smi (after region)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "unavailable"}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, read+write)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import cp360_b200
    from cp360_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(3, args.warmup)
    K = max(1, args.steps)
    B = max(1, args.batch)

    pipe = cp360_b200.SphericalPipeline(EQUI_H, EQUI_W, CUBE, CAM_C, FEAT_C, device=dev, seed=1234 + rank)
    pipe.allocate(B)
    frames = pipe.synthetic_frames(B)
    lib = _lib.lib()

    pipe.step(frames)                          # first call of every CubePad site: autotuning happens here
    torch.cuda.synchronize()
    before = _lib.launch_count()
    pipe.step(frames)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - before
    graph = None if args.no_graph else pipe.capture(frames)

    def one_step():
        if graph is not None:
            graph.replay()
        else:
            pipe.step(frames)

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(dev)
    sampler.start()
    for _ in range(W):
        one_step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    barrier()
    torch.cuda.synchronize()
    sampler.tag = "timed"
    if args.profile_range:
        torch.cuda.profiler.start()
    ev0.record()
    for _ in range(K):
        one_step()
    ev1.record()
    torch.cuda.synchronize()
    if args.profile_range:
        torch.cuda.profiler.stop()
    sampler.tag = "after"
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * K / (ms / 1e3)

    # ---- variant (row f2 of SURVEY §8): e2c and the CubePad(3) in front of conv1 as one kernel; same
    # outputs from site 0 on, the unpadded faces are never written. Reported beside the headline.
    fused_first = None
    if world == 1 and not args.no_graph:
        pipe.fuse_first_site = True
        g2 = pipe.capture(frames)
        for _ in range(W):
            g2.replay()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(K):
            g2.replay()
        ev1.record()
        torch.cuda.synchronize()
        ms2 = ev0.elapsed_time(ev1)
        fused_first = {"value": round(B * K / (ms2 / 1e3), 1), "unit": "frames/s", "ms_per_step": round(ms2 / K, 4),
                       "what": "same chain with cp360_e2c_cubepad_fwd replacing e2c + CubePad(3) (21 launches per step)"}
        pipe.fuse_first_site = False
        del g2

    # ---- per-launch CUDA events on the launching stream: which kernel dominates, and its GB/s
    names, marks = [], []

    def hook(name, site):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        names.append((name, site))
        marks.append(e)

    prof_steps = min(K, 20)
    per_class = {}
    per_site = {}
    # prof_steps + 1 steps back to back, no host sync in between, the first one dropped: with the GPU idle
    # at the start of a step the first interval would also contain the host's launch latency
    recorded = []
    for _ in range(prof_steps + 1):
        names.clear()
        marks.clear()
        pipe.step(frames, on_launch=hook)
        recorded.append((list(names), list(marks)))
    torch.cuda.synchronize()
    for step_names, step_marks in recorded[1:]:
        for i in range(len(step_marks) - 1):
            name, site = step_names[i]
            dt = step_marks[i].elapsed_time(step_marks[i + 1])
            if name == "cubepad":
                C, H, p = pipe.sites[site]
                algo = lib.cp360_cubepad_pick_algo(6 * B, C, H, H, p, p, p, p, 4, 1)
                kname = {1: "cubepad_generic_kernel", 3: "cubepad_band_kernel", 4: "cubepad_cube_kernel",
                         5: "cubepad_row_kernel", 6: "cubepad_cube2_kernel"}[algo]
                nbytes = pipe.cubepad_bytes_per_frame(pipe.sites[site]) * B
            elif name == "e2c":
                kname, nbytes = "e2c_kernel", pipe.e2c_bytes_per_frame() * B
            else:
                kname, nbytes = "c2e_small_kernel<max> (+fill)", pipe.c2e_max_bytes_per_frame() * B
            sk = "%s %s" % (kname.split("_kernel")[0], "x".join(str(v) for v in pipe.sites[site]) if name == "cubepad" else "")
            ps = per_site.setdefault(sk.strip(), {"ms": 0.0, "bytes": nbytes, "n": 0})
            ps["ms"] += dt
            ps["n"] += 1
            c = per_class.setdefault(kname, {"ms": 0.0, "bytes": 0, "launches": 0})
            c["ms"] += dt
            c["bytes"] += nbytes
            c["launches"] += 1
    del recorded
    tot_ms = sum(c["ms"] for c in per_class.values())
    peak, peak_src = measured_peak_gbs()
    kernels = {}
    for k, c in per_class.items():
        kernels[k] = {"share": round(c["ms"] / tot_ms, 4), "gbs": round(c["bytes"] / (c["ms"] * 1e-3) / 1e9, 1),
                      "launches_per_step": c["launches"] // prof_steps,
                      "avg_us": round(1e3 * c["ms"] / c["launches"], 2)}
    if os.environ.get("CP360_BENCH_SITES"):
        for k, v in per_site.items():
            us = 1e3 * v["ms"] / prof_steps
            print("site %-34s %8.1f us/step %8.1f GB/s (x%d)" % (k, us, v["bytes"] * (v["n"] // prof_steps) / (us * 1e-6) / 1e9,
                                                                  v["n"] // prof_steps), file=sys.stderr)
    dom = max(per_class, key=lambda k: per_class[k]["ms"])
    dc = per_class[dom]
    achieved = dc["bytes"] / (dc["ms"] * 1e-3) / 1e9
    roofline = {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dc["bytes"] // dc["launches"],
                "avg_launch_us": round(1e3 * dc["ms"] / dc["launches"], 2),
                "chain_gbs": round(pipe.bytes_per_frame() * B * K / (ms * 1e-3) / 1e9 * (1.0 if world == 1 else 1.0), 1),
                "chain_frac": round(pipe.bytes_per_frame() * B * K / (ms * 1e-3) / 1e9 / peak, 4)}
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_path):
        try:
            with open(traffic_path) as f:
                tj = json.load(f)
            # ncu captures are per launch at a stated batch: only valid for a run at that batch
            if int(tj.get("_frames_per_launch", 16)) == B:
                roofline["traffic"] = tj.get(dom)
                roofline["traffic_source"] = "profiles/traffic.json (ncu --set full, B=%d)" % B
        except Exception:
            pass

    # tilings the first-call autotuner settled on (per distinct CubePad site)
    import ctypes
    tuning = {}
    for (C, H, pp) in dict.fromkeys(pipe.sites):
        buf = ctypes.create_string_buffer(160)
        lib.cp360_cubepad_tune_info(6 * B, C, H, H, pp, pp, pp, pp, buf, 160)
        tuning["%dx%dx%d p%d" % (C, H, H, pp)] = buf.value.decode() or "heuristic"

    # ---- end to end: pinned host frames in, host saliency maps out, copies inside the region.
    # Headline = uint8 frames (what a video decoder hands over; converted on the GPU exactly as the
    # reference's float32(u8/255.0)); the float32-host-frame variant is reported beside it.
    e2e = None
    if not args.no_e2e:
        fw = pipe.feat_w
        E = max(4, min(K, 40))
        out_host = torch.empty((E, B, 2 * fw, 4 * fw), dtype=torch.float32).pin_memory()

        wall = []

        def run_e2e(host):
            batches = [host[i % len(host)] for i in range(E)]
            pipe.process_host(batches[:4], out_host[:4])          # warm-up (staging buffers, streams)
            torch.cuda.synchronize()
            barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()                      # the copy stream waits for the compute stream first, so e0 precedes every H2D
            pipe.process_host(batches, out_host)
            e1.record()                      # after the last D2H, which runs on the compute stream
            torch.cuda.synchronize()
            wall.append(time.perf_counter() - t0)
            dt = e0.elapsed_time(e1) / 1e3   # device time (CUDA events)
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return world * B * E / dt

        # several GPUs on one host: pinned staging buffers on the GPU's own NUMA node (best effort, advisory)
        numa = cp360_b200.prefer_gpu_numa_node(dev) if world > 1 else None
        host_u8 = [torch.randint(0, 256, (B, EQUI_H, EQUI_W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        v_u8 = run_e2e(host_u8)
        del host_u8
        host_f32 = [torch.rand((B, EQUI_H, EQUI_W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
        v_f32 = run_e2e(host_f32)
        del host_f32
        e2e = {"value": round(v_u8, 1), "unit": "frames/s",
               "h2d_bytes_per_step": B * EQUI_H * EQUI_W * 3, "d2h_bytes_per_step": B * 2 * fw * 4 * fw * 4,
               "steps": E, "api": "SphericalPipeline.process_host (pinned host uint8 frames -> host saliency maps)",
               "timing": "CUDA events on the compute stream around the whole call, max over ranks",
               "wall_clock_value": round(world * B * E / wall[0], 1),
               "f32_host_frames": {"value": round(v_f32, 1), "h2d_bytes_per_step": B * EQUI_H * EQUI_W * 3 * 4}}
        if numa is not None:
            e2e["host_numa_preference_rank0"] = numa
    sampler.stop()
    clocks = sampler.summary()

    # the single collective of the path: gather every rank's maps (outside the timed region)
    if world > 1:
        maps = cp360_b200.gather_maps(pipe.sal, world * B)
        assert maps.shape[0] == world * B

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.cpu_seconds)

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "global_frames_per_step": world * B,
                           "sharding": "frames block-partitioned over ranks, no data-path collective",
                           "launch": "CUDA graph replay" if graph is not None else "eager C-ABI launches",
                           "l2": "inputs larger than L2: %.2f GB touched per step per GPU, no flush needed"
                                 % (pipe.bytes_per_frame() * B / 1e9),
                           "algorithmic_bytes_per_frame": pipe.bytes_per_frame()},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * K),
                "roofline": roofline, "kernels": kernels, "tuning": tuning, "fused_first_site": fused_first,
                "cpu_baseline": cpu}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def claim_stdout():
    """Keep stdout for the ONE JSON line: native libraries write there too (NCCL prints its version banner on
    stdout at communicator creation), so fd 1 is pointed at stderr for the rest of the process and the line
    goes to a private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it (one process per GPU)
        import subprocess
        port = 29500 + (os.getpid() % 2000)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    claim_stdout()
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
