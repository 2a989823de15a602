#!/usr/bin/env python
"""bench.py — frames/sec of the spherical-projection hot path at 1920x960 (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--cube 256|224]
                    [--workload chain|clstm|corpus] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json configs):
  chain   (default, the headline) one step = SphericalPipeline.step over B frames per GPU: e2c, the 18 cubic-ResNet-50
          CubePad sites, the 2048-channel CubePad site, c2e + channel max. --cube 224 is the reference's own
          cube_dim (config.yaml:17); weak scaling, frames sharded over ranks. e2c and the CubePad(3) in front of
          conv1 run as ONE kernel (the same padded tensor; 20 launches per step) unless --no-fuse-first-site
          (21 launches: the faces are written, then read by the stem pad — round 1's definition).
  clstm   configs[3]: one step = one 80-frame video through the ConvLSTM-side hot path: per output frame a 5-step
          window (test_temporal.py:57-79) of the three CubePads of a cell evaluation (clstm.py:57-64) + c2e + max
          of the hidden state; B windows batched per launch. --clstm-variant reference = 1000 channels on 7x7 faces.
  corpus  configs[4]: one step = the whole 2 000-frame corpus (25 videos x 80 frames), device-resident, block-sharded
          over the ranks (strong scaling); the final gather of the maps is timed on its own.
One JSON line on stdout (rank 0). Keys beyond the base contract:
  roofline          dominant kernel: algorithmic bytes / CUDA-event duration vs MEASURED_PEAKS.json
  kernels           per-kernel-class share of the step, GB/s (same per-launch event pass)
  e2e               same metric through SphericalPipeline.process_host with page-locked HOST frame buffers (H2D of
                    every frame and D2H of every map inside the timed region), plus the upload-only ceiling of the
                    same buffers/streams (h2d_ceiling_*): e2e/ceiling says whether the box or the code is the limit
  fused_chain       the chain with every producer-side fusion (SphericalPipeline.step_fused)
  gpu_aten_baseline the reference's own GPU path (ATen cat/index_select CubePad, grid_sample c2e) on this GPU
  cpu_baseline      the reference's CPU path on a bounded sample, host cores stated (kind "reference" when the
                    unmodified reference is staged in oracle/_ref, else the library-call port oracle.ref_port)
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
EQUI_H, EQUI_W, CAM_C, FEAT_C = 960, 1920, 1000, 2048
CORPUS_FRAMES, VIDEO_FRAMES, SEQ_LEN = 2000, 80, 5


def workload_string(args):
    fw = args.cube // 32
    chain = ("chain per frame: e2c 960x1920x3 -> 6x3x%dx%d; CubePad at the 18 cubic-ResNet-50 sites "
             "(cube %d) + [6,%d,%d,%d] p1; c2e+channel-max [6,%d,%d,%d] -> [%d,%d]; fp32"
             % (args.cube, args.cube, args.cube, FEAT_C, fw, fw, CAM_C, fw, fw, 2 * fw, 4 * fw))
    if args.workload in ("chain", "corpus") and not getattr(args, "no_fuse_first_site", False):
        chain += "; e2c and the CubePad(3) in front of conv1 run as one kernel (same padded tensor, the unpadded faces are not materialised)"
    if args.workload == "chain":
        return chain
    if args.workload == "corpus":
        return "corpus of %d frames (25 videos x 80), device-resident fp32, block-sharded over ranks; per frame: %s" % (CORPUS_FRAMES, chain)
    c, w = clstm_dims(args)
    return ("ConvLSTM hot path per 80-frame video: per output frame a %d-step window of CubePad(1) on "
            "[6,%d,%d,%d] (cat of input and hidden), 2x [6,%d,%d,%d], then c2e+channel-max of the hidden state "
            "[6,%d,%d,%d]; fp32" % (SEQ_LEN, 2 * c, w, w, 4 * c, w, w, c, w, w))


def clstm_dims(args):
    return (1000, 7) if args.clstm_variant == "reference" else (FEAT_C, 8)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None,
                    help="chain: frames per step per GPU (default 32; the reference runs batch_size 1, config.yaml:33); "
                         "clstm: windows per launch (default 16); corpus: frames per launch (default: largest "
                         "divisor <= 40 of the rank's share)")
    ap.add_argument("--cube", type=int, default=256, choices=[224, 256],
                    help="face width: 256 = BASELINE.json's configuration, 224 = the reference's cube_dim (config.yaml:17)")
    ap.add_argument("--workload", default="chain", choices=["chain", "clstm", "corpus"])
    ap.add_argument("--clstm-variant", default="baseline", choices=["baseline", "reference"],
                    help="baseline: 2048 channels on 8x8 faces (BASELINE.json); reference: 1000 channels on 7x7 (config.yaml:21-22)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-aten-baseline", action="store_true")
    ap.add_argument("--no-fused", action="store_true")
    ap.add_argument("--min-untimed-steps", type=int, default=40,
                    help="at least this many untimed steps before a timed region (max with --warmup; 0 = exactly --warmup)")
    ap.add_argument("--no-fuse-first-site", action="store_true",
                    help="chain / corpus: run e2c and the CubePad(3) in front of conv1 as two launches (the unpadded faces "
                         "are written and read back) instead of the one-kernel first site (cp360_e2c_cubepad_fwd)")
    ap.add_argument("--host-alloc", default="torch_pin", choices=["torch_pin", "pinned", "write_combined", "hugepage"],
                    help="how the e2e host frame buffers are page-locked (cp360_b200.pinned_empty modes)")
    ap.add_argument("--e2e-depth", type=int, default=3, help="device staging ring depth of process_host")
    ap.add_argument("--e2e-copy-streams", type=int, default=1)
    ap.add_argument("--profile-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--ref-frames", type=int, default=2, help="frames per step of --impl reference")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {"chain": 200, "clstm": 50, "corpus": 5}[args.workload]
    if args.warmup is None:
        args.warmup = {"chain": 10, "clstm": 5, "corpus": 3}[args.workload]
    return args


# ---------------------------------------------------------------------------------------------
# CPU baseline: the reference's CPU path. The unmodified reference when it is staged (oracle/_ref, see
# oracle/ref_loader.py) — kind "reference"; else oracle.ref_port — kind "port". The one place bench.py runs oracle/.
# ---------------------------------------------------------------------------------------------
class CpuChain:
    def __init__(self, args):
        import numpy as np
        import torch
        from oracle import ref_loader
        from cp360_b200.pipeline import resnet50_cubepad_sites
        # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1 to its workers, which
        # would otherwise pin the reference's torch ops to one core
        try:
            n_cpu = len(os.sched_getaffinity(0))
        except Exception:
            n_cpu = os.cpu_count() or 1
        if torch.get_num_threads() < n_cpu:
            torch.set_num_threads(n_cpu)
        self.np, self.torch, self.args = np, torch, args
        cube, fw = args.cube, args.cube // 32
        self.clstm = args.workload == "clstm"
        if self.clstm:
            c, fw = clstm_dims(args)
            self.sites = [(2 * c, fw, 1), (4 * c, fw, 1), (4 * c, fw, 1)]
        else:
            self.sites = resnet50_cubepad_sites(cube) + [(FEAT_C, fw, 1)]
        self.kind = "port"
        if ref_loader.available() and (ref_loader.kind() == "live" or ref_loader.verify()):
            import warnings
            warnings.filterwarnings("ignore")
            cube_pad, e2c_mod, c2e_mod = ref_loader.load("cpu")
            self.kind = "reference"
            frame0 = np.zeros((EQUI_H, EQUI_W, 3), np.float32)
            self.e2c = None if self.clstm else e2c_mod.Equi2Cube(cube, frame0)
            self.pads = {p: cube_pad.CubePad(p, use_gpu=False) for p in {s[2] for s in self.sites}}
            self.c2e = c2e_mod.Cube2Equi(fw)
            self.c2e_max = lambda x: torch.max(self.c2e.to_equi_nn(x), 1)[0]            # test_temporal.py:82-84
            self.what = ("the UNMODIFIED reference (%s): Equi2Cube.to_cube (18x cv2.remap) + CubePad(use_gpu=False) x%d + "
                         "Cube2Equi.to_equi_nn (6x grid_sample, CPU-patched) + torch.max"
                         % ("oracle/_ref staging copy, sha256-verified" if ref_loader.kind() == "staged" else "/root/reference",
                            len(self.sites) * (SEQ_LEN if self.clstm else 1)))
        else:
            from oracle import ref_port
            self.e2c = None if self.clstm else ref_port.Equi2CubePort(cube, EQUI_H, EQUI_W)
            self.pads = {p: ref_port.CubePadPort(p) for p in {s[2] for s in self.sites}}
            port = ref_port.Cube2EquiPort(fw)
            self.c2e_max = port.to_equi_max
            self.what = "oracle.ref_port = cv2.remap x18 + torch slice/flip/cat CubePad + 6x grid_sample + max (reference not staged)"
        rng = np.random.default_rng(0)
        self.frame = rng.random((EQUI_H, EQUI_W, 3), dtype=np.float32)
        g = torch.Generator().manual_seed(0)
        if self.clstm:
            c, fw = clstm_dims(args)
            self.x, self.h = torch.randn((6, c, fw, fw), generator=g), torch.randn((6, c, fw, fw), generator=g)
            self.feats = [torch.randn((6, C, H, H), generator=g) for C, H, _ in self.sites[1:]]
        else:
            self.feats = [torch.randn((6, C, H, H), generator=g) for C, H, _ in self.sites[1:]]
            self.cam = torch.randn((6, CAM_C, fw, fw), generator=g)

    def one_frame(self):
        np, torch = self.np, self.torch
        with torch.no_grad():
            if self.clstm:
                for _ in range(SEQ_LEN):
                    self.pads[1](torch.cat((self.x, self.h), 1))              # clstm.py:57-58
                    for x in self.feats:
                        self.pads[1](x)
                return self.c2e_max(self.h)
            faces = self.e2c.to_cube(self.frame)
            x0 = torch.from_numpy(np.stack([faces[i] for i in range(6)])).permute(0, 3, 1, 2).contiguous()
            self.pads[self.sites[0][2]](x0)
            for (C, H, p), x in zip(self.sites[1:], self.feats):
                self.pads[p](x)
            return self.c2e_max(self.cam)

    def threads(self):
        try:
            import cv2
            cvt = cv2.getNumThreads()
        except Exception:
            cvt = 0
        return max(self.torch.get_num_threads(), cvt)


def cpu_baseline(args, budget_s):
    chain = CpuChain(args)
    chain.one_frame()                         # warm-up (allocator, cv2 thread pool)
    n, t0 = 0, time.perf_counter()
    while True:
        chain.one_frame()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 512:
            break
    return {"value": round(n / dt, 3), "unit": "frames/s", "cores": chain.threads(), "kind": chain.kind,
            "host_cpus": os.cpu_count(),
            "sample": "%d frames of the same workload, one at a time (reference batch_size 1), %.1f s; %s" % (n, dt, chain.what)}


def run_reference(args, rank):
    if rank != 0:
        return
    chain = CpuChain(args)
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
            "higher_is_better": True, "scaling": "strong" if args.workload == "corpus" else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args), "frames_per_step": S, "device": "host CPU"},
            "cpu_baseline": {"value": round(v, 3), "unit": "frames/s", "cores": chain.threads(), "kind": chain.kind,
                             "host_cpus": os.cpu_count(),
                             "sample": "%d steps x %d frames, %.1f s; %s" % (done, S, dt, chain.what)},
            "e2e": {"value": round(v, 3), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def gpu_aten_baseline(args, dev, frames=6):
    """Second baseline: the reference's own GPU path on this B200 — CubePad through ATen cat /
    index_select with its per-flip index uploads (cube_pad.py:73-76,95-216) and to_equi_nn through 6 grid_sample
    passes (cube_to_equi.py:37-66); e2c has no GPU path in the reference (cv2.remap on the host,
    equi_to_cube.py:112-129) and runs there, followed by the upload the extractor does (class_activation_model.py:58)."""
    import numpy as np
    import torch
    from oracle import ref_loader
    from cp360_b200.pipeline import resnet50_cubepad_sites
    if not (ref_loader.available() and (ref_loader.kind() == "live" or ref_loader.verify())):
        return {"unavailable": "reference not staged in oracle/_ref"}
    import warnings
    warnings.filterwarnings("ignore")
    cube_pad, e2c_mod, c2e_mod = ref_loader.load("cuda")
    cube, fw = args.cube, args.cube // 32
    sites = resnet50_cubepad_sites(cube) + [(FEAT_C, fw, 1)]
    frame = np.random.default_rng(0).random((EQUI_H, EQUI_W, 3), dtype=np.float32)
    e2c = e2c_mod.Equi2Cube(cube, frame)
    pads = {p: cube_pad.CubePad(p).to(dev) for p in {s[2] for s in sites}}
    c2e = c2e_mod.Cube2Equi(fw)
    g = torch.Generator(device=dev).manual_seed(0)
    feats = [torch.randn((6, C, H, H), device=dev, generator=g) for C, H, _ in sites[1:]]
    cam = torch.randn((6, CAM_C, fw, fw), device=dev, generator=g)

    def one(timers=None):
        t0 = time.perf_counter()
        faces = e2c.to_cube(frame)
        x0 = torch.from_numpy(np.stack([faces[i] for i in range(6)])).permute(0, 3, 1, 2).contiguous().to(dev)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        with torch.no_grad():
            pads[sites[0][2]](x0)
            for (C, H, p), x in zip(sites[1:], feats):
                pads[p](x)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            sal = torch.max(c2e.to_equi_nn(cam), 1)[0]
            sal.cpu()
        t3 = time.perf_counter()
        if timers is not None:
            timers.append((t1 - t0, t2 - t1, t3 - t2))
    one()
    one()
    timers = []
    t0 = time.perf_counter()
    for _ in range(frames):
        one(timers)
    dt = time.perf_counter() - t0
    med = lambda k: round(1e3 * statistics.median(t[k] for t in timers), 3)                # noqa: E731
    return {"value": round(frames / dt, 2), "unit": "frames/s", "frames": frames,
            "e2c_host_plus_upload_ms": med(0), "cubepad_x19_gpu_ms": med(1), "c2e_max_gpu_ms": med(2),
            "what": "the unmodified reference (%s) on this GPU, one frame at a time (batch_size 1): Equi2Cube.to_cube on the "
                    "host (its only path) + upload, CubePad (ATen cat/index_select, use_gpu=True) at the 19 sites, "
                    "to_equi_nn (6x grid_sample + masked scatter) + torch.max; wall clock with synchronize"
                    % ("oracle/_ref" if ref_loader.kind() == "staged" else "/root/reference")}


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
# shared measurement helpers of the B200 arm
# ---------------------------------------------------------------------------------------------
class Ctx:
    """Rank / device plumbing and the contract's timed loop."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.world = rank, world
        self.dev = torch.device("cuda", local_rank)
        torch.cuda.set_device(self.dev)
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.W, self.K = max(3, args.warmup), max(1, args.steps)
        # steady state: a run with few warm-up steps reads 2-3 % low (tools/gpu_r2_call41.sh: 20 steps after 3 warm-up steps
        # 25.9-26.2 k frames/s, after 50: 27.0 k, 200 steps after 10: 26.5 k on the same box) — the first replays of a graph
        # and the clock ramp out of idle. At least this many untimed steps precede every timed region (stated in `config`).
        self.min_untimed = 0 if args.profile_range else max(0, args.min_untimed_steps)
        self.sampler = ClockSampler(self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def timed(self, one_step, steps=None, warmup=None, tag=True, profile=False):
        """W untimed steps, then exactly K steps between barrier + synchronize; CUDA events, max over ranks -> ms."""
        torch = self.torch
        K = self.K if steps is None else steps
        W = self.W if warmup is None else warmup
        for _ in range(max(W, self.min_untimed)):
            one_step()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        self.barrier()
        torch.cuda.synchronize()
        if tag:
            self.sampler.tag = "timed"
        if profile:
            torch.cuda.profiler.start()
        ev0.record()
        for _ in range(K):
            one_step()
        ev1.record()
        torch.cuda.synchronize()
        if profile:
            torch.cuda.profiler.stop()
        if tag:
            self.sampler.tag = "after"
        self.barrier()
        return self.max_over_ranks(ev0.elapsed_time(ev1))


def event_pass(torch, run_with_hook, prof_steps):
    """Per-launch CUDA events on the launching stream: run_with_hook(hook) performs one eager step and calls
    hook(name, site) before every C-ABI launch and once at the end. prof_steps + 1 steps back to back, the first
    dropped (with the GPU idle the first interval would also contain the host's launch latency).
    Returns {(name, site): [total_ms, count]} over prof_steps steps."""
    recorded = []
    names, marks = [], []

    def hook(name, site):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        names.append((name, site))
        marks.append(e)
    for _ in range(prof_steps + 1):
        names.clear()
        marks.clear()
        run_with_hook(hook)
        recorded.append((list(names), list(marks)))
    torch.cuda.synchronize()
    acc = {}
    for step_names, step_marks in recorded[1:]:
        for i in range(len(step_marks) - 1):
            a = acc.setdefault(step_names[i], [0.0, 0])
            a[0] += step_marks[i].elapsed_time(step_marks[i + 1])
            a[1] += 1
    return acc


CUBEPAD_KERNELS = {1: "cubepad_generic_kernel", 3: "cubepad_band_kernel", 4: "cubepad_cube_kernel",
                   5: "cubepad_row_kernel", 6: "cubepad_cube2_kernel"}


def cubepad_kernel_name(lib, n_faces, C, H, p):
    return CUBEPAD_KERNELS[lib.cp360_cubepad_pick_algo(n_faces, C, H, H, p, p, p, p, 4, 1)]


def tuning_info(lib, n_faces, sites):
    import ctypes
    out = {}
    for (C, H, pp) in dict.fromkeys(sites):
        buf = ctypes.create_string_buffer(200)
        lib.cp360_cubepad_tune_info(n_faces, C, H, H, pp, pp, pp, pp, buf, 200)
        out["%dx%dx%d p%d" % (C, H, H, pp)] = buf.value.decode() or "heuristic"
    return out


def roofline_from(per_class, peak, peak_src, extra=None):
    dom = max(per_class, key=lambda k: per_class[k]["ms"])
    dc = per_class[dom]
    achieved = dc["bytes"] / (dc["ms"] * 1e-3) / 1e9
    r = {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
         "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
         "algorithmic_bytes_per_launch": dc["bytes"] // dc["launches"],
         "avg_launch_us": round(1e3 * dc["ms"] / dc["launches"], 2)}
    if extra:
        r.update(extra)
    return r, dom


def kernels_table(per_class, prof_steps):
    tot_ms = sum(c["ms"] for c in per_class.values())
    return {k: {"share": round(c["ms"] / tot_ms, 4), "gbs": round(c["bytes"] / (c["ms"] * 1e-3) / 1e9, 1),
                "launches_per_step": c["launches"] // prof_steps, "avg_us": round(1e3 * c["ms"] / c["launches"], 2)}
            for k, c in per_class.items()}


def attach_traffic(roofline, dom, B, cube, fused_first=True):
    """ncu captures are per launch at a stated batch and face width: only valid for a run at that configuration
    (and, for a class average, at that composition of the class: with or without the stem site)."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            tj = json.load(f)
        key = dom if fused_first or (dom + "__first_site_unfused") not in tj else dom + "__first_site_unfused"
        if int(tj.get("_frames_per_launch", 16)) == B and int(tj.get("_cube", 256)) == cube and tj.get(key) is not None:
            roofline["traffic"] = tj.get(key)
            roofline["traffic_source"] = "profiles/traffic.json (%s)" % tj.get("_source", "ncu --set full, B=%d" % B)
    except Exception:
        pass


def host_frames(torch, args, shape, dtype, n):
    """n page-locked host frame batches, allocated the way --host-alloc says."""
    import cp360_b200
    out = []
    for _ in range(n):
        if args.host_alloc == "torch_pin":
            t = torch.empty(shape, dtype=dtype).pin_memory()
        else:
            t = cp360_b200.pinned_empty(shape, dtype, mode=args.host_alloc)
        if dtype == torch.uint8:
            t.copy_(torch.randint(0, 256, shape, dtype=torch.uint8))
        else:
            t.copy_(torch.rand(shape, dtype=torch.float32))
        out.append(t)
    return out


# ---------------------------------------------------------------------------------------------
# B200 arm: chain (headline)
# ---------------------------------------------------------------------------------------------
def run_chain(args, ctx):
    import cp360_b200
    from cp360_b200 import _lib
    torch, dist, dev, world, rank = ctx.torch, ctx.dist, ctx.dev, ctx.world, ctx.rank
    W, K = ctx.W, ctx.K
    B = max(1, args.batch or 32)
    cube = args.cube
    pipe = cp360_b200.SphericalPipeline(EQUI_H, EQUI_W, cube, CAM_C, FEAT_C, device=dev, seed=1234 + rank,
                                        fuse_first_site=not args.no_fuse_first_site)
    pipe.allocate(B)
    frames = pipe.synthetic_frames(B)
    lib = _lib.lib()

    pipe.step(frames)                          # first call of every CubePad site: tiling lookup / tuning happens here
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

    ctx.sampler.start()
    ms = ctx.timed(one_step, profile=args.profile_range)
    value = world * B * K / (ms / 1e3)
    peak, peak_src = measured_peak_gbs()

    # ---- per-launch CUDA events on the launching stream: which kernel dominates, and its GB/s
    prof_steps = min(K, 20)
    acc = event_pass(torch, lambda hook: pipe.step(frames, on_launch=hook), prof_steps)
    per_class, per_site = {}, {}
    for (name, site), (tms, cnt) in acc.items():
        if name == "cubepad":
            C, H, p = pipe.sites[site]
            kname = cubepad_kernel_name(lib, 6 * B, C, H, p)
            nbytes = pipe.cubepad_bytes_per_frame(pipe.sites[site]) * B
            skey = "%s %dx%dx%d" % (kname.split("_kernel")[0], C, H, p)
        elif name == "e2c" and pipe.fuse_first_site:
            kname, nbytes, skey = "e2c_cubepad_kernel", pipe.e2c_cubepad_bytes_per_frame() * B, "e2c_cubepad"
        elif name == "e2c":
            kname, nbytes, skey = "e2c_kernel", pipe.e2c_bytes_per_frame() * B, "e2c"
        else:
            kname, nbytes, skey = "c2e_max_kernel", pipe.c2e_max_bytes_per_frame() * B, "c2e_max"
        ps = per_site.setdefault(skey, {"ms": 0.0, "bytes": nbytes, "n": 0})
        ps["ms"] += tms
        ps["n"] += cnt
        c = per_class.setdefault(kname, {"ms": 0.0, "bytes": 0, "launches": 0})
        c["ms"] += tms
        c["bytes"] += nbytes * cnt
        c["launches"] += cnt
    sites_tbl = {}
    for k, v in per_site.items():
        us = 1e3 * v["ms"] / v["n"]
        sites_tbl[k] = {"us": round(us, 1), "gbs": round(v["bytes"] / (us * 1e-6) / 1e9, 1),
                        "frac": round(v["bytes"] / (us * 1e-6) / 1e9 / peak, 3), "launches_per_step": v["n"] // prof_steps}
    if os.environ.get("CP360_BENCH_SITES"):
        for k, v in sites_tbl.items():
            print("site %-34s %8.1f us %8.1f GB/s %.2f (x%d)" % (k, v["us"], v["gbs"], v["frac"], v["launches_per_step"]), file=sys.stderr)
        if os.environ.get("CP360_BENCH_SITES") == "2":          # where the allocator put each site's tensors (placement experiments)
            for i, (C, H, p) in enumerate(pipe.sites):
                a, b = pipe.site_in[i].data_ptr(), pipe.site_out[i].data_ptr()
                print("ptrs site %d %dx%d in %#x out %#x  in%%2M %7d KB out%%2M %7d KB  (out-in)%%2M %7d KB" % (
                    i, C, H, a, b, (a >> 10) % 2048, (b >> 10) % 2048, ((b - a) >> 10) % 2048), file=sys.stderr)
    chain_gbs = pipe.bytes_per_frame() * B * K / (ms * 1e-3) / 1e9
    roofline, dom = roofline_from(per_class, peak, peak_src,
                                  {"chain_gbs": round(chain_gbs / world, 1), "chain_frac": round(chain_gbs / world / peak, 4)})
    attach_traffic(roofline, dom, B, cube, pipe.fuse_first_site)
    kernels = kernels_table(per_class, prof_steps)

    # ---- the fused chain (SURVEY.md §8 row f2 / north_star (b) "or fused into the producer"), beside the headline
    fused = None
    if not args.no_fused and not args.no_graph:
        g2 = pipe.capture(frames, fused=True)
        b0 = _lib.launch_count()
        pipe.step_fused(frames)
        torch.cuda.synchronize()
        fl = _lib.launch_count() - b0
        ms2 = ctx.timed(g2.replay, tag=False)
        fb, ub = pipe.fused_bytes_per_frame()
        facc = event_pass(torch, lambda hook: pipe.step_fused(frames, on_launch=hook), prof_steps)
        fsites = {}
        for (name, site), (tms, cnt) in facc.items():
            us = 1e3 * tms / cnt
            if name in ("cubepad_bn_relu", "cubepad_cat"):
                C, H, p = pipe.sites[site]
                nb = pipe.cubepad_bytes_per_frame((C, H, p)) * B // (2 if name == "cubepad_cat" else 1)
                key = "%s %dx%dx%d" % (name, C, H, p)
            elif name == "e2c_cubepad":
                w, p0 = cube, pipe.sites[0][2]
                nb = (pipe.e2c_bytes_per_frame() - 6 * w * w * 12 + 6 * 3 * (w + 2 * p0) ** 2 * 4) * B
                key = name
            else:
                nb, key = pipe.c2e_max_bytes_per_frame() * B, name
            fsites[key] = {"us": round(us, 1), "gbs": round(nb / (us * 1e-6) / 1e9, 1), "frac": round(nb / (us * 1e-6) / 1e9 / peak, 3)}
        v2 = world * B * K / (ms2 / 1e3)
        fused = {"value": round(v2, 1), "unit": "frames/s", "ms_per_step": round(ms2 / K, 4), "launches_per_step": int(fl),
                 "algorithmic_bytes_per_frame": fb, "unfused_equivalent_bytes_per_frame": ub,
                 "gbs": round(fb * B * K / (ms2 * 1e-3) / 1e9, 1), "frac": round(fb * B * K / (ms2 * 1e-3) / 1e9 / peak, 4),
                 "unfused_equivalent_gbs": round(ub * B * K / (ms2 * 1e-3) / 1e9, 1),
                 "sites": fsites,
                 "sites_below_half_of_peak": sorted(k for k, v in fsites.items() if v["frac"] < 0.5),
                 "what": "SphericalPipeline.step_fused: e2c+CubePad(3) in one kernel, BN-affine+ReLU folded into the 17 ResNet "
                         "pads, the ConvLSTM site written from its two cat sources; unfused_equivalent = the bytes the unfused "
                         "network moves for the same tensors (own BN+ReLU pass, cat pass, faces round trip)"}
        del g2

    # ---- end to end: page-locked host frames in, host saliency maps out, copies inside the region.
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
            kw = dict(depth=args.e2e_depth, copy_streams=args.e2e_copy_streams)
            pipe.process_host(batches[:4], out_host[:4], **kw)          # warm-up (staging buffers, streams)
            torch.cuda.synchronize()
            ctx.barrier()
            # upload-only ceiling with the same buffers, ring and streams, all ranks at once — taken before AND after
            # the end-to-end run (the uplink is shared with the box's other GPUs, whose tenants come and go)
            c0 = pipe.h2d_probe(batches)
            ctx.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            e0.record()                      # the copy streams wait for the compute stream first, so e0 precedes every H2D
            pipe.process_host(batches, out_host, **kw)
            e1.record()                      # after the last D2H, which runs on the compute stream
            torch.cuda.synchronize()
            wall.append(time.perf_counter() - t0)
            dt = ctx.max_over_ranks(e0.elapsed_time(e1) / 1e3)   # device time (CUDA events), max over ranks
            ctx.barrier()
            c1 = pipe.h2d_probe(batches)
            agg = [world * c[2] / ctx.max_over_ranks(c[1]) / 1e9 for c in (c0, c1)]
            return world * B * E / dt, max(c0[0], c1[0]), max(agg)

        shape = (B, EQUI_H, EQUI_W, 3)
        host_u8 = host_frames(torch, args, shape, torch.uint8, 2)
        v_u8, ceil_u8, ceil_u8_sum = run_e2e(host_u8)
        del host_u8
        host_f32 = host_frames(torch, args, shape, torch.float32, 2)
        v_f32, ceil_f32, _ = run_e2e(host_f32)
        del host_f32
        h2d_bytes = B * EQUI_H * EQUI_W * 3
        agg_gbs = v_u8 * (h2d_bytes / B) / 1e9
        e2e = {"value": round(v_u8, 1), "unit": "frames/s",
               "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": B * 2 * fw * 4 * fw * 4,
               "steps": E, "api": "SphericalPipeline.process_host (page-locked host uint8 frames -> host saliency maps)",
               "timing": "CUDA events on the compute stream around the whole call, max over ranks",
               "wall_clock_value": round(world * B * E / wall[0], 1),
               "h2d_gbs_aggregate": round(agg_gbs, 1),
               "h2d_ceiling_gbs_rank0": round(ceil_u8, 1), "h2d_ceiling_gbs_aggregate": round(ceil_u8_sum, 1),
               "frac_of_h2d_ceiling": round(agg_gbs / ceil_u8_sum, 3),
               "h2d_ceiling_what": "the same page-locked batches through the same staging ring and copy streams, no kernels, "
                                   "all ranks at once (SphericalPipeline.h2d_probe), before and after the end-to-end run "
                                   "(the larger); aggregate = all ranks' bytes / the slowest rank's time: what this box can "
                                   "upload at this GPU count",
               "device_rate_frames_per_s": round(value, 1),
               "host_alloc": args.host_alloc, "staging_depth": args.e2e_depth, "copy_streams": args.e2e_copy_streams,
               "f32_host_frames": {"value": round(v_f32, 1), "h2d_bytes_per_step": h2d_bytes * 4,
                                   "h2d_ceiling_gbs_rank0": round(ceil_f32, 1)}}
    ctx.sampler.stop()
    clocks = ctx.sampler.summary()

    # the single collective of the path: gather every rank's maps (outside the timed region)
    if world > 1:
        maps = cp360_b200.gather_maps(pipe.sal, world * B)
        assert maps.shape[0] == world * B

    cpu = aten = None
    if rank == 0 and world == 1:
        if not args.no_aten_baseline:
            try:
                aten = gpu_aten_baseline(args, dev)
            except Exception as e:                  # noqa: BLE001 - a baseline must never take the bench line down
                aten = {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}
        if not args.no_cpu_baseline:
            cpu = cpu_baseline(args, args.cpu_seconds)

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_string(args), "cube": cube, "frames_per_step_per_gpu": B,
                           "global_frames_per_step": world * B,
                           "sharding": "frames block-partitioned over ranks, no data-path collective",
                           "launch": "CUDA graph replay" if graph is not None else "eager C-ABI launches",
                           "untimed_steps_before_timing": max(W, ctx.min_untimed),
                           "l2": "inputs larger than L2: %.2f GB touched per step per GPU, no flush needed"
                                 % (pipe.bytes_per_frame() * B / 1e9) if pipe.bytes_per_frame() * B > 300e6 else
                                 "%.0f MB touched per step per GPU: part of it stays in the 126 MB L2 between steps "
                                 "(small-batch regime, launch-bound)" % (pipe.bytes_per_frame() * B / 1e6),
                           "algorithmic_bytes_per_frame": pipe.bytes_per_frame()},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * K),
                "roofline": roofline, "kernels": kernels, "sites": sites_tbl, "tuning": tuning_info(lib, 6 * B, pipe.sites),
                "fused_chain": fused, "gpu_aten_baseline": aten, "cpu_baseline": cpu}
        emit(line)


# ---------------------------------------------------------------------------------------------
# B200 arm: ConvLSTM sequence (BASELINE.json configs[3])
# ---------------------------------------------------------------------------------------------
def run_clstm(args, ctx):
    import cp360_b200
    from cp360_b200 import _lib
    torch, dev, world, rank = ctx.torch, ctx.dev, ctx.world, ctx.rank
    W, K = ctx.W, ctx.K
    B = max(1, args.batch or 16)
    c, fw = clstm_dims(args)
    seq = cp360_b200.TemporalCubePadSequence(c, c, fw, SEQ_LEN, device=dev, seed=4321 + rank)
    seq.allocate(B)
    lib = _lib.lib()
    nb = (VIDEO_FRAMES + B - 1) // B                     # window batches per 80-frame video
    seq.window_batch()
    torch.cuda.synchronize()
    b0 = _lib.launch_count()
    seq.window_batch()
    torch.cuda.synchronize()
    launches = (_lib.launch_count() - b0) * nb
    graph = None if args.no_graph else seq.capture()

    def one_step():                                      # one video: 80 output frames (the last batch may be partly padding)
        for _ in range(nb):
            if graph is not None:
                graph.replay()
            else:
                seq.window_batch()

    ctx.sampler.start()
    ms = ctx.timed(one_step, profile=args.profile_range)
    value = world * VIDEO_FRAMES * K / (ms / 1e3)
    peak, peak_src = measured_peak_gbs()
    prof_steps = min(K, 10)
    acc = event_pass(torch, lambda hook: seq.window_batch(on_launch=hook), prof_steps)
    per_class = {}
    sites = seq.sites()
    for (name, site), (tms, cnt) in acc.items():
        if name in ("cubepad", "cubepad_cat"):
            C, H, p = sites[site]
            kname = cubepad_kernel_name(lib, 6 * B, C if name == "cubepad" else C // 2, H, p)
            nbytes = 6 * B * C * (H * H + (H + 2) * (H + 2)) * 4 // (2 if name == "cubepad_cat" else 1)
        else:
            kname, nbytes = "c2e_max_kernel", (6 * c * fw * fw * 4 + 8 * fw * fw * 4) * B
        cl = per_class.setdefault(kname, {"ms": 0.0, "bytes": 0, "launches": 0})
        cl["ms"] += tms
        cl["bytes"] += nbytes * cnt
        cl["launches"] += cnt
    processed = nb * B                                   # windows actually computed per video (>= 80)
    chain_gbs = seq.bytes_per_frame() * processed * K / (ms * 1e-3) / 1e9
    roofline, dom = roofline_from(per_class, peak, peak_src, {"sequence_gbs": round(chain_gbs / world, 1),
                                                              "sequence_frac": round(chain_gbs / world / peak, 4)})

    e2e = None
    if not args.no_e2e:
        # host side of test_temporal.py:60-88: the window's cube feature frames come from host memory
        # (np.load -> FloatTensor -> .cuda()), the equirect map goes back to the host (.cpu().numpy()).
        E = max(2, min(K, 10))
        feat_shape = (SEQ_LEN, 6 * B, c, fw, fw)
        host = [torch.randn(feat_shape, dtype=torch.float32).pin_memory() for _ in range(2)]
        out_host = torch.empty((E * nb, B, 2 * fw, 4 * fw), dtype=torch.float32).pin_memory()
        stage = [torch.empty(feat_shape, dtype=torch.float32, device=dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        compute = torch.cuda.current_stream(dev)

        def run(n_batches):
            freed = [None, None]
            copy_stream.wait_stream(compute)
            for i in range(n_batches):
                k = i & 1
                with torch.cuda.stream(copy_stream):
                    if freed[k] is not None:
                        copy_stream.wait_event(freed[k])
                    stage[k].copy_(host[k], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                compute.wait_event(ev)
                seq.x = list(stage[k].unbind(0))
                sal = seq.window_batch()
                freed[k] = torch.cuda.Event()
                freed[k].record(compute)
                out_host[i].copy_(sal, non_blocking=True)
            compute.synchronize()
        run(2)
        torch.cuda.synchronize()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(E * nb)
        e1.record()
        torch.cuda.synchronize()
        dt = ctx.max_over_ranks(e0.elapsed_time(e1) / 1e3)
        seq.allocate(B)
        e2e = {"value": round(world * VIDEO_FRAMES * E / dt, 1), "unit": "frames/s",
               "h2d_bytes_per_step": nb * SEQ_LEN * 6 * B * c * fw * fw * 4, "d2h_bytes_per_step": nb * B * 2 * fw * 4 * fw * 4,
               "steps": E, "api": "TemporalCubePadSequence.window_batch fed from pinned host cube-feature windows "
                                  "(test_temporal.py:60-88), maps back to the host",
               "timing": "CUDA events around the whole loop, max over ranks"}
    ctx.sampler.stop()
    clocks = ctx.sampler.summary()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, args.cpu_seconds)
    if rank == 0:
        line = {"metric": METRIC.replace("equi->cube->CubePad->cube->equi", "ConvLSTM-side CubePad x3 per step + cube->equi"),
                "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_string(args), "windows_per_launch": B, "window_batches_per_video": nb,
                           "frames_per_step_per_gpu": VIDEO_FRAMES, "seq_len": SEQ_LEN, "variant": args.clstm_variant,
                           "sharding": "videos over ranks (one video per rank per step), no data-path collective",
                           "launch": "CUDA graph replay" if graph is not None else "eager C-ABI launches",
                           "untimed_steps_before_timing": max(W, ctx.min_untimed),
                           "l2": "%.2f GB touched per window batch, larger than L2" % (seq.bytes_per_frame() * B / 1e9),
                           "algorithmic_bytes_per_frame": seq.bytes_per_frame()},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches * K), "roofline": roofline,
                "kernels": kernels_table(per_class, prof_steps), "tuning": tuning_info(lib, 6 * B, sites),
                "cpu_baseline": cpu}
        emit(line)


# ---------------------------------------------------------------------------------------------
# B200 arm: the 2 000-frame corpus, strong scaling (BASELINE.json configs[4])
# ---------------------------------------------------------------------------------------------
def run_corpus(args, ctx):
    import cp360_b200
    from cp360_b200 import _lib
    torch, dev, world, rank = ctx.torch, ctx.dev, ctx.world, ctx.rank
    W, K = ctx.W, ctx.K
    a, b = cp360_b200.shard_range(CORPUS_FRAMES, rank, world)
    n_local = b - a
    B = args.batch or max(d for d in range(1, 41) if n_local % d == 0)
    pipe = cp360_b200.SphericalPipeline(EQUI_H, EQUI_W, args.cube, CAM_C, FEAT_C, device=dev, seed=1234,
                                        fuse_first_site=not args.no_fuse_first_site)
    pipe.allocate(B)
    lib = _lib.lib()
    # the rank's share of the corpus, resident in HBM (44 GB of fp32 frames on one GPU, SURVEY.md §8d-5: frames are
    # generated on-device from the seed, no disk / PCIe in the timed loop)
    corpus = torch.empty((n_local, EQUI_H, EQUI_W, 3), dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(99)
    for i0 in range(0, n_local, 25):
        g.manual_seed(99 + a + i0)                       # frame content depends on the global index only
        corpus[i0:i0 + 25].copy_(torch.rand((min(25, n_local - i0), EQUI_H, EQUI_W, 3), dtype=torch.float32, device=dev, generator=g))
    maps = torch.empty((n_local, 2 * pipe.feat_w, 4 * pipe.feat_w), dtype=torch.float32, device=dev)
    nbatches = (n_local + B - 1) // B
    st = {"launches": 0}

    def one_pass():
        for i in range(nbatches):
            lo = min(i * B, n_local - B)                 # a ragged tail re-covers the last B frames
            sal = pipe.step(corpus[lo:lo + B])
            maps[lo:lo + B].copy_(sal)
    one_pass()
    torch.cuda.synchronize()
    b0 = _lib.launch_count()
    one_pass()
    torch.cuda.synchronize()
    st["launches"] = _lib.launch_count() - b0
    ctx.sampler.start()
    ms = ctx.timed(one_pass, profile=args.profile_range)
    value = CORPUS_FRAMES * K / (ms / 1e3)
    peak, peak_src = measured_peak_gbs()
    # the one collective of the path, timed on its own (NCCL all_gather of [n_local,16,32] maps)
    gather_ms = None
    if world > 1:
        for _ in range(3):
            allm = cp360_b200.gather_maps(maps, CORPUS_FRAMES)
        torch.cuda.synchronize()
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            allm = cp360_b200.gather_maps(maps, CORPUS_FRAMES)
        e1.record()
        torch.cuda.synchronize()
        gather_ms = ctx.max_over_ranks(e0.elapsed_time(e1) / 10)
        assert allm.shape[0] == CORPUS_FRAMES
    prof = event_pass(torch, lambda hook: pipe.step(corpus[:B], on_launch=hook), 5)
    per_class = {}
    for (name, site), (tms, cnt) in prof.items():
        if name == "cubepad":
            C, H, p = pipe.sites[site]
            kname, nbytes = cubepad_kernel_name(lib, 6 * B, C, H, p), pipe.cubepad_bytes_per_frame(pipe.sites[site]) * B
        elif name == "e2c" and pipe.fuse_first_site:
            kname, nbytes = "e2c_cubepad_kernel", pipe.e2c_cubepad_bytes_per_frame() * B
        elif name == "e2c":
            kname, nbytes = "e2c_kernel", pipe.e2c_bytes_per_frame() * B
        else:
            kname, nbytes = "c2e_max_kernel", pipe.c2e_max_bytes_per_frame() * B
        cl = per_class.setdefault(kname, {"ms": 0.0, "bytes": 0, "launches": 0})
        cl["ms"] += tms
        cl["bytes"] += nbytes * cnt
        cl["launches"] += cnt
    gbs = pipe.bytes_per_frame() * n_local * K / (ms * 1e-3) / 1e9
    roofline, dom = roofline_from(per_class, peak, peak_src, {"chain_gbs_per_gpu": round(gbs, 1), "chain_frac": round(gbs / peak, 4)})
    ctx.sampler.stop()
    clocks = ctx.sampler.summary()
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, args.cpu_seconds)
    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 1), "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": round(ms / K, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_string(args), "cube": args.cube, "corpus_frames": CORPUS_FRAMES,
                           "frames_per_rank": n_local, "frames_per_launch": B,
                           "sharding": "contiguous frame blocks over ranks (shard_range), no collective inside a pass",
                           "launch": "eager C-ABI launches",
                           "l2": "every pass streams %.1f GB of frames + %.1f GB of features per GPU" %
                                 (n_local * EQUI_H * EQUI_W * 12 / 1e9, (pipe.bytes_per_frame() - (pipe.e2c_cubepad_bytes_per_frame() if pipe.fuse_first_site else pipe.e2c_bytes_per_frame())) * n_local / 1e9),
                           "algorithmic_bytes_per_frame": pipe.bytes_per_frame()},
                "corpus_wall_ms": round(ms / K, 3), "gather_ms": None if gather_ms is None else round(gather_ms, 4),
                "gather": None if gather_ms is None else "all_gather_into_tensor of [%d,%d,%d] fp32 maps (NCCL), timed on its own over 10 calls"
                          % (n_local, 2 * pipe.feat_w, 4 * pipe.feat_w),
                "clocks": clocks, "e2e": None, "gpu_launches": int(st["launches"] * K), "roofline": roofline,
                "kernels": kernels_table(per_class, 5), "tuning": tuning_info(lib, 6 * B, pipe.sites), "cpu_baseline": cpu}
        emit(line)


def run_b200(args, rank, world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    ctx = Ctx(args, rank, world, local_rank)
    {"chain": run_chain, "clstm": run_clstm, "corpus": run_corpus}[args.workload](args, ctx)
    if world > 1:
        ctx.dist.destroy_process_group()


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
