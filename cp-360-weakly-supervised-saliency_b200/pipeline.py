"""SphericalPipeline — the batched, frame-sharded chain the benchmark measures
(BASELINE.json metric: frames/sec equi->cube->CubePad->cube->equi @1920x960).

One *step* processes B frames that are independent of each other, in the order
the reference's static-inference loop touches the three operators
(static_model/dataset_feat_extractor.py:145,161,174 + model/resnet_cubic.py:165-170,92):

  K1  Equi2Cube            frames[B,Hin,Win,3]  -> faces[6B,3,w,w]                (e2c.cu)
  K2  CubePad x 18         the 18 CubePad sites of one cubic ResNet-50 forward; site 0 pads the
                           faces K1 just produced, sites 1..17 pad device-resident feature tensors
                           of the exact site shapes (the convolutions between them are cuDNN's
                           business and out of scope, so the features are synthetic stand-ins)
  K2  CubePad(1) 2048-ch   [6B,2048,w/32,w/32]  -> [6B,2048,w/32+2,w/32+2]  (BASELINE's 2048-channel
                           CubePad; the ConvLSTM-side site, model/clstm.py:58-64)
  K3m Cube2Equi + max      cam[6B,1000,w/32,w/32] -> sal[B,2w/32,4w/32]            (c2e.cu)

All launches go through the C-ABI (include/cp360.h) on the caller's current stream; buffers are
allocated once per (B, device) and reused; the step can be captured in a CUDA graph.
Frames are partitioned over ranks in contiguous blocks; no collective is needed inside a step —
``gather_maps`` is the single final exchange.
"""
import numpy as np
import torch

from . import _lib
from .cube_to_equi import Cube2Equi
from .equi_to_cube import Equi2Cube


def resnet50_cubepad_sites(cube):
    """(C, H, pad) of the 18 CubePad calls of one cubic ResNet-50 forward at face width `cube`
    (model/resnet_cubic.py:71,92,116-117,165,169)."""
    d = int(cube)
    return ([(3, d, 3), (64, d // 2, 1)] + [(64, d // 4, 1)] * 3 + [(128, d // 4, 1)] +
            [(128, d // 8, 1)] * 3 + [(256, d // 8, 1)] + [(256, d // 16, 1)] * 5 +
            [(512, d // 16, 1)] + [(512, d // 32, 1)] * 2)


def shard_range(n_items, rank, world):
    """Contiguous block partition of n_items over `world` ranks -> (start, stop) of `rank`."""
    base, rem = divmod(int(n_items), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_maps(local_maps, n_total, group=None):
    """The one collective of the path: every rank contributes its [n_local,h,w] saliency maps,
    every rank gets the [n_total,h,w] stack in frame order (NCCL on GPU tensors, gloo on CPU)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_maps
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    n_max = max(b - a for a, b in sizes)
    pad = torch.zeros((n_max,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype, device=local_maps.device)
    pad[:local_maps.shape[0]] = local_maps
    out = torch.empty((world * n_max,) + tuple(local_maps.shape[1:]), dtype=local_maps.dtype,
                      device=local_maps.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * n_max:r * n_max + (b - a)] for r, (a, b) in enumerate(sizes)], 0)


def gpu_numa_node(device):
    """NUMA node the GPU's PCIe function hangs off (sysfs), or -1 when the platform does not say."""
    props = torch.cuda.get_device_properties(device)
    try:
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as f:
            return int(f.read().strip())
    except (AttributeError, OSError, ValueError):
        return -1


def prefer_gpu_numa_node(device):
    """Best effort, for multi-GPU hosts: make this process PREFER (MPOL_PREFERRED — falls back to other nodes,
    never fails an allocation) host memory on the GPU's own NUMA node, so that pinned staging buffers allocated
    afterwards (``process_host``'s host batches) are read by the GPU's copy engines without crossing the socket
    interconnect. One process per GPU is the deployment model, so a process-wide policy is the
    right grain. Returns a small dict saying what was done; never raises."""
    import ctypes
    import os
    info = {"node": -1, "applied": False}
    try:
        node = gpu_numa_node(device)
        info["node"] = node
        if node < 0 or not os.path.isdir("/sys/devices/system/node/node%d" % node):
            info["why"] = "GPU NUMA node unknown"
            return info
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        SYS_set_mempolicy, MPOL_PREFERRED = 238, 1          # x86_64
        if os.uname().machine != "x86_64":
            info["why"] = "set_mempolicy syscall number only known for x86_64"
            return info
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), 16 * 64)
        if rc != 0:
            info["why"] = "set_mempolicy: " + os.strerror(ctypes.get_errno())
            return info
        info["applied"] = True
    except Exception as e:                                  # noqa: BLE001 - advisory only
        info["why"] = "%s: %s" % (type(e).__name__, e)
    return info


class SphericalPipeline:
    def __init__(self, equi_h=960, equi_w=1920, cube=256, cam_channels=1000, feat_channels=2048,
                 device=None, seed=1234, fuse_first_site=False, align_corners=False):
        """fuse_first_site: run e2c and the CubePad(3) in front of conv1 as ONE kernel
        (cp360_e2c_cubepad_fwd); the unpadded faces are then never written.
        align_corners: grid_sample convention of the back-projection (False = what the unmodified reference
        computes under the installed torch; True = the torch<=1.2 behaviour it was written against)."""
        self.fuse_first_site = bool(fuse_first_site)
        if not torch.cuda.is_available():
            raise RuntimeError("SphericalPipeline needs a CUDA device (sm_100a); no CPU fallback")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.equi_h, self.equi_w, self.cube = int(equi_h), int(equi_w), int(cube)
        self.feat_w = self.cube // 32
        self.cam_channels, self.feat_channels = int(cam_channels), int(feat_channels)
        self.e2c = Equi2Cube(self.cube, np.empty((self.equi_h, self.equi_w, 3), np.float32))
        self.c2e = Cube2Equi(self.feat_w, align_corners=align_corners)
        self.sites = resnet50_cubepad_sites(self.cube) + [(self.feat_channels, self.feat_w, 1)]
        self.seed = seed
        self.B = 0
        self._lib = _lib.lib()

    # ---------------------------------------------------------------- accounting
    def _e2c_touched(self):
        sx, sy = self.e2c.sx.astype(np.int64).reshape(-1), self.e2c.sy.astype(np.int64).reshape(-1)
        x0, y0, fx, fy = sx >> 5, sy >> 5, sx & 31, sy & 31
        W = self.equi_w
        t = [y0 * W + x0, (y0 * W + x0 + 1)[fx > 0], ((y0 + 1) * W + x0)[fy > 0],
             ((y0 + 1) * W + x0 + 1)[(fx > 0) & (fy > 0)]]
        return np.unique(np.concatenate(t))

    def e2c_bytes_per_frame(self):
        """Algorithmic bytes of K1: unique input pixels touched * C * 4 + faces written."""
        if getattr(self, "_e2c_bytes", None) is None:
            w = self.cube
            self._e2c_bytes = int(self._e2c_touched().size) * 3 * 4 + 6 * w * w * 3 * 4
        return self._e2c_bytes

    def e2c_sector_bytes_per_frame(self, elem=4):
        """What DRAM must deliver at its 32 B access granularity: every sector holding a touched pixel once,
        plus the faces written — the physical floor of a gather that cannot fetch less than a sector."""
        px = self._e2c_touched() * 3 * elem                       # byte offset of each touched pixel (3 channels)
        sectors = np.unique(np.concatenate([px // 32, (px + 3 * elem - 1) // 32]))
        w = self.cube
        return int(sectors.size) * 32 + 6 * w * w * 3 * 4

    def cubepad_bytes_per_frame(self, site):
        C, H, p = site
        return 6 * C * (H * H + (H + 2 * p) * (H + 2 * p)) * 4

    def c2e_max_bytes_per_frame(self):
        w = self.feat_w
        return 6 * self.cam_channels * w * w * 4 + 8 * w * w * 4

    def e2c_cubepad_bytes_per_frame(self):
        """Algorithmic bytes of the fused first site (cp360_e2c_cubepad_fwd): unique input pixels touched + the
        padded conv1 input written; the unpadded faces are neither written nor read."""
        w, p0 = self.cube, self.sites[0][2]
        return self.e2c_bytes_per_frame() - 6 * w * w * 3 * 4 + 6 * 3 * (w + 2 * p0) ** 2 * 4

    def bytes_per_frame(self):
        if self.fuse_first_site:
            return (self.e2c_cubepad_bytes_per_frame() + sum(self.cubepad_bytes_per_frame(s) for s in self.sites[1:]) +
                    self.c2e_max_bytes_per_frame())
        return (self.e2c_bytes_per_frame() + sum(self.cubepad_bytes_per_frame(s) for s in self.sites) +
                self.c2e_max_bytes_per_frame())

    def fused_bytes_per_frame(self):
        """(algorithmic bytes of the fused chain, bytes the UNFUSED network moves for the same tensors).

        Fused chain (step_fused): e2c + CubePad(3) in one kernel (the faces are never written), BN-affine + ReLU
        folded into the pad at the 17 ResNet sites, the ConvLSTM site written from its two cat sources.
        The unfused network pays for the same results: e2c + pad (faces written, then read), a BN+ReLU pass of
        its own per site (read x, write x': resnet_cubic.py:89-92) before the pad reads x', and a torch.cat pass
        (read both sources, write the cat: clstm.py:57) before the pad reads the cat."""
        w, p0 = self.cube, self.sites[0][2]
        faces = 6 * w * w * 3 * 4
        touched = self.e2c_bytes_per_frame() - faces
        padded0 = 6 * 3 * (w + 2 * p0) ** 2 * 4
        fused = touched + padded0
        unfused = touched + faces + faces + padded0
        for (C, H, p) in self.sites[1:]:
            x, y = 6 * C * H * H * 4, 6 * C * (H + 2 * p) ** 2 * 4
            fused += x + y
            unfused += 3 * x + y
        fused += self.c2e_max_bytes_per_frame()
        unfused += self.c2e_max_bytes_per_frame()
        return fused, unfused

    # ---------------------------------------------------------------- buffers
    def allocate(self, B):
        """Device buffers for B frames per step; synthetic resident features (seeded)."""
        dev, w = self.device, self.cube
        g = torch.Generator(device=dev).manual_seed(self.seed)
        self.B = int(B)
        n = 6 * self.B
        self.faces = torch.empty((n, 3, w, w), dtype=torch.float32, device=dev)
        self.site_in, self.site_out = [self.faces], []
        for i, (C, H, p) in enumerate(self.sites):
            if i > 0:
                self.site_in.append(torch.randn((n, C, H, H), dtype=torch.float32, device=dev, generator=g))
            self.site_out.append(torch.empty((n, C, H + 2 * p, H + 2 * p), dtype=torch.float32, device=dev))
        fw = self.feat_w
        self.cam = torch.randn((n, self.cam_channels, fw, fw), dtype=torch.float32, device=dev, generator=g)
        self.sal = torch.empty((self.B, 2 * fw, 4 * fw), dtype=torch.float32, device=dev)
        self._packed = self.e2c._map_on(dev)
        self._taps, self._wts = self.c2e._plan_on(dev)
        # fused chain (step_fused): folded eval-mode BatchNorm parameters per ResNet site, and the ConvLSTM
        # site's input as the two tensors the reference concatenates (input_, h_cur: clstm.py:57)
        self.bn_scale, self.bn_shift = [None], [None]
        for (C, H, p) in self.sites[1:-1]:
            self.bn_scale.append(torch.rand(C, dtype=torch.float32, device=dev, generator=g) + 0.5)
            self.bn_shift.append(torch.randn(C, dtype=torch.float32, device=dev, generator=g))
        Cf, Hf, _ = self.sites[-1]
        half = Cf // 2
        self.cat_src = [self.site_in[-1][:, :half].contiguous(), self.site_in[-1][:, half:].contiguous()]
        return self

    def synthetic_frames(self, B, generator=None):
        """U[0,1) fp32 frames [B,Hin,Win,3] on the device (Wild-360-shaped)."""
        g = generator or torch.Generator(device=self.device).manual_seed(self.seed + 1)
        return torch.rand((B, self.equi_h, self.equi_w, 3), dtype=torch.float32, device=self.device, generator=g)

    # ---------------------------------------------------------------- the step
    def launches_per_step(self):
        return 1 + len(self.sites) + 1 - int(self.fuse_first_site)   # e2c, CubePads, c2e + channel max (one cluster kernel)

    def step(self, frames, on_launch=None):
        """frames [B,Hin,Win,3] fp32 (or uint8) on self.device -> sal [B,2fw,4fw] (buffer reused each step).

        on_launch(name, site_index): optional hook called before every C-ABI call and once after
        the last (bench.py uses it to drop CUDA events between kernels)."""
        if frames.shape[0] != self.B:
            self.allocate(frames.shape[0])
        lib, chk = self._lib, _lib.check
        st = torch.cuda.current_stream(self.device).cuda_stream
        n = 6 * self.B
        if on_launch:
            on_launch("e2c", -1)
        first = 0
        if self.fuse_first_site:
            p = self.sites[0][2]
            chk(lib.cp360_e2c_cubepad_fwd(frames.data_ptr(), int(frames.dtype == torch.uint8), self._packed.data_ptr(),
                                          self.site_out[0].data_ptr(), self.B, self.equi_h, self.equi_w, 3, self.cube,
                                          p, p, p, p, 255.0, None, None, st))
            first = 1
        elif frames.dtype == torch.uint8:    # decoded video frames: converted as float32(u8)/255 on the fly
            chk(lib.cp360_e2c_fwd_u8(frames.data_ptr(), self._packed.data_ptr(), self.faces.data_ptr(), self.B,
                                     self.equi_h, self.equi_w, 3, self.cube, _lib.LAYOUT_NCHW, 255.0, None, None, st))
        else:
            chk(lib.cp360_e2c_fwd(frames.data_ptr(), self._packed.data_ptr(), self.faces.data_ptr(), self.B,
                                  self.equi_h, self.equi_w, 3, self.cube, _lib.LAYOUT_NCHW, None, None, st))
        for i, (C, H, p) in enumerate(self.sites):
            if i < first:
                continue
            if on_launch:
                on_launch("cubepad", i)
            chk(lib.cp360_cubepad_fwd(self.site_in[i].data_ptr(), self.site_out[i].data_ptr(), n, C, H, H,
                                      p, p, p, p, 4, st))
        if on_launch:
            on_launch("c2e_max", -1)
        chk(lib.cp360_c2e_max_fwd(self.cam.data_ptr(), self._taps.data_ptr(), self._wts.data_ptr(),
                                  self.sal.data_ptr(), self.B, self.cam_channels, self.feat_w, st))
        if on_launch:
            on_launch("end", -1)
        return self.sal

    def step_fused(self, frames, on_launch=None):
        """The same chain with every producer-side fusion the library offers (north_star's
        "fused into the producer of the conv input so no separate pad copy exists"):
          cp360_e2c_cubepad_fwd            frames -> CubePad(3)(faces), the faces are never written
          cp360_cubepad_fused_fwd x 17     CubePad(relu(x * scale[c] + shift[c])) — BN + ReLU + pad in one pass
          cp360_cubepad_fused_fwd x 2      the ConvLSTM site as CubePad(cat(a, b)) written one source at a time
          cp360_c2e_max_fwd                back-projection + channel max
        Site outputs equal CubePad applied to the affine+ReLU'd / concatenated inputs bit for bit
        (tests/test_pipeline_gpu.py::test_step_fused_equals_pad_of_the_producer_ops)."""
        if frames.shape[0] != self.B:
            self.allocate(frames.shape[0])
        lib, chk = self._lib, _lib.check
        st = torch.cuda.current_stream(self.device).cuda_stream
        n = 6 * self.B
        if on_launch:
            on_launch("e2c_cubepad", 0)
        p = self.sites[0][2]
        chk(lib.cp360_e2c_cubepad_fwd(frames.data_ptr(), int(frames.dtype == torch.uint8), self._packed.data_ptr(),
                                      self.site_out[0].data_ptr(), self.B, self.equi_h, self.equi_w, 3, self.cube,
                                      p, p, p, p, 255.0, None, None, st))
        last = len(self.sites) - 1
        for i, (C, H, p) in enumerate(self.sites):
            if i == 0:
                continue
            if i < last:
                if on_launch:
                    on_launch("cubepad_bn_relu", i)
                chk(lib.cp360_cubepad_fused_fwd(self.site_in[i].data_ptr(), self.site_out[i].data_ptr(), n, C, H, H,
                                                p, p, p, p, self.bn_scale[i].data_ptr(), self.bn_shift[i].data_ptr(),
                                                1, 0, 0, st))
            else:
                off = 0
                for src in self.cat_src:
                    if on_launch:
                        on_launch("cubepad_cat", i)
                    chk(lib.cp360_cubepad_fused_fwd(src.data_ptr(), self.site_out[i].data_ptr(), n, src.shape[1], H, H,
                                                    p, p, p, p, None, None, 0, C, off, st))
                    off += src.shape[1]
        if on_launch:
            on_launch("c2e_max", -1)
        chk(lib.cp360_c2e_max_fwd(self.cam.data_ptr(), self._taps.data_ptr(), self._wts.data_ptr(),
                                  self.sal.data_ptr(), self.B, self.cam_channels, self.feat_w, st))
        if on_launch:
            on_launch("end", -1)
        return self.sal

    def capture(self, frames, fused=False):
        """Capture one step over `frames` (a fixed device buffer) in a CUDA graph; returns the
        graph (call .replay()). The first eager step doubles as warm-up (function attributes)."""
        fn = self.step_fused if fused else self.step
        fn(frames)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn(frames)
        return graph

    # ---------------------------------------------------------------- host-buffer entry point
    def run_host(self, frames_host, out_host=None):
        """End-to-end call with HOST buffers: frames_host [B,Hin,Win,3] fp32 (pinned for async
        copies) -> saliency maps [B,2fw,4fw] on the host. H2D + chain + D2H on the current stream."""
        frames = frames_host.to(self.device, non_blocking=True)
        sal = self.step(frames)
        if out_host is None:
            return sal.cpu()
        out_host.copy_(sal, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out_host

    def process_host(self, host_batches, out_host, depth=3, copy_streams=1, fused=False):
        """Streamed end-to-end path: host_batches is a sequence of pinned [B,Hin,Win,3] host tensors
        (uint8 video frames or float32), out_host a pinned [len(host_batches),B,2fw,4fw] tensor. Uploads run on
        `copy_streams` copy streams (each batch split evenly between them) into a ring of `depth` device buffers,
        up to depth-1 batches ahead of the chain; every batch's maps are copied back to the host on the compute
        stream. Returns after everything has landed."""
        dev = self.device
        B = host_batches[0].shape[0]
        if B != self.B:
            self.allocate(B)
        dt = host_batches[0].dtype
        depth, copy_streams = max(2, int(depth)), max(1, int(copy_streams))
        st = getattr(self, "_stage", None)
        if st is None or len(st) != depth or st[0].shape[0] != B or st[0].dtype != dt or len(self._copy_streams) != copy_streams:
            self._stage = [torch.empty((B, self.equi_h, self.equi_w, 3), dtype=dt, device=dev) for _ in range(depth)]
            self._copy_streams = [torch.cuda.Stream(device=dev) for _ in range(copy_streams)]
        compute = torch.cuda.current_stream(dev)
        freed = [None] * depth
        part = (B + copy_streams - 1) // copy_streams
        for cs in self._copy_streams:
            cs.wait_stream(compute)
        run = self.step_fused if fused else self.step
        for i, hb in enumerate(host_batches):
            k = i % depth
            for j, cs in enumerate(self._copy_streams):
                with torch.cuda.stream(cs):
                    if freed[k] is not None:
                        cs.wait_event(freed[k])
                    self._stage[k][j * part:(j + 1) * part].copy_(hb[j * part:(j + 1) * part], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                compute.wait_event(ev)
            sal = run(self._stage[k])
            freed[k] = torch.cuda.Event()
            freed[k].record(compute)
            out_host[i].copy_(sal, non_blocking=True)
        compute.synchronize()
        return out_host

    def h2d_probe(self, host_batches, repeats=1):
        """Upload-only ceiling of process_host: the same pinned batches through the same staging ring and copy
        streams, no kernels, no D2H. Returns (GB/s, seconds, bytes) — CUDA events on the compute stream; with several
        ranks on one host the aggregate is all ranks' bytes over the SLOWEST rank's seconds (ranks that finish early
        hand their share of the uplink to the others, so per-rank rates must not be summed)."""
        dev = self.device
        if getattr(self, "_stage", None) is None:
            raise RuntimeError("call process_host once first (it owns the staging ring)")
        compute = torch.cuda.current_stream(dev)
        B = host_batches[0].shape[0]
        part = (B + len(self._copy_streams) - 1) // len(self._copy_streams)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(compute)
        for cs in self._copy_streams:
            cs.wait_event(e0)
        nbytes = 0
        for _ in range(repeats):
            for i, hb in enumerate(host_batches):
                k = i % len(self._stage)
                for j, cs in enumerate(self._copy_streams):
                    with torch.cuda.stream(cs):
                        self._stage[k][j * part:(j + 1) * part].copy_(hb[j * part:(j + 1) * part], non_blocking=True)
                nbytes += hb.numel() * hb.element_size()
        for cs in self._copy_streams:
            compute.wait_stream(cs)
        e1.record(compute)
        compute.synchronize()
        sec = e0.elapsed_time(e1) * 1e-3
        return nbytes / sec / 1e9, sec, nbytes

class TemporalCubePadSequence:
    """The hot-path work of the ConvLSTM temporal model (BASELINE.json configs[3]) for B
    windows at a time: per time step the three CubePads of one cell evaluation (model/clstm.py:57-64)

        CubePad(cat(input_, h_cur))   [6B, feat+hid, w, w]   written from its two sources (cubepad_cat)
        CubePad(relu(Conv1(.)))       [6B, 4*hid,   w, w]
        CubePad(relu(Conv2(.)))       [6B, 4*hid,   w, w]

    and, after `seq_len` steps, the back-projection + channel max of the hidden state
    (temporal_model/test_temporal.py:82-84). Windows are independent — the state is reset for every output
    frame (test_temporal.py:69-73) — so they batch along the cube dimension. The three convolutions and the gate
    math are cuDNN's / PyTorch's business (out of scope): their outputs are device-resident stand-ins of the
    exact shapes. feat = hid = 2048 on 8x8 faces is BASELINE's synthetic width; feat = hid = 1000 on 7x7
    faces is what the reference runs (config.yaml:21-22)."""

    def __init__(self, feat_channels=2048, hidden_channels=2048, feat_w=8, seq_len=5, device=None, seed=4321,
                 align_corners=False, fused_cat=True):
        if not torch.cuda.is_available():
            raise RuntimeError("TemporalCubePadSequence needs a CUDA device (sm_100a); no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.feat, self.hid, self.w, self.seq_len = int(feat_channels), int(hidden_channels), int(feat_w), int(seq_len)
        self.fused_cat = bool(fused_cat)
        self.c2e = Cube2Equi(self.w, align_corners=align_corners)
        self.seed = seed
        self.B = 0
        self._lib = _lib.lib()

    def sites(self):
        return [(self.feat + self.hid, self.w, 1), (4 * self.hid, self.w, 1), (4 * self.hid, self.w, 1)]

    def bytes_per_frame(self):
        """Algorithmic bytes per OUTPUT frame (= one window): seq_len x the three pads + one c2e+max."""
        w = self.w
        pads = sum(6 * C * (w * w + (w + 2) * (w + 2)) * 4 for C, _, _ in self.sites())
        return self.seq_len * pads + 6 * self.hid * w * w * 4 + 8 * w * w * 4

    def launches_per_window_batch(self):
        return self.seq_len * (4 if self.fused_cat else 3) + 2      # + fill and c2e_max

    def allocate(self, B):
        dev, w = self.device, self.w
        g = torch.Generator(device=dev).manual_seed(self.seed)
        self.B = int(B)
        n = 6 * self.B
        r = lambda c: torch.randn((n, c, w, w), dtype=torch.float32, device=dev, generator=g)   # noqa: E731
        # one set of stand-ins PER TIME STEP (as in the model, where every step's tensors are fresh conv outputs):
        # nothing a step reads was touched by the previous step, so no input is served from L2 by accident
        self.x = [r(self.feat) for _ in range(self.seq_len)]         # the window's feature frames
        self.hs = [r(self.hid) for _ in range(self.seq_len + 1)]     # hidden state before each step / after the last
        self.mid = [[r(4 * self.hid), r(4 * self.hid)] for _ in range(self.seq_len)]   # Conv1 / Conv2 outputs
        self.h = self.hs[-1]
        self.cat = torch.empty((n, self.feat + self.hid, w, w), dtype=torch.float32, device=dev) if not self.fused_cat else None
        self.out_cat = torch.empty((n, self.feat + self.hid, w + 2, w + 2), dtype=torch.float32, device=dev)
        self.out_mid = [torch.empty((n, 4 * self.hid, w + 2, w + 2), dtype=torch.float32, device=dev) for _ in range(2)]
        self.sal = torch.empty((self.B, 2 * w, 4 * w), dtype=torch.float32, device=dev)
        self._taps, self._wts = self.c2e._plan_on(dev)
        return self

    def window_batch(self, on_launch=None):
        """One batch of B windows: seq_len cell evaluations' CubePads, then c2e + max of the hidden state."""
        lib, chk = self._lib, _lib.check
        st = torch.cuda.current_stream(self.device).cuda_stream
        n, w = 6 * self.B, self.w
        for t in range(self.seq_len):
            if self.fused_cat:
                off = 0
                for src in (self.x[t], self.hs[t]):
                    if on_launch:
                        on_launch("cubepad_cat", 0)
                    chk(lib.cp360_cubepad_fused_fwd(src.data_ptr(), self.out_cat.data_ptr(), n, src.shape[1], w, w,
                                                    1, 1, 1, 1, None, None, 0, self.feat + self.hid, off, st))
                    off += src.shape[1]
            else:
                torch.cat((self.x[t], self.hs[t]), 1, out=self.cat)
                if on_launch:
                    on_launch("cubepad", 0)
                chk(lib.cp360_cubepad_fwd(self.cat.data_ptr(), self.out_cat.data_ptr(), n, self.feat + self.hid, w, w,
                                          1, 1, 1, 1, 4, st))
            for j in range(2):
                if on_launch:
                    on_launch("cubepad", 1 + j)
                chk(lib.cp360_cubepad_fwd(self.mid[t][j].data_ptr(), self.out_mid[j].data_ptr(), n, 4 * self.hid, w, w,
                                          1, 1, 1, 1, 4, st))
        if on_launch:
            on_launch("c2e_max", -1)
        chk(lib.cp360_c2e_max_fwd(self.h.data_ptr(), self._taps.data_ptr(), self._wts.data_ptr(),
                                  self.sal.data_ptr(), self.B, self.hid, w, st))
        if on_launch:
            on_launch("end", -1)
        return self.sal

    def capture(self):
        self.window_batch()
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.window_batch()
        return graph
