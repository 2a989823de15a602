#!/usr/bin/env python
"""Build container only: time the UNMODIFIED reference (executed where it lies under /root/reference, through
tests/golden/_ref_loader.py) on the benchmark chain, next to bench.py's CPU port (oracle/ref_port.py), same inputs,
same thread count, one frame at a time — and check that both produce the same outputs while at it.

The port is what `bench.py --impl reference` and the `cpu_baseline` leg time on the GPU box (the reference is pure
Python and cannot travel there). This script is the evidence that the port is a fair stand-in for the
reference's own CPU path (same outputs, same speed):

    python tests/golden/compare_port_vs_reference.py [frames, default 15]  >  profiles/r01_port_vs_reference.txt
"""
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
import torch  # noqa: E402

import _ref_loader  # noqa: E402
import bench  # noqa: E402


def main(frames):
    if not _ref_loader.available():
        raise SystemExit("reference sources not present: this tool runs in the build container only")
    cube_pad, e2c_mod, c2e_mod = _ref_loader.load()
    port = bench.CpuChain()                                  # also raises torch's thread count to the host's
    sites = port.sites
    ref_e2c = e2c_mod.Equi2Cube(bench.CUBE, port.frame)
    ref_pads = {p: cube_pad.CubePad(p, use_gpu=False) for p in {s[2] for s in sites}}
    ref_c2e = c2e_mod.Cube2Equi(bench.CUBE // 32)

    def ref_frame():
        faces = ref_e2c.to_cube(port.frame)                                              # equi_to_cube.py:112-129
        x0 = torch.from_numpy(np.stack([faces[i] for i in range(6)])).permute(0, 3, 1, 2).contiguous()
        outs = [ref_pads[sites[0][2]](x0)]                                               # cube_pad.py:23-216
        for (C, H, p), x in zip(sites[1:], port.feats):
            outs.append(ref_pads[p](x))
        equi = ref_c2e.to_equi_nn(port.cam)                                              # cube_to_equi.py:37-66
        return faces, outs, torch.max(equi, 1)[0].squeeze(0)                             # test_temporal.py:82-84

    def port_frame():
        faces = port.e2c.to_cube(port.frame)
        x0 = torch.from_numpy(np.stack([faces[i] for i in range(6)])).permute(0, 3, 1, 2).contiguous()
        outs = [port.pads[sites[0][2]](x0)]
        for (C, H, p), x in zip(sites[1:], port.feats):
            outs.append(port.pads[p](x))
        return faces, outs, port.c2e.to_equi_max(port.cam)

    # same outputs first
    fr, orf, sr = ref_frame()
    fp, opf, sp = port_frame()
    assert all(np.array_equal(fr[i], fp[i]) for i in range(6)), "e2c faces differ"
    assert all(torch.equal(a, b) for a, b in zip(orf, opf)), "CubePad outputs differ"
    assert float((sr.detach() - sp).abs().max()) <= 4e-6, "back-projected maps differ"
    print("outputs: e2c faces bit-identical, 19 CubePad outputs bit-identical, c2e+max max-abs diff %.1e"
          % float((sr.detach() - sp).abs().max()))

    # interleaved, per-frame times, median and best: the host is shared, back-to-back blocks drift by +-20 %
    ref_frame(), port_frame()
    t_ref, t_port = [], []
    for _ in range(frames):
        t0 = time.perf_counter(); ref_frame(); t1 = time.perf_counter(); port_frame(); t2 = time.perf_counter()
        t_ref.append(t1 - t0)
        t_port.append(t2 - t1)
    med = lambda v: sorted(v)[len(v) // 2]                                                # noqa: E731
    print("threads: torch %d, cv2 %s, host cpus %d; %d interleaved frames each" % (torch.get_num_threads(), port.threads(),
                                                                                  os.cpu_count(), frames))
    print("unmodified reference : median %7.1f ms per frame (%.2f frames/s), best %7.1f ms" % (1e3 * med(t_ref), 1 / med(t_ref), 1e3 * min(t_ref)))
    print("oracle/ref_port      : median %7.1f ms per frame (%.2f frames/s), best %7.1f ms" % (1e3 * med(t_port), 1 / med(t_port), 1e3 * min(t_port)))
    print("port / reference speed: %.2fx by medians, %.2fx by best frames"
          % (med(t_ref) / med(t_port), min(t_ref) / min(t_port)))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 15)
