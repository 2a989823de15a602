"""CPU, world_size 2 over gloo: the frame partition and the single final exchange of the path
(SphericalPipeline's N>1 logic). No kernels are launched."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, ROOT)
    import cp360_b200
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = cp360_b200.shard_range(n_total, rank, world)
    # "saliency map" of frame i is filled with i so the order of the gathered stack is checkable
    local = torch.arange(a, b, dtype=torch.float32).view(-1, 1, 1).expand(-1, 4, 8).contiguous()
    full = cp360_b200.gather_maps(local, n_total)
    ok = full.shape == (n_total, 4, 8) and torch.equal(full[:, 0, 0], torch.arange(n_total, dtype=torch.float32))
    q.put((rank, (a, b), bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [2000, 7, 2])
def test_shard_and_gather_world2(n_total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, _, ok in res)
    ranges = [r for _, r, _ in res]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_total and ranges[0][1] == ranges[1][0]


def test_shard_range_properties():
    import cp360_b200
    for n in (0, 1, 5, 2000, 2001):
        for world in (1, 2, 4, 8):
            parts = [cp360_b200.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_prefer_gpu_numa_node_is_advisory(monkeypatch):
    """The NUMA preference of the multi-GPU host path never raises: unknown node -> not applied;
    a known node -> MPOL_PREFERRED (allocations keep working either way)."""
    import importlib
    import cp360_b200
    pl = importlib.import_module(cp360_b200.SphericalPipeline.__module__)
    monkeypatch.setattr(pl, "gpu_numa_node", lambda d: -1)
    info = pl.prefer_gpu_numa_node(0)
    assert info["applied"] is False and "why" in info
    monkeypatch.setattr(pl, "gpu_numa_node", lambda d: 0)
    info = pl.prefer_gpu_numa_node(0)
    assert info["node"] == 0 and (info["applied"] or "why" in info)
    assert torch.empty(1 << 16).fill_(1).sum().item() == float(1 << 16)
    monkeypatch.setattr(pl, "gpu_numa_node", lambda d: 1000)          # no such node on this host
    assert pl.prefer_gpu_numa_node(0)["applied"] is False
