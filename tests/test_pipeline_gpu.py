"""GPU tests of the batched chain (cp360_b200.SphericalPipeline) — the object bench.py times and the entry point of the
end-to-end number (process_host). Every stage's buffer is compared with the single-operator API applied to the same
input (those operators are pinned to the oracle / the reference's golden vectors in tests/test_gpu_parity.py), bit for
bit: the chain adds batching, buffer reuse, graph capture and staging, never arithmetic."""
import numpy as np
import pytest
import torch

import cp360_b200

pytestmark = pytest.mark.gpu

EQUI_H, EQUI_W, CUBE, CAM_C, FEAT_C = 96, 192, 64, 24, 32        # cube 64 -> 2x2 feature faces, sites down to 2x2


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def make(dev, B, **kw):
    pipe = cp360_b200.SphericalPipeline(EQUI_H, EQUI_W, CUBE, CAM_C, FEAT_C, device=dev, seed=7, **kw)
    pipe.allocate(B)
    return pipe, pipe.synthetic_frames(B)


def check_pads(pipe, first):
    for i, (C, H, p) in enumerate(pipe.sites):
        if i < first:
            continue
        assert torch.equal(pipe.site_out[i], cp360_b200.CubePad(p)(pipe.site_in[i])), "site %d %s" % (i, (C, H, p))


@pytest.mark.parametrize("B", [1, 3])
def test_step_equals_the_single_operators(dev, B):
    pipe, frames = make(dev, B)
    sal = pipe.step(frames)
    assert tuple(sal.shape) == (B, 2 * pipe.feat_w, 4 * pipe.feat_w)
    assert torch.equal(pipe.faces, pipe.e2c.to_cube_tensor(frames))          # K1, [6B,3,w,w]
    check_pads(pipe, 0)                                                        # K2 at every site (site 0 reads the faces)
    assert torch.equal(sal, pipe.c2e.to_equi_max(pipe.cam))                    # K3m
    assert pipe.launches_per_step() == 1 + len(pipe.sites) + 1


def test_first_site_fused_into_e2c_gives_the_same_padded_tensor(dev):
    """bench.py's default chain: e2c and the CubePad(3) in front of conv1 as ONE kernel (cp360_e2c_cubepad_fwd)."""
    ref, frames = make(dev, 2)
    ref.step(frames)
    pipe, _ = make(dev, 2, fuse_first_site=True)
    pipe.faces.fill_(float("nan"))                                             # never written on this path
    sal = pipe.step(frames)
    assert torch.equal(pipe.site_out[0], ref.site_out[0])
    assert torch.isnan(pipe.faces).all()
    check_pads(pipe, 1)
    assert torch.equal(sal, ref.sal)
    assert pipe.launches_per_step() == ref.launches_per_step() - 1
    w, p0 = CUBE, pipe.sites[0][2]
    assert ref.bytes_per_frame() - pipe.bytes_per_frame() == 2 * 6 * 3 * w * w * 4    # the faces' write and read
    # uint8 frames (what the end-to-end path uploads): float32(u8) / 255 on the fly, in both forms
    u8 = (frames * 255).to(torch.uint8)
    ref.step(u8)
    pipe.step(u8)
    assert torch.equal(pipe.site_out[0], ref.site_out[0])
    assert torch.equal(ref.faces, ref.e2c.to_cube_tensor(u8))


def test_step_fused_equals_pad_of_the_producer_ops(dev):
    pipe, frames = make(dev, 2)
    sal = pipe.step_fused(frames)
    faces = pipe.e2c.to_cube_tensor(frames)
    assert torch.equal(pipe.site_out[0], cp360_b200.CubePad(pipe.sites[0][2])(faces))
    last = len(pipe.sites) - 1
    for i, (C, H, p) in enumerate(pipe.sites):
        if i == 0:
            continue
        if i < last:                                                           # eval-mode BN affine + ReLU, then the pad
            x = pipe.site_in[i] * pipe.bn_scale[i].view(1, -1, 1, 1) + pipe.bn_shift[i].view(1, -1, 1, 1)
            want = cp360_b200.CubePad(p)(torch.relu(x))
        else:                                                                  # ConvLSTM site: pad of the cat of two sources
            want = cp360_b200.CubePad(p)(torch.cat(pipe.cat_src, 1))
        assert torch.equal(pipe.site_out[i], want), "site %d" % i
    assert torch.equal(sal, pipe.c2e.to_equi_max(pipe.cam))


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("first", [False, True])
def test_graph_replay_equals_eager(dev, fused, first):
    pipe, frames = make(dev, 2, fuse_first_site=first)
    (pipe.step_fused if fused else pipe.step)(frames)
    want = [t.clone() for t in pipe.site_out] + [pipe.sal.clone()]
    graph = pipe.capture(frames, fused=fused)
    for t in pipe.site_out:
        t.zero_()
    pipe.sal.zero_()
    graph.replay()
    torch.cuda.synchronize()
    for a, b in zip(pipe.site_out + [pipe.sal], want):
        assert torch.equal(a, b)


@pytest.mark.parametrize("dtype", [torch.uint8, torch.float32])
@pytest.mark.parametrize("depth,streams", [(2, 1), (3, 2)])
def test_process_host_streams_every_batch_through_the_chain(dev, dtype, depth, streams):
    """The end-to-end entry point: page-locked host batches in, host maps out, uploads running ahead of the chain in a
    staging ring. The maps come from the resident stand-in features (every batch must deliver them); that the uploads
    really pass through e2c is checked on the padded conv1 input the LAST batch leaves behind, twice (ring reuse)."""
    B, n = 2, 5
    pipe, _ = make(dev, B, fuse_first_site=True)
    rng = np.random.default_rng(3)
    if dtype == torch.uint8:
        host = [torch.from_numpy(rng.integers(0, 256, (B, EQUI_H, EQUI_W, 3), dtype=np.uint8)).pin_memory() for _ in range(n)]
    else:
        host = [torch.from_numpy(rng.random((B, EQUI_H, EQUI_W, 3), dtype=np.float32)).pin_memory() for _ in range(n)]
    out = torch.full((n, B, 2 * pipe.feat_w, 4 * pipe.feat_w), float("nan")).pin_memory()
    pipe.process_host(host, out, depth=depth, copy_streams=streams)
    torch.cuda.synchronize()
    want_sal = pipe.c2e.to_equi_max(pipe.cam).cpu()
    for i in range(n):
        assert torch.equal(out[i], want_sal)                                   # the chain's maps come from the resident features
    last = host[-1].to(dev)
    want0 = cp360_b200.CubePad(pipe.sites[0][2])(pipe.e2c.to_cube_tensor(last))
    assert torch.equal(pipe.site_out[0], want0)                                # the last upload really went through e2c
    # staging ring reuse: a second call with other data leaves the new last batch's tensor
    host2 = [h.clone().pin_memory() for h in reversed(host)]
    pipe.process_host(host2, out, depth=depth, copy_streams=streams)
    torch.cuda.synchronize()
    want1 = cp360_b200.CubePad(pipe.sites[0][2])(pipe.e2c.to_cube_tensor(host2[-1].to(dev)))
    assert torch.equal(pipe.site_out[0], want1)
