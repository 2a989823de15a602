"""SURVEY.md §8 row a-7 on the GPU: the reference's OWN call sites, run unmodified, with cp360_b200 swapped in.

The reference sources are read from oracle/_ref/ (the byte-identical, sha256-pinned staging copy that
``__graft_entry__.build()`` makes from /root/reference — it travels to the GPU box like the built .so) or from
/root/reference where that exists. Compared on the same CUDA device, same seeded weights and inputs:

 * model/resnet_cubic.py resnet50(): the network built with the reference's own ``model.cube_pad.CubePad`` (≈41
   ATen launches + 8 index uploads per 6-face group, cube_pad.py:73-76,95-216) vs the network built after
   ``model.resnet_cubic.CubePad = cp360_b200.CubePad`` (INTEGRATION.md option A) — logits and the layer4 feature
   map BIT-identical (CubePad is pure data movement; the cuDNN convolutions see identical inputs);
 * model/clstm.py ConvLSTMCell: five steps (config.yaml seq_len) forward bit-identical; backward through the whole
   sequence (train_temporal.py:167-170) equal to fp32 summation-order tolerance;
 * utils/equi_to_cube.py Equi2Cube.to_cube (cv2.remap on the host) vs cp360_b200.Equi2Cube.to_cube: bit-exact faces
   and integer maps; utils/cube_to_equi.py Cube2Equi.to_equi_nn (the reference's CUDA grid_sample path, only
   ``async`` -> ``non_blocking`` patched in memory) vs cp360_b200.Cube2Equi.to_equi_nn: max-abs <= 1e-5, and the
   channel max the call sites take right after (test_temporal.py:82-84) vs to_equi_max.
"""
import warnings

import numpy as np
import pytest
import torch

import cp360_b200
from oracle import ref_loader

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason="reference neither at /root/reference nor staged in "
                                                                    "oracle/_ref (run __graft_entry__.build() first)")]

C2E_TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda", 0)


@pytest.fixture(scope="module")
def ref():
    warnings.filterwarnings("ignore")
    if ref_loader.kind() == "staged":
        assert ref_loader.verify(), "oracle/_ref does not match its sha256 manifest"
    cube_pad, resnet, clstm = ref_loader.load_models()
    assert not getattr(cube_pad.CubePad, "_cp360_cpu_default", False)
    _, e2c_mod, _ = ref_loader.load()
    c2e_cuda = ref_loader.load("cuda")[2]
    return {"cube_pad": cube_pad, "resnet": resnet, "clstm": clstm, "e2c": e2c_mod, "c2e": c2e_cuda}


class _swapped:
    """``with _swapped(module): ...`` — the reference module's global ``CubePad`` name bound to cp360_b200.CubePad,
    exactly what a maintainer's one-line import change does (INTEGRATION.md option A)."""

    def __init__(self, module):
        self.module = module

    def __enter__(self):
        self.saved = self.module.CubePad
        self.module.CubePad = cp360_b200.CubePad

    def __exit__(self, *exc):
        self.module.CubePad = self.saved


def _count(model, cls):
    return sum(isinstance(m, cls) for m in model.modules())


@pytest.mark.parametrize("groups", [1, 2])
def test_reference_resnet50_with_cp360_cubepad_bit_identical(dev, ref, groups):
    resnet, cube_pad = ref["resnet"], ref["cube_pad"]
    torch.manual_seed(10)
    model_ref = resnet.resnet50(pretrained=False).to(dev).eval()
    assert _count(model_ref, cube_pad.CubePad) >= 18 and _count(model_ref, cp360_b200.CubePad) == 0
    with _swapped(resnet):
        model_new = resnet.resnet50(pretrained=False).to(dev).eval()
    assert _count(model_new, cp360_b200.CubePad) == _count(model_ref, cube_pad.CubePad)
    assert _count(model_new, cube_pad.CubePad) == 0
    model_new.load_state_dict(model_ref.state_dict())          # CubePad has no parameters: same keys
    feats = {}
    for tag, m in (("ref", model_ref), ("new", model_new)):
        m.layer4.register_forward_hook(lambda mod, i, o, tag=tag: feats.__setitem__(tag, o.detach().clone()))
    x = torch.randn(6 * groups, 3, 224, 224, device=dev)       # the reference's cube_dim (config.yaml:17)
    before = cp360_b200._lib.launch_count()
    with torch.no_grad():
        want = model_ref(x)
        mid = cp360_b200._lib.launch_count()
        got = model_new(x)
    torch.cuda.synchronize()
    assert mid == before, "the reference network must not touch libcp360"
    assert cp360_b200._lib.launch_count() - mid >= 18, "expected one libcp360 launch per CubePad site"
    assert feats["ref"].shape == (6 * groups, 2048, 7, 7)
    assert torch.equal(feats["new"], feats["ref"]), "layer4 features differ"
    assert torch.equal(got, want), "logits differ"


def test_reference_convlstm_with_cp360_cubepad(dev, ref):
    clstm, cube_pad = ref["clstm"], ref["cube_pad"]
    torch.manual_seed(11)
    feat = hid = 96
    cell_ref = clstm.ConvLSTMCell(feat, hid).to(dev)
    with _swapped(clstm):
        cell_new = clstm.ConvLSTMCell(feat, hid).to(dev)
    assert isinstance(cell_new.pad, cp360_b200.CubePad) and isinstance(cell_ref.pad, cube_pad.CubePad)
    cell_new.load_state_dict(cell_ref.state_dict())
    seq = [torch.randn(6, feat, 7, 7, device=dev) for _ in range(5)]            # seq_len 5, [6,C,7,7] (test_temporal.py:57-79)
    h0 = torch.randn(6, hid, 7, 7, device=dev)

    def run(cell, grad):
        for p in cell.parameters():
            p.grad = None
        xs = [s.clone().requires_grad_(grad) for s in seq]
        state = (h0.clone(), h0.clone())
        outs = []
        for x in xs:
            state = cell(x, state)
            outs.append(state[0])
        if grad:
            torch.stack(outs).square().sum().backward()
        return outs, xs

    with torch.no_grad():
        want, _ = run(cell_ref, False)
        got, _ = run(cell_new, False)
    for a, b in zip(got, want):
        assert torch.equal(a, b), "ConvLSTM hidden state differs"
    # training path: the gradient crosses 15 CubePads (cp360_cubepad_bwd_f32 vs autograd through cat/index_select)
    want, xs_ref = run(cell_ref, True)
    g_ref = [p.grad.clone() for p in cell_ref.parameters()]
    got, xs_new = run(cell_new, True)
    for a, b in zip(got, want):
        assert torch.equal(a.detach(), b.detach())
    for p, g in zip(cell_new.parameters(), g_ref):
        scale = float(g.abs().max()) + 1e-12
        assert float((p.grad - g).abs().max()) <= 2e-5 * scale, "weight gradient differs"
    for a, b in zip(xs_new, xs_ref):
        scale = float(b.grad.abs().max()) + 1e-12
        assert float((a.grad - b.grad).abs().max()) <= 2e-5 * scale, "input gradient differs"


def test_reference_convlstm_reference_width(dev, ref):
    """The shapes the reference actually runs: 1000 -> 1000 channels on 7x7 faces (config.yaml:21-22, clstm.py:57-64):
    CubePad on [6,2000,7,7] and 2x [6,4000,7,7] per step."""
    clstm = ref["clstm"]
    torch.manual_seed(12)
    cell_ref = clstm.ConvLSTMCell(1000, 1000).to(dev).eval()
    with _swapped(clstm):
        cell_new = clstm.ConvLSTMCell(1000, 1000)
    cell_new.load_state_dict(cell_ref.state_dict())
    cell_new = cell_new.to(dev).eval()
    x = torch.randn(6, 1000, 7, 7, device=dev)
    h = torch.randn(6, 1000, 7, 7, device=dev)
    with torch.no_grad():
        want = cell_ref(x, (h, h))
        got = cell_new(x, (h, h))
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])


def test_reference_extractor_call_sites(dev, ref):
    """dataset_feat_extractor.py:132,145 (Equi2Cube / to_cube) and :170,174-175 + test_temporal.py:82-84
    (Cube2Equi / to_equi_nn / channel max), reference objects vs cp360_b200 objects on the same data."""
    rng = np.random.default_rng(5)
    H, W, w = 960, 1920, 224
    img = rng.random((H, W, 3), dtype=np.float32)
    e_ref = ref["e2c"].Equi2Cube(w, img)
    e_new = cp360_b200.Equi2Cube(w, img)
    for f in range(6):
        assert np.array_equal(np.asarray(e_ref.inXs[f]).reshape(-1), e_new.inXs[f]) or \
            float(np.abs(np.asarray(e_ref.inXs[f]).reshape(-1) - e_new.inXs[f]).max()) <= 1e-9
        sx = np.rint(np.asarray(e_ref.inXs[f], dtype=np.float32).reshape(-1) * np.float32(32)).astype(np.int32)
        assert np.array_equal(sx, e_new.sx[f].reshape(-1)), "integer sampling map differs (face %d)" % f
    want = e_ref.to_cube(img)
    got = e_new.to_cube(img)
    for f in range(6):
        assert got[f].dtype == want[f].dtype and np.array_equal(got[f], want[f]), "e2c face %d differs" % f

    fw = 7
    c_ref = ref["c2e"].Cube2Equi(fw)
    c_new = cp360_b200.Cube2Equi(fw)                       # default follows the installed torch (align_corners=False)
    assert np.array_equal(c_ref.face_map, c_new.face_map)
    assert float(np.abs(c_ref.out_coord - c_new.out_coord).max()) <= 1e-12
    torch.manual_seed(6)
    hidden = torch.randn(6, 1000, fw, fw, device=dev)
    with torch.no_grad():
        want = c_ref.to_equi_nn(hidden)                   # the reference's own CUDA path
        got = c_new.to_equi_nn(hidden)
        got_max = c_new.to_equi_max(hidden)
    assert want.shape == got.shape == (1, 1000, 2 * fw, 4 * fw)
    assert float((got - want).abs().max()) <= C2E_TOL
    want_max = torch.max(want, 1)[0]                       # test_temporal.py:83
    assert float((got_max - want_max).abs().max()) <= C2E_TOL
