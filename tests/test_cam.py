"""CAM contraction feeding the back-projection.

CPU: the numpy oracle (oracle/cam.py) against fixtures produced by the reference's own CAM()
(tests/golden/make_golden_cam.py). GPU: cp360_b200.cam_scores / SaliencyHead against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import cam as ocam
from oracle import c2e as oc2e

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_cam.npz"))
TAGS = ("a", "b", "c")


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_cam_matches_reference(tag):
    fc, feat = GOLD["cam_%s_fc" % tag], GOLD["cam_%s_feat" % tag]
    np.testing.assert_array_equal(ocam.cam_weight(fc), GOLD["cam_%s_wshift" % tag])
    # same numpy dot on the same operands: identical bits
    np.testing.assert_array_equal(ocam.cam_scores(feat, fc), GOLD["cam_%s_score" % tag])


def test_oracle_heatmap():
    equi = np.random.default_rng(0).standard_normal((5, 4, 8)).astype(np.float32)
    h = ocam.heatmap(equi)
    np.testing.assert_array_equal(h, np.max(equi, 0) ** 2)
    n = ocam.heatmap(equi, normalize=True)
    assert n.min() == 0.0 and n.max() == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("tag", TAGS)
def test_cam_scores_gpu_vs_golden(tag):
    import cp360_b200
    dev = torch.device("cuda", 0)
    fc, feat = GOLD["cam_%s_fc" % tag], GOLD["cam_%s_feat" % tag]
    got = cp360_b200.cam_scores(torch.from_numpy(feat).to(dev), torch.from_numpy(fc).to(dev))
    want = GOLD["cam_%s_score" % tag]
    assert got.shape == want.shape and got.is_cuda
    # fp32 GEMM with a different summation order than numpy's dot: relative 1e-5 of the score scale
    tol = 1e-5 * max(1.0, float(np.abs(want).max()))
    assert float(np.abs(got.cpu().numpy() - want).max()) <= tol
    np.testing.assert_array_equal(cp360_b200.cam_weight(torch.from_numpy(fc)).numpy(), GOLD["cam_%s_wshift" % tag])


@pytest.mark.gpu
def test_saliency_head_full_shapes():
    """ResNet-50 shapes: [96,2048,8,8] features, fc [1000,2048] -> [16,16,32] saliency; against the
    oracle chain cam -> to_equi -> max -> **2 on one frame, and the fused path's internal consistency."""
    import cp360_b200
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(11)
    feat = torch.randn(96, 2048, 8, 8, device=dev, generator=g).abs_()          # post-ReLU features
    fc = torch.randn(1000, 2048, device=dev, generator=g) * 0.02
    head = cp360_b200.SaliencyHead(fc, 8)
    sal = head(feat)
    assert tuple(sal.shape) == (16, 16, 32)
    scores = head.scores(feat)
    assert tuple(scores.shape) == (96, 1000, 8, 8)
    assert torch.equal(sal, head.c2e.to_equi_max(scores) ** 2)
    # frame 3 against the numpy oracle
    face, coord = oc2e.build_maps(8)
    s_np = ocam.cam_scores(feat[18:24].cpu().numpy(), fc.cpu().numpy())
    want = ocam.heatmap(oc2e.to_equi(s_np, face, coord)[0])
    err = float(np.abs(sal[3].cpu().numpy() - want).max())
    assert err <= 1e-4 * float(np.abs(want).max()), err
    n = head(feat, normalize=True)
    assert float(n.min()) == 0.0 and float(n.max()) == 1.0
