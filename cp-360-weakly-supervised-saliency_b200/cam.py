"""Class-activation scores feeding the back-projection — the GPU side of CAM()
(static_model/class_activation_model.py:46-52, 76-90) and of the heat-map post-processing
(static_model/dataset_feat_extractor.py:174-176, utils/utils.py:15-17).

The reference copies the layer4 features to the host through a forward hook, min-shifts the fc
weight, and runs one numpy `weight.dot(features[idx])` per face. Here the contraction stays on the
device as ONE batched library GEMM (torch.matmul -> cuBLAS, a plain GEMM and not the product of
this repo) that writes the `[6B, classes, h, w]` layout the fused back-projection + channel-max
kernel (cp360_c2e_max_fwd) consumes directly, so the class scores never leave HBM.
"""
import torch

from .cube_to_equi import Cube2Equi


def cam_weight(fc_weight):
    """`weight_softmax` of the reference: squeeze, and shift by the minimum if any entry is negative
    (class_activation_model.py:46-52)."""
    w = fc_weight.detach().float()
    w = w.reshape(w.shape[0], -1)
    m = w.min()
    return w - m if bool(m < 0) else w


def cam_scores(features, fc_weight, shifted=False):
    """features [6B, nc, h, w] cuda, fc_weight [classes, nc] -> cube scores [6B, classes, h, w]
    (class_activation_model.py:76-90), fp32, IEEE (no TF32)."""
    if not features.is_cuda:
        raise RuntimeError("cam_scores: features are on %s; CUDA only, no CPU fallback" % features.device)
    n, nc, h, w = features.shape
    wt = fc_weight if shifted else cam_weight(fc_weight)
    wt = wt.to(features.device)
    if wt.shape[1] != nc:
        raise ValueError("fc weight has %d input channels, features have %d" % (wt.shape[1], nc))
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = torch.matmul(wt, features.float().reshape(n, nc, h * w))          # [classes,nc] x [6B,nc,hw]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out.reshape(n, wt.shape[0], h, w)


class SaliencyHead:
    """features -> equirect saliency, all on the device:
    cam_scores -> Cube2Equi.to_equi_max (K3m) -> **2 (dataset_feat_extractor.py:176) [-> min-max normalise
    (utils/utils.py:15-17)]."""

    def __init__(self, fc_weight, feat_w, align_corners=False):
        self.weight = cam_weight(fc_weight)
        self.c2e = Cube2Equi(int(feat_w), align_corners=align_corners)

    def scores(self, features):
        return cam_scores(features, self.weight, shifted=True)

    def __call__(self, features, normalize=False):
        sal = self.c2e.to_equi_max(self.scores(features))                        # [B, 2w, 4w]
        sal = sal * sal
        if normalize:
            flat = sal.reshape(sal.shape[0], -1)
            lo = flat.min(dim=1, keepdim=True)[0]
            flat = flat - lo
            sal = (flat / flat.max(dim=1, keepdim=True)[0]).reshape(sal.shape)
        return sal
