"""TEST INFRASTRUCTURE — numpy restatement of the reference's CAM contraction and heat-map
post-processing (static_model/class_activation_model.py:46-52, 76-90;
static_model/dataset_feat_extractor.py:174-176; utils/utils.py:15-17). Checker only: imported by
tests/ (never by the product path). Pinned against the reference's own CAM() on a stub model by
tests/golden/make_golden_cam.py -> tests/golden/golden_cam.npz."""
import numpy as np


def cam_weight(fc_weight):
    w = np.squeeze(np.array(fc_weight, dtype=np.float32, copy=True))       # :46-50
    if np.min(w) < 0:
        w -= np.min(w)                                                      # :51-52
    return w


def cam_scores(features, fc_weight):
    """features [bz, nc, h, w] -> [bz, classes, h, w]: one weight.dot(features[idx]) per face (:76-90)."""
    w = cam_weight(fc_weight)
    bz, nc, h, ww = features.shape
    f = features.reshape(bz, nc, h * ww)
    out = np.stack([w.dot(f[i]) for i in range(bz)], 0)
    return out.reshape(bz, w.shape[0], h, ww)


def heatmap(equi, normalize=False):
    """equi [C, H, W] -> np.max over channels, squared (dataset_feat_extractor.py:175-176)
    [-> min-max normalised as overlay() does, utils/utils.py:15-17]."""
    s = np.max(equi, 0)
    s = s[:, :] ** 2
    if normalize:
        s = s - np.min(s)
        s = s / np.max(s)
    return s
