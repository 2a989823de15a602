"""CPU port of the reference hot path built from the SAME LIBRARY CALLS the reference makes
(TEST INFRASTRUCTURE / CPU BASELINE ONLY — see oracle/__init__.py).

The reference is pure Python and cannot travel to the GPU box (/root/reference does not exist
there), so bench.py's ``cpu_baseline`` and ``--impl reference`` legs time this port
(cpu_baseline.kind = "port"). It keeps the reference's work structure so the timing is
representative of the reference's own CPU path:

  CubePadPort     per 6-face group: slice / flip / transpose plates, repeat corners, torch.cat
                  vertically then horizontally, final torch.cat over groups
                  (model/cube_pad.py:28-42, :95-216)
  Equi2CubePort   6 faces x C channels of cv2.remap(INTER_LINEAR) on strided channel planes with
                  per-call float64->float32 map casts (utils/equi_to_cube.py:112-129)
  Cube2EquiPort   6 full-grid F.grid_sample passes + boolean-mask assignment
                  (utils/cube_to_equi.py:37-66), then torch.max over channels
                  (temporal_model/test_temporal.py:82-84)

tests/test_oracle_golden.py checks each port against the fixtures generated from the reference
(bit-exact for CubePad and e2c faces, <= 4e-6 for c2e).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import c2e as _c2e
from . import e2c as _e2c
from .cubepad import B, D, F as FF, L, R, T, get_pad_size

# (face', transposed, flip_rows, flip_cols, which slice) per face for each plate; the slice is
# described by name and resolved against the pad size at call time.
#   'top'   rows [0,p)        'bot'   rows [H-p,H)      'lft' cols [0,p)     'rgt' cols [W-p,W)
_TOP = {B: (T, 'top', False, False, True), D: (FF, 'bot', False, False, False),
        FF: (T, 'bot', False, False, False), L: (T, 'lft', True, False, False),
        R: (T, 'rgt', True, False, True), T: (B, 'top', False, False, True)}
_DOWN = {B: (D, 'bot', False, False, True), D: (B, 'bot', False, False, True),
         FF: (D, 'top', False, False, False), L: (D, 'lft', True, False, True),
         R: (D, 'rgt', True, False, False), T: (FF, 'top', False, False, False)}
_LEFT = {B: (R, 'rgt', False, False, False), D: (L, 'bot', True, True, False),
         FF: (L, 'rgt', False, False, False), L: (B, 'rgt', False, False, False),
         R: (FF, 'rgt', False, False, False), T: (L, 'top', True, False, False)}
_RIGHT = {B: (L, 'lft', False, False, False), D: (R, 'bot', True, False, False),
          FF: (R, 'lft', False, False, False), L: (FF, 'lft', False, False, False),
          R: (B, 'lft', False, False, False), T: (R, 'top', True, True, False)}


def _plate(x6, spec, p):
    """One padding plate [C, rows, cols] of a face from its neighbour (slice/transpose/flip)."""
    face, which, transposed, flip_rows, flip_cols = spec
    src = x6[face]
    H, W = src.shape[-2:]
    if which == 'top':
        s = src[:, :p, :]
    elif which == 'bot':
        s = src[:, H - p:, :]
    elif which == 'lft':
        s = src[:, :, :p]
    else:
        s = src[:, :, W - p:]
    if transposed:
        s = s.transpose(1, 2)
    if flip_rows:
        s = torch.flip(s, [1])
    if flip_cols:
        s = torch.flip(s, [2])
    return s


def _corner(td, lr, top, left):
    """make_cubepad_edge (cube_pad.py:83-90): repeat the l/r plate row or the t/d plate column."""
    td_pad, lr_pad = td.shape[1], lr.shape[2]
    if td_pad > lr_pad:
        row = lr[:, :1, :] if top else lr[:, -1:, :]
        return row.repeat(1, td_pad, 1)
    col = td[:, :, :1] if left else td[:, :, -1:]
    return col.repeat(1, 1, lr_pad)


class CubePadPort:
    def __init__(self, lrtd_pad):
        self.pads = get_pad_size(lrtd_pad)

    def _group(self, x6):
        pl, pr, pt, pd = self.pads
        out = []
        for f in range(6):
            t, d = _plate(x6, _TOP[f], pt), _plate(x6, _DOWN[f], pd)
            l, r = _plate(x6, _LEFT[f], pl), _plate(x6, _RIGHT[f], pr)
            mid = torch.cat([l, x6[f], r], dim=2)
            rows = []
            if pt:
                rows.append(torch.cat([_corner(t, l, True, True), t, _corner(t, r, True, False)], dim=2))
            rows.append(mid)
            if pd:
                rows.append(torch.cat([_corner(d, l, False, True), d, _corner(d, r, False, False)], dim=2))
            out.append(torch.cat(rows, dim=1))
        return torch.stack(out, 0)

    def __call__(self, x):
        if x.shape[0] % 6:
            raise ValueError("CubePad size mismatch!")
        return torch.cat([self._group(x[6 * i:6 * i + 6]) for i in range(x.shape[0] // 6)], dim=0)


class Equi2CubePort:
    def __init__(self, out_w, in_h, in_w, vfov=90):
        self.w = out_w
        self.inXs, self.inYs = _e2c.build_maps(out_w, in_h, in_w, vfov)

    def to_cube(self, img):
        import cv2
        w = self.w
        out = {}
        for f in range(6):
            face = np.zeros((w, w, img.shape[2]), img.dtype)
            for c in range(img.shape[2]):
                face[:, :, c] = cv2.remap(img[:, :, c], self.inXs[f].astype('float32').reshape(w, w),
                                          self.inYs[f].astype('float32').reshape(w, w), cv2.INTER_LINEAR)
            out[f] = face
        return out


class Cube2EquiPort:
    def __init__(self, w):
        self.w = w
        self.face_map, self.out_coord = _c2e.build_maps(w)

    def to_equi_nn(self, cube):
        grid = torch.from_numpy(self.out_coord.astype(np.float32))
        fmap = torch.from_numpy(self.face_map.astype(np.int64))
        M = torch.max(grid)
        gn = ((grid - M / 2) / (M / 2)).unsqueeze(0)
        C = cube.shape[1]
        out = torch.zeros(1, C, 2 * self.w, 4 * self.w)
        for f in range(6):
            mask = (fmap == f).unsqueeze(0).unsqueeze(0).expand(1, C, -1, -1)
            s = F.grid_sample(cube[f:f + 1], gn, mode='bilinear', padding_mode='zeros', align_corners=False)
            out[mask] = s[mask]
        return out

    def to_equi_max(self, cube):
        return torch.max(self.to_equi_nn(cube), 1)[0].squeeze(0)
