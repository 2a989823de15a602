"""Oracle: equirectangular -> 6 cube faces (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates utils/equi_to_cube.py:12-129 (+ utils/sph_utils.py:23-38) of the reference and the
arithmetic of the third-party call on the path, ``cv2.remap(..., INTER_LINEAR)`` with float32
maps (opencv 4.13 as installed; 1/32-pixel fixed-point bilinear — the reference pins no
version, README.md:11-16 says "cv2 3.4.2", same algorithm).

  build_maps      float64 sampling coordinates inXs/inYs per face        equi_to_cube.py:12-110
  fixed_point     what cv2 actually consumes: round(f32(coord)*32)        equi_to_cube.py:122-125
  pack_map        the product's packed uint32 map (x0:11|y0:10|fx:5|fy:5) from sx,sy
  to_cube         fixed-point bilinear gather, fp32, no FMA               equi_to_cube.py:112-129
"""
import math

import numpy as np

# yaw, pitch, roll in degrees — equi_to_cube.py:17-22 (Back, Bottom, Front, Left, Right, Top)
VIEWS_DEG = ((180, 0, 0), (0, -90, 0), (0, 0, 0), (-90, 0, 0), (90, 0, 0), (0, 90, 0))
INTER_BITS = 5
INTER_TAB_SIZE = 1 << INTER_BITS


def _rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def build_maps(out_w, in_h, in_w, vfov_deg=90):
    """Returns (inXs, inYs): two lists of 6 float64 arrays [out_w*out_w], 1-based coords."""
    assert in_h * 2 == in_w                                   # equi_to_cube.py:15
    vfov = vfov_deg * np.pi / 180
    views = np.array(VIEWS_DEG) * np.pi / 180
    t = math.tan(vfov / 2)
    top_left = np.array([-t * (out_w / out_w), -t, 1])
    uv = np.array([-2 * top_left[0] / out_w, -2 * top_left[1] / out_w, 0])

    # inverse-trig lookup tables, equi_to_cube.py:49-56
    res_acos, res_atan = 2 * in_w, 2 * in_h
    step_acos, step_atan = np.pi / res_acos, np.pi / res_atan
    lut_acos = np.append(-np.cos(np.arange(0, res_acos) * step_acos), 1.0)
    lut_atan = np.append(
        np.append(np.tan(step_atan / 2 - np.pi / 2),
                  np.tan(np.arange(1, res_atan) * step_atan - np.pi / 2)),
        np.tan(-step_atan / 2 + np.pi / 2))
    idx_acos = np.arange(0, res_acos + 1)
    idx_atan = np.arange(0, res_atan + 1)

    X, Y = np.meshgrid(range(out_w), range(out_w))
    X, Y = X.flatten(), Y.flatten()
    n = X.shape[0]
    pts = np.stack([top_left[0] + uv[0] * X, top_left[1] + uv[1] * Y,
                    top_left[2] + uv[2] * np.ones(n)], axis=0)

    in_xs, in_ys = [], []
    for yaw, pitch, roll in views:
        tf = np.dot(np.dot(_rot_y(yaw), _rot_x(pitch)), _rot_z(roll))
        mv = np.dot(tf, pts)
        xp, yp, zp = mv[0], mv[1], mv[2]
        nxz = np.sqrt(xp ** 2 + zp ** 2)
        phi, theta = np.zeros(n), np.zeros(n)
        pole = nxz < 10e-10                                    # equi_to_cube.py:86
        phi[pole & (yp > 0)] = np.pi / 2
        phi[pole & (yp <= 0)] = -np.pi / 2
        ok = ~pole
        # scipy interp1d(kind='linear') on a 1-D float64 abscissa is numpy.interp
        # (scipy/interpolate/_interpolate.py: _call_linear_np); bounds_error => raise.
        qa = yp[ok] / nxz[ok]
        qc = -zp[ok] / nxz[ok]
        if qa.min() < lut_atan[0] or qa.max() > lut_atan[-1] or qc.min() < lut_acos[0] \
                or qc.max() > lut_acos[-1]:
            raise ValueError("A value in x_new is outside the interpolation range.")
        phi[ok] = np.interp(qa, lut_atan, idx_atan) * step_atan - (np.pi / 2)
        theta[ok] = np.interp(qc, lut_acos, idx_acos) * step_acos
        neg = ok & (xp < 0)
        theta[neg] = -theta[neg]
        in_x = (theta / np.pi) * (in_w / 2) + (in_w / 2) + 1     # 1-based, :100-101
        in_y = (phi / (np.pi / 2)) * (in_h / 2) + (in_h / 2) + 1
        in_x[in_x < 1] = 1
        in_x[in_x >= in_w - 1] = in_w - 1
        in_y[in_y < 1] = 1
        in_y[in_y >= in_h - 1] = in_h - 1
        in_xs.append(in_x)
        in_ys.append(in_y)
    return in_xs, in_ys


def fixed_point(coord_f64):
    """cv2.remap's map conversion: cvRound(float32(v) * 32), round-half-to-even -> int32."""
    v32 = np.asarray(coord_f64).astype(np.float32)
    return np.rint(v32 * np.float32(INTER_TAB_SIZE)).astype(np.int32)


def fixed_maps(out_w, in_h, in_w, vfov_deg=90):
    """int32 (sx, sy), each [6, out_w, out_w]."""
    xs, ys = build_maps(out_w, in_h, in_w, vfov_deg)
    sx = np.stack([fixed_point(a).reshape(out_w, out_w) for a in xs])
    sy = np.stack([fixed_point(a).reshape(out_w, out_w) for a in ys])
    return sx, sy


def pack_map(sx, sy):
    """uint32 x0<<20 | y0<<10 | fx<<5 | fy (x0 < 2048, y0 < 1024)."""
    x0, fx = sx >> INTER_BITS, sx & (INTER_TAB_SIZE - 1)
    y0, fy = sy >> INTER_BITS, sy & (INTER_TAB_SIZE - 1)
    assert x0.max() < 2048 and y0.max() < 1024 and x0.min() >= 0 and y0.min() >= 0
    return ((x0.astype(np.uint32) << 20) | (y0.astype(np.uint32) << 10)
            | (fx.astype(np.uint32) << 5) | fy.astype(np.uint32))


def bilinear_fixed(img, sx, sy):
    """cv2.remap(INTER_LINEAR, BORDER_CONSTANT=0) arithmetic for an fp32 H x W x C image.

    img [H,W,C] float32; sx, sy int32 [...]; returns [..., C] float32.
    Weights are the products of {1-f, f} with f = k/32 (exact in fp32); the four taps are
    accumulated left to right with separate multiply and add (no FMA) in fp32.
    """
    img = np.asarray(img, dtype=np.float32)
    H, W = img.shape[:2]
    x0, y0 = sx >> INTER_BITS, sy >> INTER_BITS
    fx = (sx & (INTER_TAB_SIZE - 1)).astype(np.float32) / np.float32(INTER_TAB_SIZE)
    fy = (sy & (INTER_TAB_SIZE - 1)).astype(np.float32) / np.float32(INTER_TAB_SIZE)
    one = np.float32(1)
    w00 = ((one - fy) * (one - fx))[..., None]
    w01 = ((one - fy) * fx)[..., None]
    w10 = (fy * (one - fx))[..., None]
    w11 = (fy * fx)[..., None]

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok[..., None], v, np.float32(0))

    acc = tap(y0, x0) * w00
    acc = acc + tap(y0, x0 + 1) * w01
    acc = acc + tap(y0 + 1, x0) * w10
    acc = acc + tap(y0 + 1, x0 + 1) * w11
    return acc.astype(np.float32)


def to_cube(img, sx, sy):
    """img [H,W,C] fp32 -> faces [6, w, w, C] fp32 (face-major stack of the reference's dict)."""
    return bilinear_fixed(img, sx, sy)
