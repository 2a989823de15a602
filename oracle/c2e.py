"""Oracle: 6 cube faces -> equirectangular (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates utils/cube_to_equi.py:12-66 and utils/sph_utils.py:53-153 of the reference, and the
arithmetic of the third-party call on the path, ``torch.nn.functional.grid_sample`` (bilinear,
padding_mode='zeros'; installed torch 2.11 default ``align_corners=False`` is what the
unmodified reference computes today — torch/include/ATen/native/GridSampler.h:27-36).

  build_maps   face_map [2w,4w] (0..5) and out_coord [2w,4w,2] float64   cube_to_equi.py:12-35
  sample_plan  fp32 per-pixel taps/weights as torch's CUDA kernel forms them
  to_equi      single pass equivalent of the 6x grid_sample + masked scatter  cube_to_equi.py:37-66
"""
import numpy as np

FACE_B, FACE_D, FACE_F, FACE_L, FACE_R, FACE_T = range(6)


def build_maps(w):
    out_w, out_h = 4 * w, 2 * w
    XX, YY = np.meshgrid(range(out_w), range(out_h))
    # xy2angle, sph_utils.py:53-60
    theta = (2 * (XX + 0.5) / float(out_w) - 1) * np.pi
    phi = (1 - 2 * (YY + 0.5) / float(out_h)) * np.pi / 2
    # pruned_inf, sph_utils.py:70-77
    err = 10e-9
    for a in (theta, phi):
        a[a == 0.0] = err
        a[a == np.pi] = np.pi - err
        a[a == -np.pi] = -np.pi + err
        a[a == np.pi / 2] = np.pi / 2 - err
        a[a == -np.pi / 2] = -np.pi / 2 + err
    # to_3dsphere, sph_utils.py:63-67 (R = 1)
    x = 1 * np.cos(phi) * np.cos(theta)
    y = 1 * np.sin(phi)
    z = 1 * np.cos(phi) * np.sin(theta)

    # get_face, sph_utils.py:88-111. NB np.maximum(a, b, c) treats c as out=, so the
    # reference's "max of three" is max(|x|,|y|) only — replicated, not fixed.
    eps = 10e-9
    m = np.maximum(np.abs(x), np.abs(y))
    xf, yf, zf = m - np.abs(x) < eps, m - np.abs(y) < eps, m - np.abs(z) < eps
    face = np.zeros((out_h, out_w))
    face[(x >= 0) & xf] = FACE_F
    face[(x <= 0) & xf] = FACE_B
    face[(y >= 0) & yf] = FACE_T
    face[(y <= 0) & yf] = FACE_D
    face[(z >= 0) & zf] = FACE_R
    face[(z <= 0) & zf] = FACE_L

    # face_to_cube_coord, sph_utils.py:114-146
    d = np.zeros((out_h, out_w, 3))
    for fid, (a, b, c) in {FACE_F: (z, y, x), FACE_B: (-z, y, x), FACE_T: (z, -x, y),
                           FACE_D: (z, x, y), FACE_R: (-x, y, z), FACE_L: (x, y, z)}.items():
        k = face == fid
        d[k, 0], d[k, 1], d[k, 2] = a[k], b[k], c[k]
    x_on = (d[:, :, 0] / np.abs(d[:, :, 2]) + 1) / 2
    y_on = (-d[:, :, 1] / np.abs(d[:, :, 2]) + 1) / 2
    coord = np.transpose(np.array([x_on, y_on]), (1, 2, 0))
    # norm_to_cube, sph_utils.py:149-153
    coord = coord * (w - 1)
    coord[coord < 0.] = 0.
    coord[coord > (w - 1)] = (w - 1)
    return face, coord


def sample_plan(out_coord, w, align_corners=False):
    """fp32 plan: (x0, y0 int32 [2w,4w], weights float32 [2w,4w,4] order nw,ne,sw,se, M)."""
    g = np.asarray(out_coord).astype(np.float32)
    M = g.max()                                              # cube_to_equi.py:58 (both coords)
    half = M / np.float32(2)
    gn = (g - half) / half
    one, two = np.float32(1), np.float32(2)
    if align_corners:
        pix = ((gn + one) / two) * np.float32(w - 1)
    else:
        pix = ((gn + one) * np.float32(w) - one) / two
    ix, iy = pix[..., 0], pix[..., 1]
    x_w, y_n = np.floor(ix), np.floor(iy)
    x_e, y_s = x_w + one, y_n + one
    wts = np.stack([(x_e - ix) * (y_s - iy), (ix - x_w) * (y_s - iy),
                    (x_e - ix) * (iy - y_n), (ix - x_w) * (iy - y_n)], axis=-1)
    return x_w.astype(np.int32), y_n.astype(np.int32), wts.astype(np.float32), float(M)


def to_equi(cube, face_map, out_coord, align_corners=False):
    """cube [6,C,w,w] fp32 -> [1,C,2w,4w] fp32."""
    cube = np.asarray(cube, dtype=np.float32)
    _, C, w, _ = cube.shape
    x0, y0, wts, _ = sample_plan(out_coord, w, align_corners)
    f = np.asarray(face_map).astype(np.int64)
    out = np.zeros((C,) + f.shape, dtype=np.float32)
    for k, (dy, dx) in enumerate(((0, 0), (0, 1), (1, 0), (1, 1))):
        yy, xx = y0 + dy, x0 + dx
        ok = (yy >= 0) & (yy < w) & (xx >= 0) & (xx < w)
        v = cube[f, :, np.clip(yy, 0, w - 1), np.clip(xx, 0, w - 1)]      # [2w,4w,C]
        v = np.where(ok[..., None], v * wts[..., k:k + 1], np.float32(0))
        out = out + np.transpose(v, (2, 0, 1))
    return out[None]


def to_equi_max(cube, face_map, out_coord, align_corners=False):
    """Channel max of to_equi -> [2w,4w] (test_temporal.py:82-84, train_temporal.py:105-106)."""
    return to_equi(cube, face_map, out_coord, align_corners)[0].max(axis=0)


# ------------------------------------------------------------------------------------------------
# to_equi_cv2 — utils/cube_to_equi.py:68-91: cv2.remap(face[..., 4d:4d+4], out_coord_x, out_coord_y,
# INTER_CUBIC) per face and channel quad, masked by face_map. Coordinates are used as they are
# (face pixels, NOT normalised like to_equi_nn). The third-party arithmetic restated here is
# OpenCV's remapBicubic for float sources (imgproc/src/imgwarp.cpp; installed cv2 4.13, the
# algorithm is unchanged since 2.x):
#   * float maps -> 1/32-pixel fixed point: s = cvRound(float32(coord) * 32); tap origin (s >> 5) - 1,
#     fraction s & 31 picks a row of the bicubic coefficient table (A = -0.75, fp32);
#   * the 4x4 weight is fl(wy[i] * wx[j]) (table of products, fp32);
#   * origin with all 16 taps inside the face: sum = r0; sum += r1; sum += r2; sum += r3 with
#     r_i = ((S0*w0 + S1*w1) + S2*w2) + S3*w3 (fp32, no FMA);
#   * otherwise (BORDER_CONSTANT, value 0): sum = 0, then sum += S*w tap by tap in row-major order,
#     skipping taps outside the face.
# ------------------------------------------------------------------------------------------------
def cubic_table():
    """[32,4] fp32 coefficients of OpenCV's interpolateCubic at x = k/32."""
    f = np.float32
    x = np.arange(32, dtype=np.float32) * f(1.0 / 32)
    A = f(-0.75)
    x1 = x + f(1)
    c0 = ((A * x1 - f(5) * A) * x1 + f(8) * A) * x1 - f(4) * A
    c1 = ((A + f(2)) * x - (A + f(3))) * x * x + f(1)
    xm = f(1) - x
    c2 = ((A + f(2)) * xm - (A + f(3))) * xm * xm + f(1)
    c3 = f(1) - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], axis=1).astype(np.float32)


def cubic_plan(out_coord):
    """(x0, y0, fx, fy) int32 [2w,4w]: tap origin (top-left of the 4x4 window) and 1/32 fractions."""
    g = np.asarray(out_coord).astype(np.float32)
    s = np.rint(g.astype(np.float64) * 32).astype(np.int64)      # cvRound: round half to even
    sx, sy = s[..., 0], s[..., 1]
    return ((sx >> 5) - 1).astype(np.int32), ((sy >> 5) - 1).astype(np.int32), \
        (sx & 31).astype(np.int32), (sy & 31).astype(np.int32)


def to_equi_cv2(cube, face_map, out_coord):
    """cube [6,C,w,w] fp32 -> [C,2w,4w] fp32 (bicubic, zeros outside the face)."""
    cube = np.asarray(cube, dtype=np.float32)
    _, C, w, _ = cube.shape
    x0, y0, fx, fy = cubic_plan(out_coord)
    tab = cubic_table()
    wx, wy = tab[fx], tab[fy]                                   # [2w,4w,4]
    f = np.asarray(face_map).astype(np.int64)
    lim = max(w - 3, 0)
    inside = (x0 >= 0) & (x0 < lim) & (y0 >= 0) & (y0 < lim)
    zero = np.float32(0)
    acc_in = None                                               # all-taps-inside order
    acc_bd = np.zeros((C,) + f.shape, dtype=np.float32)         # border order
    for i in range(4):
        row = None
        yy = y0 + i
        for j in range(4):
            xx = x0 + j
            ok = (yy >= 0) & (yy < w) & (xx >= 0) & (xx < w)
            v = cube[f, :, np.clip(yy, 0, w - 1), np.clip(xx, 0, w - 1)]          # [2w,4w,C]
            wt = (wy[..., i] * wx[..., j]).astype(np.float32)
            p = np.transpose(v * wt[..., None], (2, 0, 1))
            row = p if row is None else row + p
            acc_bd = np.where(ok[None], acc_bd + p, acc_bd)
        acc_in = row if acc_in is None else acc_in + row
    return np.where(inside[None], acc_in, acc_bd + zero).astype(np.float32)
