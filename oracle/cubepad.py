"""Oracle: CubePad as a closed-form index map (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates model/cube_pad.py:95-216 of the reference. Instead of the reference's chain of
cat / index_select / permute / repeat, every output pixel is given its source
(face, row, col); applying the map is a single fancy-index gather.

Face order (cube_pad.py:49,106-111): 0=Back 1=Down 2=Front 3=Left 4=Right 5=Top.
"""
import numpy as np

B, D, F, L, R, T = range(6)


def get_pad_size(lrtd_pad):
    """cube_pad.py:12-20 — int (all four sides) or [p_l, p_r, p_t, p_d]."""
    if isinstance(lrtd_pad, (int, np.integer)):
        p = int(lrtd_pad)
        return p, p, p, p
    p_l, p_r, p_t, p_d = (int(v) for v in lrtd_pad)
    return p_l, p_r, p_t, p_d


def _strip_top(f, r, c, H, W, pt):
    """(face, row, col) of top-strip element (r in [0,pt), c in [0,W)). cube_pad.py:114-126."""
    if f == B:
        return T, r, W - 1 - c          # flip(top[:pt], cols)
    if f == D:
        return F, H - pt + r, c         # front[-pt:]
    if f == F:
        return T, H - pt + r, c         # top[-pt:]
    if f == L:
        return T, c, r                  # top[:, :pt].T
    if f == R:
        return T, H - 1 - c, W - pt + r  # flip(top[:, -pt:].T, cols)
    return B, r, W - 1 - c              # f == T: flip(back[:pt], cols)


def _strip_down(f, r, c, H, W, pd):
    """Down strip (r in [0,pd), c in [0,W)). cube_pad.py:127-138."""
    if f == B:
        return D, H - pd + r, W - 1 - c
    if f == D:
        return B, H - pd + r, W - 1 - c
    if f == F:
        return D, r, c
    if f == L:
        return D, H - 1 - c, r
    if f == R:
        return D, c, W - pd + r
    return F, r, c


def _strip_left(f, r, c, H, W, pl):
    """Left strip (r in [0,H), c in [0,pl)). cube_pad.py:139-150."""
    if f == B:
        return R, r, W - pl + c
    if f == D:
        return L, H - pl + c, W - 1 - r
    if f == F:
        return L, r, W - pl + c
    if f == L:
        return B, r, W - pl + c
    if f == R:
        return F, r, W - pl + c
    return L, c, r


def _strip_right(f, r, c, H, W, pr):
    """Right strip (r in [0,H), c in [0,pr)). cube_pad.py:151-162."""
    if f == B:
        return L, r, c
    if f == D:
        return R, H - pr + c, r
    if f == F:
        return R, r, c
    if f == L:
        return F, r, c
    if f == R:
        return B, r, c
    return R, c, W - 1 - r


def source_of(f, oy, ox, H, W, pl, pr, pt, pd):
    """Source (face,row,col) of output pixel (oy,ox) of face f. Scalar, slow, obviously right."""
    y, x = oy - pt, ox - pl
    in_y, in_x = 0 <= y < H, 0 <= x < W
    if in_y and in_x:
        return f, y, x
    if in_x:
        return _strip_top(f, oy, x, H, W, pt) if y < 0 else _strip_down(f, y - H, x, H, W, pd)
    if in_y:
        return _strip_left(f, y, ox, H, W, pl) if x < 0 else _strip_right(f, y, x - W, H, W, pr)
    # corner — make_cubepad_edge, cube_pad.py:83-90,165-176
    top, left = y < 0, x < 0
    td = pt if top else pd
    lr = pl if left else pr
    if td > lr:
        # repeat the left/right strip's first (top) / last (bottom) row down the corner
        rr = 0 if top else H - 1
        return (_strip_left(f, rr, ox, H, W, pl) if left
                else _strip_right(f, rr, x - W, H, W, pr))
    # repeat the top/down strip's edge column across the corner
    cc = 0 if left else W - 1
    return (_strip_top(f, oy, cc, H, W, pt) if top
            else _strip_down(f, y - H, cc, H, W, pd))


def index_map(H, W, lrtd_pad):
    """int64 [6, Ho, Wo]: flat index into one channel's [6, H, W] cube for every output pixel."""
    if H != W:
        raise ValueError("CubePad needs square faces (reference cat fails for H != W)")
    pl, pr, pt, pd = get_pad_size(lrtd_pad)
    if max(pl, pr, pt, pd) > H or min(pl, pr, pt, pd) < 0:
        raise ValueError("pad must be in [0, H]")
    Ho, Wo = H + pt + pd, W + pl + pr
    m = np.empty((6, Ho, Wo), dtype=np.int64)
    for f in range(6):
        # interior in one shot, the O(perimeter) halo per element
        yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        m[f, pt:pt + H, pl:pl + W] = f * H * W + yy * W + xx
        for oy in range(Ho):
            inner_row = pt <= oy < pt + H
            cols = (list(range(pl)) + list(range(pl + W, Wo))) if inner_row else range(Wo)
            for ox in cols:
                sf, sy, sx = source_of(f, oy, ox, H, W, pl, pr, pt, pd)
                m[f, oy, ox] = sf * H * W + sy * W + sx
    return m


def cubepad(x, lrtd_pad, imap=None):
    """x: ndarray [6N, C, H, W] -> [6N, C, Ho, Wo] (same dtype). cube_pad.py:28-42."""
    x = np.asarray(x)
    n6, C, H, W = x.shape
    if n6 % 6:
        raise ValueError("CubePad size mismatch!")  # cube_pad.py:33-35 prints + exit()
    if imap is None:
        imap = index_map(H, W, lrtd_pad)
    _, Ho, Wo = imap.shape
    g = x.reshape(n6 // 6, 6, C, H * W).transpose(0, 2, 1, 3).reshape(n6 // 6, C, 6 * H * W)
    out = g[:, :, imap.reshape(-1)]                       # [N, C, 6*Ho*Wo]
    out = out.reshape(n6 // 6, C, 6, Ho, Wo).transpose(0, 2, 1, 3, 4)
    return np.ascontiguousarray(out).reshape(n6, C, Ho, Wo)
