// CubePad geometry as a 4x6 table of affine source maps (host + device).
//
// The reference builds the four padding plates of each face by slicing / transposing /
// flipping a neighbouring face (model/cube_pad.py:114-162) and the corners by repeating a
// plate edge (make_cubepad_edge, :83-90, :165-176). Every plate element is therefore
//     src = (face', y', x')  with  y' = yr*r + yc*c + y0,   x' = xr*r + xc*c + x0
// where (r, c) are the element's coordinates inside the plate and the coefficients depend only
// on (plate, face, H, W, pads). The table is filled once per launch on the host and handed to
// the kernels by value; the kernels never branch on the face id.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CP360_HD __host__ __device__ __forceinline__
#else
#define CP360_HD inline
#endif

namespace cp360 {

enum Face { FB = 0, FD = 1, FF = 2, FL = 3, FR = 4, FT = 5 };
enum Plate { P_TOP = 0, P_DOWN = 1, P_LEFT = 2, P_RIGHT = 3 };

struct PlateMap {      // source = (face, pix) with pix = base + sr*r + sc*c  (pix = y'*W + x')
  int32_t face;        // face'
  int32_t base;        // y0*W + x0
  int32_t sr;          // yr*W + xr
  int32_t sc;          // yc*W + xc
};

// The same geometry seen from the SOURCE side ("push" form, used by the row kernel): face f is the
// source of exactly four plates of other faces. Entry k of face f describes one of them as a
// rectangle of destination elements (u, v) in [0,U) x [0,V) at (oy0 + u, ox0 + v) of face `dface`
// — the plate plus the corner elements that replicate its edge — and the affine map from plate
// coordinates to the source pixel:
//     cmode 0 (top/down plates):   r = u,                        c = clamp(v + off, 0, W-1)
//     cmode 1 (left/right plates): r = clamp(u + off, 0, H-1),   c = v
//     y' = yr*r + yc*c + y0,  x' = xr*r + xc*c + x0
// Exactly one of yr, yc is +-1: `drive` tells whether the source ROW is a function of u (0) or
// of v (1), so the destination elements fed by a band of source rows form a sub-rectangle.
struct PushEntry {
  int32_t dface, oy0, ox0, U, V;
  int32_t cmode, off;
  int32_t yr, yc, y0, xr, xc, x0;
  int32_t drive, L, I, doff;   // driving index i: w = clamp(i + doff, 0, L-1), i in [0, I)
};

struct CubePadGeom {
  int32_t H, W, pl, pr, pt, pd, Ho, Wo;
  int32_t corner_uses_lr[4];   // [tl, tr, dl, dr]: 1 -> corner repeats the l/r plate row, 0 -> t/d plate column
  PlateMap plate[4][6];
  PushEntry push[6][4];
};

// rows: what the reference slices for each face, written as (face', y(r,c), x(r,c)).
// y = yr*r + yc*c + y0 ; x = xr*r + xc*c + x0
struct AffineSrc { int face, yr, yc, y0, xr, xc, x0; };

inline PlateMap make_plate(const AffineSrc& a, int W) {
  PlateMap m;
  m.face = a.face;
  m.base = a.y0 * W + a.x0;
  m.sr = a.yr * W + a.xr;
  m.sc = a.yc * W + a.xc;
  return m;
}

// Fills the table. Returns false for shapes the reference cannot pad (H != W, pad > H, pad < 0).
inline bool make_geom(int H, int W, int pl, int pr, int pt, int pd, CubePadGeom* g) {
  if (H <= 0 || W <= 0 || H != W) return false;
  if (pl < 0 || pr < 0 || pt < 0 || pd < 0) return false;
  if (pl > W || pr > W || pt > H || pd > H) return false;
  g->H = H; g->W = W; g->pl = pl; g->pr = pr; g->pt = pt; g->pd = pd;
  g->Ho = H + pt + pd; g->Wo = W + pl + pr;
  // top plate, r in [0,pt), c in [0,W)                      cube_pad.py:114-126
  const AffineSrc top[6] = {
      /*B*/ {FT, 1, 0, 0, 0, -1, W - 1},          // flip(top[:pt, :])
      /*D*/ {FF, 1, 0, H - pt, 0, 1, 0},          // front[-pt:, :]
      /*F*/ {FT, 1, 0, H - pt, 0, 1, 0},          // top[-pt:, :]
      /*L*/ {FT, 0, 1, 0, 1, 0, 0},               // top[:, :pt]^T
      /*R*/ {FT, 0, -1, H - 1, 1, 0, W - pt},     // flip(top[:, -pt:]^T)
      /*T*/ {FB, 1, 0, 0, 0, -1, W - 1}};         // flip(back[:pt, :])
  // down plate, r in [0,pd), c in [0,W)                     cube_pad.py:127-138
  const AffineSrc down[6] = {
      /*B*/ {FD, 1, 0, H - pd, 0, -1, W - 1},     // flip(down[-pd:, :])
      /*D*/ {FB, 1, 0, H - pd, 0, -1, W - 1},     // flip(back[-pd:, :])
      /*F*/ {FD, 1, 0, 0, 0, 1, 0},               // down[:pd, :]
      /*L*/ {FD, 0, -1, H - 1, 1, 0, 0},          // flip(down[:, :pd]^T)
      /*R*/ {FD, 0, 1, 0, 1, 0, W - pd},          // down[:, -pd:]^T
      /*T*/ {FF, 1, 0, 0, 0, 1, 0}};              // front[:pd, :]
  // left plate, r in [0,H), c in [0,pl)                     cube_pad.py:139-150
  const AffineSrc left[6] = {
      /*B*/ {FR, 1, 0, 0, 0, 1, W - pl},          // right[:, -pl:]
      /*D*/ {FL, 0, 1, H - pl, -1, 0, W - 1},     // flip(left[-pl:, :]^T, rows)
      /*F*/ {FL, 1, 0, 0, 0, 1, W - pl},          // left[:, -pl:]
      /*L*/ {FB, 1, 0, 0, 0, 1, W - pl},          // back[:, -pl:]
      /*R*/ {FF, 1, 0, 0, 0, 1, W - pl},          // front[:, -pl:]
      /*T*/ {FL, 0, 1, 0, 1, 0, 0}};              // left[:pl, :]^T
  // right plate, r in [0,H), c in [0,pr)                    cube_pad.py:151-162
  const AffineSrc right[6] = {
      /*B*/ {FL, 1, 0, 0, 0, 1, 0},               // left[:, :pr]
      /*D*/ {FR, 0, 1, H - pr, 1, 0, 0},          // right[-pr:, :]^T
      /*F*/ {FR, 1, 0, 0, 0, 1, 0},               // right[:, :pr]
      /*L*/ {FF, 1, 0, 0, 0, 1, 0},               // front[:, :pr]
      /*R*/ {FB, 1, 0, 0, 0, 1, 0},               // back[:, :pr]
      /*T*/ {FR, 0, 1, 0, -1, 0, W - 1}};         // flip(right[:pr, :]^T, rows)
  for (int f = 0; f < 6; ++f) {
    g->plate[P_TOP][f] = make_plate(top[f], W);
    g->plate[P_DOWN][f] = make_plate(down[f], W);
    g->plate[P_LEFT][f] = make_plate(left[f], W);
    g->plate[P_RIGHT][f] = make_plate(right[f], W);
  }
  // make_cubepad_edge: td_pad > lr_pad -> repeat the l/r plate's row, else the t/d plate's column
  g->corner_uses_lr[0] = pt > pl;
  g->corner_uses_lr[1] = pt > pr;
  g->corner_uses_lr[2] = pd > pl;
  g->corner_uses_lr[3] = pd > pr;
  // push table
  int n_push[6] = {0, 0, 0, 0, 0, 0};
  const AffineSrc* tables[4] = {top, down, left, right};
  const int* cl = g->corner_uses_lr;
  for (int P = 0; P < 4; ++P)
    for (int f = 0; f < 6; ++f) {
      const AffineSrc& a = tables[P][f];
      if (n_push[a.face] >= 4) return false;   // cannot happen: every face feeds exactly 4 plates
      PushEntry& e = g->push[a.face][n_push[a.face]++];
      e.dface = f;
      e.yr = a.yr; e.yc = a.yc; e.y0 = a.y0; e.xr = a.xr; e.xc = a.xc; e.x0 = a.x0;
      if (P == P_TOP || P == P_DOWN) {
        const int k0 = P == P_TOP ? 0 : 2;                 // corners [tl,tr] or [dl,dr]
        const int xa = cl[k0] ? pl : 0, xb = cl[k0 + 1] ? pl + W : g->Wo;
        e.oy0 = P == P_TOP ? 0 : pt + H; e.U = P == P_TOP ? pt : pd;
        e.ox0 = xa; e.V = xb - xa;
        e.cmode = 0; e.off = xa - pl;
      } else {
        const int kt = P == P_LEFT ? 0 : 1, kd = kt + 2;   // corners [tl,dl] or [tr,dr]
        const int ya = cl[kt] ? 0 : pt, yb = cl[kd] ? g->Ho : pt + H;
        e.ox0 = P == P_LEFT ? 0 : pl + W; e.V = P == P_LEFT ? pl : pr;
        e.oy0 = ya; e.U = yb - ya;
        e.cmode = 1; e.off = ya - pt;
      }
      e.drive = a.yr != 0 ? 0 : 1;
      e.I = e.drive == 0 ? e.U : e.V;
      const bool clamped = (e.drive == 0) == (e.cmode == 1);
      e.L = clamped ? (e.cmode == 1 ? H : W) : e.I;
      e.doff = clamped ? e.off : 0;
    }
  for (int f = 0; f < 6; ++f)
    if (n_push[f] != 4) return false;
  return true;
}

// Source of output pixel (oy, ox) of face f: returns the pixel offset y'*W + x' inside the
// source face plane and writes the source face id to *src_face.
CP360_HD int32_t cubepad_src(const CubePadGeom& g, int f, int oy, int ox, int* src_face) {
  const int y = oy - g.pt, x = ox - g.pl;
  const bool top = y < 0, bot = y >= g.H, lft = x < 0, rgt = x >= g.W;
  if (!(top | bot | lft | rgt)) {
    *src_face = f;
    return y * g.W + x;
  }
  int plate, r, c;
  if (!(lft | rgt)) {                       // top / down plate
    plate = top ? P_TOP : P_DOWN; r = top ? oy : y - g.H; c = x;
  } else if (!(top | bot)) {                // left / right plate
    plate = lft ? P_LEFT : P_RIGHT; r = y; c = lft ? ox : x - g.W;
  } else {                                  // corner
    const int k = (bot ? 2 : 0) + (rgt ? 1 : 0);
    if (g.corner_uses_lr[k]) {
      plate = lft ? P_LEFT : P_RIGHT; r = top ? 0 : g.H - 1; c = lft ? ox : x - g.W;
    } else {
      plate = top ? P_TOP : P_DOWN; r = top ? oy : y - g.H; c = lft ? 0 : g.W - 1;
    }
  }
  const PlateMap& m = g.plate[plate][f];
  *src_face = m.face;
  return m.base + m.sr * r + m.sc * c;
}

// The transpose of the map above (backward pass, tests): calls fn(dface, oy, ox) for every HALO
// output position of the cube that copies source pixel (f, y, x) — the four push entries of face f
// inverted. Entry e covers destination (u, v) in [0,U) x [0,V) with
//   cmode 0: r = u, c = clamp(v + off, 0, W-1);   cmode 1: r = clamp(u + off, 0, H-1), c = v
//   y' = yr*r + yc*c + y0,  x' = xr*r + xc*c + x0   (a signed permutation: one of yr, yc is +-1)
// so (r, c) follows from (y, x) directly and a clamped coordinate at 0 / its maximum stands for the
// whole run of corner positions that replicate it. Positions are visited in a fixed order
// (entry, u, v ascending), which makes a sum over them reproducible.
template <class F>
CP360_HD void cubepad_for_each_copy(const CubePadGeom& g, int f, int y, int x, F&& fn) {
  for (int e = 0; e < 4; ++e) {
    const PushEntry& pe = g.push[f][e];
    if (pe.U <= 0 || pe.V <= 0) continue;
    int r, c;
    if (pe.yr != 0) { r = (y - pe.y0) * pe.yr; c = (x - pe.x0) * pe.xc; }
    else { c = (y - pe.y0) * pe.yc; r = (x - pe.x0) * pe.xr; }
    int ulo, uhi, vlo, vhi;
    if (pe.cmode == 0) {
      if (r < 0 || r >= pe.U || c < 0 || c > g.W - 1) continue;
      ulo = uhi = r;
      vlo = c == 0 ? 0 : c - pe.off;
      vhi = c == g.W - 1 ? pe.V - 1 : c - pe.off;
    } else {
      if (c < 0 || c >= pe.V || r < 0 || r > g.H - 1) continue;
      vlo = vhi = c;
      ulo = r == 0 ? 0 : r - pe.off;
      uhi = r == g.H - 1 ? pe.U - 1 : r - pe.off;
    }
    if (ulo < 0) ulo = 0;
    if (vlo < 0) vlo = 0;
    if (uhi > pe.U - 1) uhi = pe.U - 1;
    if (vhi > pe.V - 1) vhi = pe.V - 1;
    for (int u = ulo; u <= uhi; ++u)
      for (int v = vlo; v <= vhi; ++v) fn(pe.dface, pe.oy0 + u, pe.ox0 + v);
  }
}

}  // namespace cp360
