// CubePad row kernel (32-bit elements) — included by cubepad.cu.
//
// Work unit ("tile"): a band of consecutive output rows of ONE plane, or k whole planes when
// planes are small. Either way the input rows it copies are one contiguous, 16 B-alignable range
// of the NCHW tensor, fetched by ONE TMA bulk load, and every input element is read from DRAM
// exactly once.
//
// Every warp is its own pipeline: a private ring of `slots` shared-memory buffers with one
// mbarrier each. Lane 0 keeps `slots` bulk loads in flight; when a tile has landed the warp
// copies its rows — LDS.32 (conflict-free) -> coalesced STG.32; the (pl + row*Wo)-word shift
// between the two layouts is what keeps this from being a pure TMA copy, and the (W+2p)*4 B
// output pitch is why the store side cannot be a tensor map — and immediately re-arms the slot
// with the tile `slots` steps ahead. There is no __syncthreads, no producer/consumer handshake
// and no cross-warp dependency anywhere in the steady state.
// The few elements that are not a straight copy (side columns, top/bottom pad rows, corners) are
// gathered from L2 through the affine plate table; their loads are issued BEFORE the warp waits
// for its tile and consumed after the rows are copied, so their latency hides behind the copy.
#pragma once
#include "common.cuh"
#include "cubepad_geom.h"
#include "tma.cuh"

namespace cp360 {

constexpr int kRowWarps = 8;
constexpr int kRowThreads = kRowWarps * 32;
constexpr int kRowMaxSlots = 8;
constexpr int kRowBarBytes = kRowWarps * kRowMaxSlots * 8;

struct RowArgs {
  const uint32_t* x;
  uint32_t* y;
  int64_t total_in_words;   // n_planes * H * W
  int32_t n_planes;
  int32_t n_tiles;
  int32_t C;
  int32_t nb;               // bands per plane (1 when a tile is k whole planes)
  int32_t Rb;               // output rows per band
  int32_t k;                // planes per tile (1 when nb > 1)
  int32_t slot_words;       // capacity of one ring slot
  int32_t slots;
  int32_t any_corner_lr;    // some corner repeats the l/r plate (asymmetric pads only)
};

struct RowTile {
  int32_t p0, np;           // planes [p0, p0 + np)
  int32_t oyA, oyB;         // output rows [oyA, oyB) of each of them
  int32_t ya, yb;           // interior input rows [ya, yb)
  int32_t shift;            // word offset of row ya inside the slot
  int32_t words;            // words to load (multiple of 4; 0 = nothing to copy)
  int64_t w0;               // first word to load (multiple of 4)
};

__device__ __forceinline__ RowTile row_tile(const RowArgs& a, const CubePadGeom& g, int t) {
  RowTile d;
  if (a.nb > 1) {
    d.p0 = t / a.nb;
    const int b = t - d.p0 * a.nb;
    d.np = 1;
    d.oyA = b * a.Rb;
    d.oyB = min(d.oyA + a.Rb, g.Ho);
  } else {
    d.p0 = t * a.k;
    d.np = min(a.k, a.n_planes - d.p0);
    d.oyA = 0;
    d.oyB = g.Ho;
  }
  d.ya = min(max(d.oyA - g.pt, 0), g.H);
  d.yb = min(max(d.oyB - g.pt, 0), g.H);
  const int HW = g.H * g.W;
  const int64_t s0 = (int64_t)d.p0 * HW + d.ya * g.W;
  const int64_t s1 = (int64_t)(d.p0 + d.np - 1) * HW + d.yb * g.W;
  d.w0 = s0 & ~(int64_t)3;
  const int64_t w1 = min((s1 + 3) & ~(int64_t)3, a.total_in_words);
  d.shift = (int)(s0 - d.w0);
  d.words = s1 > s0 ? (int)(w1 - d.w0) : 0;
  return d;
}

// Source of a top / bottom pad-row element (incl. corners).
__device__ __forceinline__ int row_pad_src(const RowArgs& a, const CubePadGeom& g, int f, int oy,
                                           int ox, int* sf) {
  if (a.any_corner_lr) return cubepad_src(g, f, oy, ox, sf);
  const bool top = oy < g.pt;
  const PlateMap& m = g.plate[top ? P_TOP : P_DOWN][f];
  const int r = top ? oy : oy - g.pt - g.H;
  const int cc = min(max(ox - g.pl, 0), g.W - 1);      // corners repeat the plate's edge column
  *sf = m.face;
  return m.base + m.sr * r + m.sc * cc;
}

__global__ void __launch_bounds__(kRowThreads)
cubepad_row_kernel(const RowArgs a, const __grid_constant__ CubePadGeom g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw) + warp * kRowMaxSlots;
  uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + kRowBarBytes) +
                   (size_t)warp * a.slots * a.slot_words;
  const int H = g.H, W = g.W, Ho = g.Ho, Wo = g.Wo, HW = H * W, HoWo = Ho * Wo;
  const int slots = a.slots;
  const int gw = blockIdx.x * kRowWarps + warp, GW = gridDim.x * kRowWarps;
  const int64_t face_stride = (int64_t)a.C * HW;
  const int nside = g.pl + g.pr;

  auto issue = [&](int t, int s) {
    const RowTile d = row_tile(a, g, t);
    tma::mbar_expect_tx(&bar[s], (uint32_t)d.words * 4u);
    if (d.words) tma::bulk_load(ring + (size_t)s * a.slot_words, a.x + d.w0, (uint32_t)d.words * 4u, &bar[s]);
  };

  if (lane == 0) {
    for (int s = 0; s < slots; ++s) tma::mbar_init(&bar[s], 1);
    tma::fence_mbar_init();
    int t = gw;
    for (int s = 0; s < slots && t < a.n_tiles; ++s, t += GW) issue(t, s);
  }
  __syncwarp();

  int it = 0;
  for (int t = gw; t < a.n_tiles; t += GW, ++it) {
    const int s = it % slots;
    const RowTile d = row_tile(a, g, t);
    const uint32_t* in_s = ring + (size_t)s * a.slot_words + d.shift;
    int nf = d.p0 / a.C, c = d.p0 - nf * a.C, f = nf % 6;
    const int ntop = max(min(d.oyB, g.pt) - d.oyA, 0);
    const int bot0 = max(d.oyA, g.pt + H), nbot = max(d.oyB - bot0, 0);
    const int n_side = (d.yb - d.ya) * nside, n_pad = (ntop + nbot) * Wo;

    uint32_t hv0 = 0, hv1 = 0;                 // gathered values in flight
    uint32_t *hp0 = nullptr, *hp1 = nullptr;   // where they go
    bool landed = false;
    for (int j = 0; j < d.np; ++j) {
      const uint32_t* __restrict__ cube = a.x + ((int64_t)(nf - f) * a.C + c) * HW;
      uint32_t* __restrict__ outp = a.y + (int64_t)(d.p0 + j) * HoWo;

      // ---- A. side columns of the interior rows: lanes = (row, side column)
      for (int q0 = 0; q0 < n_side; q0 += 32) {
        if (hp0) { __stcs(hp0, hv0); hp0 = nullptr; }
        const int q = q0 + lane;
        if (q < n_side) {
          const int i = q / nside, sc = q - i * nside, y = d.ya + i;
          const bool left = sc < g.pl;
          const PlateMap& m = g.plate[left ? P_LEFT : P_RIGHT][f];
          const int pix = m.base + m.sr * y + m.sc * (left ? sc : sc - g.pl);
          hv0 = __ldg(cube + m.face * face_stride + pix);
          hp0 = outp + (y + g.pt) * Wo + (left ? sc : W + sc);
        }
      }
      // ---- B. top / bottom pad rows (incl. corners) of this band
      for (int q0 = 0; q0 < n_pad; q0 += 32) {
        if (hp1) { __stcs(hp1, hv1); hp1 = nullptr; }
        const int q = q0 + lane;
        if (q < n_pad) {
          const int rr = q / Wo, ox = q - rr * Wo;
          const int oy = rr < ntop ? d.oyA + rr : bot0 + (rr - ntop);
          int sf;
          const int pix = row_pad_src(a, g, f, oy, ox, &sf);
          hv1 = __ldg(cube + sf * face_stride + pix);
          hp1 = outp + oy * Wo + ox;
        }
      }
      // ---- C. interior rows: shifted copy shared -> global
      if (!landed) { tma::mbar_wait(&bar[s], (uint32_t)((it / slots) & 1)); landed = true; }
      const uint32_t* __restrict__ sp = in_s + j * HW + lane;
      uint32_t* __restrict__ dp = outp + (d.ya + g.pt) * Wo + g.pl + lane;
      const int nr = d.yb - d.ya;
      int r = 0;
      if (W >= 128) {
        for (; r < nr; ++r, sp += W, dp += Wo) {
          int jj = 0;
          for (; jj + 96 + lane < W; jj += 128) {
            const uint32_t v0 = sp[jj], v1 = sp[jj + 32], v2 = sp[jj + 64], v3 = sp[jj + 96];
            __stcs(dp + jj, v0); __stcs(dp + jj + 32, v1); __stcs(dp + jj + 64, v2); __stcs(dp + jj + 96, v3);
          }
          for (; jj + lane < W; jj += 32) __stcs(dp + jj, sp[jj]);
        }
      } else {
        for (; r + 4 <= nr; r += 4, sp += 4 * W, dp += 4 * Wo) {
          for (int jj = 0; jj + lane < W; jj += 32) {
            const uint32_t v0 = sp[jj], v1 = sp[jj + W], v2 = sp[jj + 2 * W], v3 = sp[jj + 3 * W];
            __stcs(dp + jj, v0); __stcs(dp + jj + Wo, v1); __stcs(dp + jj + 2 * Wo, v2); __stcs(dp + jj + 3 * Wo, v3);
          }
        }
        for (; r < nr; ++r, sp += W, dp += Wo)
          for (int jj = 0; jj + lane < W; jj += 32) __stcs(dp + jj, sp[jj]);
      }
      if (++c == a.C) { c = 0; ++nf; if (++f == 6) f = 0; }
    }
    if (!landed) tma::mbar_wait(&bar[s], (uint32_t)((it / slots) & 1));   // keep the phases in step
    __syncwarp();                                  // every lane is done reading slot s
    if (lane == 0) {
      const int tn = t + slots * GW;
      if (tn < a.n_tiles) issue(tn, s);
    }
    if (hp0) __stcs(hp0, hv0);
    if (hp1) __stcs(hp1, hv1);
  }
}

}  // namespace cp360
