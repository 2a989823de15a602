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
// Integer divisions by run-time constants use host-computed multiply-shift pairs.
#pragma once
#include "common.cuh"
#include "cubepad_geom.h"
#include "tma.cuh"

namespace cp360 {

constexpr int kRowWarps = 8;
constexpr int kRowThreads = kRowWarps * 32;
constexpr int kRowMaxSlots = 8;
constexpr int kRowBarBytes = kRowWarps * kRowMaxSlots * 8;

// n / d for 0 <= n < 2^31 as umulhi(n, m) >> s  (m == 0: d == 1)
struct FastDiv { uint32_t m, s; };

inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f = {0u, 0u};
  if (d <= 1) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;                       // smallest l with 2^l >= d
  f.m = (uint32_t)(((1ull << (31 + l)) / d) + 1);    // < 2^32 because d > 2^(l-1)
  f.s = l - 1;
  return f;
}

__device__ __forceinline__ int fdiv(int n, const FastDiv& d) {
  return d.m ? (int)(__umulhi((uint32_t)n, d.m) >> d.s) : n;
}

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
  FastDiv d_nb, d_C, d_nside, d_Wo;
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
    d.p0 = fdiv(t, a.d_nb);
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

// nr rows of W words: sp (shared, row pitch W) -> dp (global, row pitch Wo); both already
// offset by the lane. NJ = ceil(W / 32) at compile time (0: run-time loop); FULL: W % 32 == 0.
template <int NJ, bool FULL>
__device__ __forceinline__ void row_copy(const uint32_t* __restrict__ sp, uint32_t* __restrict__ dp,
                                         int nr, int W, int Wo, int lane) {
  if (NJ == 0) {
    for (int r = 0; r < nr; ++r, sp += W, dp += Wo)
      for (int jj = 0; jj + lane < W; jj += 32) __stcs(dp + jj, sp[jj]);
    return;
  }
  constexpr int NJc = NJ > 0 ? NJ : 1;
  const bool tail_ok = FULL || (NJc - 1) * 32 + lane < W;
  if (NJc >= 4) {
    // wide rows: one row per step, NJ loads in flight
#pragma unroll 1
    for (int r = 0; r < nr; ++r, sp += W, dp += Wo) {
      uint32_t v[NJc];
#pragma unroll
      for (int j = 0; j < NJc; ++j)
        if (j < NJc - 1 || tail_ok) v[j] = sp[j * 32];
#pragma unroll
      for (int j = 0; j < NJc; ++j)
        if (j < NJc - 1 || tail_ok) __stcs(dp + j * 32, v[j]);
    }
  } else {
    // narrow rows: four rows per step
    int r = 0;
#pragma unroll 1
    for (; r + 4 <= nr; r += 4, sp += 4 * W, dp += 4 * Wo) {
      uint32_t v[4][NJc];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < NJc; ++j)
          if (j < NJc - 1 || tail_ok) v[k][j] = sp[k * W + j * 32];
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < NJc; ++j)
          if (j < NJc - 1 || tail_ok) __stcs(dp + k * Wo + j * 32, v[k][j]);
    }
#pragma unroll 1
    for (; r < nr; ++r, sp += W, dp += Wo) {
#pragma unroll
      for (int j = 0; j < NJc; ++j)
        if (j < NJc - 1 || tail_ok) __stcs(dp + j * 32, sp[j * 32]);
    }
  }
}

template <int NJ, bool FULL>
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
    if (d.words) tma::bulk_load(ring + s * a.slot_words, a.x + d.w0, (uint32_t)d.words * 4u, &bar[s]);
  };

  if (lane == 0) {
    for (int s = 0; s < slots; ++s) tma::mbar_init(&bar[s], 1);
    tma::fence_mbar_init();
    int t = gw;
    for (int s = 0; s < slots && t < a.n_tiles; ++s, t += GW) issue(t, s);
  }
  __syncwarp();

  int s = 0;
  uint32_t phase = 0;
#pragma unroll 1
  for (int t = gw; t < a.n_tiles; t += GW) {
    const RowTile d = row_tile(a, g, t);
    const uint32_t* in_s = ring + s * a.slot_words + d.shift;
    int nf = fdiv(d.p0, a.d_C), c = d.p0 - nf * a.C, f = nf % 6;
    const int ntop = max(min(d.oyB, g.pt) - d.oyA, 0);
    const int bot0 = max(d.oyA, g.pt + H), nbot = max(d.oyB - bot0, 0);
    const int n_side = (d.yb - d.ya) * nside, n_pad = (ntop + nbot) * Wo;

    const int n_halo = n_side + n_pad;
    bool landed = false;
#pragma unroll 1
    for (int j = 0; j < d.np; ++j) {
      const uint32_t* __restrict__ cube = a.x + ((int64_t)(nf - f) * a.C + c) * HW;
      uint32_t* __restrict__ outp = a.y + (int64_t)(d.p0 + j) * HoWo;

      // ---- A. everything that is not a straight row copy — side columns of the interior rows,
      //         then the top / bottom pad rows incl. corners — as one list, four gathers per lane
      //         in flight. The first batch is issued before the tile is awaited and stored after
      //         its rows are copied.
      uint32_t hv[4];
      int ho[4];
      auto gather = [&](int q0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = q0 + u * 32 + lane;
          ho[u] = -1;
          if (q < n_side) {
            const int i = fdiv(q, a.d_nside), sc = q - i * nside, y = d.ya + i;
            const bool left = sc < g.pl;
            const PlateMap& m = g.plate[left ? P_LEFT : P_RIGHT][f];
            const int pix = m.base + m.sr * y + m.sc * (left ? sc : sc - g.pl);
            hv[u] = __ldg(cube + m.face * face_stride + pix);
            ho[u] = (y + g.pt) * Wo + (left ? sc : W + sc);
          } else if (q < n_halo) {
            const int qp = q - n_side;
            const int rr = fdiv(qp, a.d_Wo), ox = qp - rr * Wo;
            const int oy = rr < ntop ? d.oyA + rr : bot0 + (rr - ntop);
            int sf;
            const int pix = row_pad_src(a, g, f, oy, ox, &sf);
            hv[u] = __ldg(cube + sf * face_stride + pix);
            ho[u] = oy * Wo + ox;
          }
        }
      };
      auto flush = [&]() {
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ho[u] >= 0) __stcs(outp + ho[u], hv[u]);
      };
      gather(0);
      // ---- B. interior rows: shifted copy shared -> global
      if (!landed) { tma::mbar_wait(&bar[s], phase); landed = true; }
      row_copy<NJ, FULL>(in_s + j * HW + lane, outp + (d.ya + g.pt) * Wo + g.pl + lane, d.yb - d.ya, W, Wo,
                         lane);
      if (j == d.np - 1) {
        __syncwarp();                              // every lane is done reading slot s:
        if (lane == 0) {                           // re-arm it with the tile `slots` steps ahead
          const int tn = t + slots * GW;
          if (tn < a.n_tiles) issue(tn, s);
        }
      }
      flush();
#pragma unroll 1
      for (int q0 = 128; q0 < n_halo; q0 += 128) { gather(q0); flush(); }
      if (++c == a.C) { c = 0; ++nf; if (++f == 6) f = 0; }
    }
    if (++s == slots) { s = 0; phase ^= 1u; }
  }
}

}  // namespace cp360
