// CubePad row kernel (32-bit elements) — included by cubepad.cu.
//
// Tile: a band of consecutive rows of ONE input plane, or k whole planes when planes are small.
// Either way it is one contiguous, 16 B-alignable range of the NCHW tensor, fetched by ONE TMA
// bulk load, and every input element is read from DRAM exactly once. A warp walks units of `ub`
// consecutive bands of a plane so that per-plane bookkeeping is paid once per unit.
//
// Every warp is its own pipeline: a private ring of `slots` shared-memory buffers with one
// mbarrier each. Lane 0 keeps `slots` bulk loads in flight; when a tile has landed the warp
// copies its rows — LDS.32 (conflict-free) -> coalesced STG.32; the (pl + row*Wo)-word shift
// between the two layouts is what keeps this from being a pure TMA copy, and the (W+2p)*4 B
// output pitch is why the store side cannot be a tensor map — and immediately re-arms the slot
// with the tile `slots` steps ahead. There is no __syncthreads, no producer/consumer handshake
// and no cross-warp dependency anywhere in the steady state.
// Halo elements are PUSHED, not gathered: every halo pixel of the padded tensor is a copy of an
// interior pixel of a neighbouring face, so while a band of face f sits in shared memory the warp
// also writes the halo elements of the (up to four) plates of other faces that are fed by those
// rows (cubepad_geom.h: push table, the plate maps seen from the source side; corners ride along
// as clamped plate coordinates). The kernel therefore issues no global loads besides the bulk
// copies: DRAM traffic is exactly one read of the input and one write of the output.
// Integer divisions by run-time constants use host-computed multiply-shift pairs.
#pragma once
#include "common.cuh"
#include "cubepad_geom.h"
#include "tma.cuh"

namespace cp360 {

#ifdef CP360_ROW_ST_DEFAULT
#define CP360_ROW_ST(ptr, v) (*(ptr) = (v))
#else
#define CP360_ROW_ST(ptr, v) __stcs((ptr), (v))   // streaming store: the output is not re-read here
#endif

// Optional per-channel epilogue fused into the copy (fp32 tensors): v -> act(v * scale[c] + shift[c]),
// the eval-mode BatchNorm affine + ReLU that precedes CubePad in the cubic ResNet
// (model/resnet_cubic.py:89-92). Separate multiply and add, so a numpy fp32 restatement is bit-exact.
struct Epi { float sc, sh; int relu; };

template <bool EPI>
__device__ __forceinline__ uint32_t epi_apply(uint32_t v, const Epi& e) {
  if (!EPI) return v;
  float f = __fadd_rn(__fmul_rn(__uint_as_float(v), e.sc), e.sh);
  if (e.relu) f = f < 0.0f ? 0.0f : f;               // NaN stays NaN, like torch.relu / np.maximum
  return __float_as_uint(f);
}

#ifndef CP360_ROW_WARPS
#define CP360_ROW_WARPS 8                              // measured on B200: 12 / 16 warps per CTA (smaller rings each) are slower — profiles/README.md
#endif
constexpr int kRowWarps = CP360_ROW_WARPS;
constexpr int kRowThreads = kRowWarps * 32;
constexpr int kRowMaxSlots = 8;
constexpr int kRowMetaOff = kRowWarps * kRowMaxSlots * 8;            // per-slot (plane, band) of the tile in flight
constexpr int kRowCtrOff = 2 * kRowMetaOff;                            // CTA-wide unit counter
constexpr int kRowGrpOff = kRowCtrOff + 16;                          // order 4: ring of 8 {group base, tag}
constexpr int kRowBarBytes = kRowGrpOff + 64;

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
  int32_t n_units;
  int32_t C;
  int32_t Rb;               // interior rows per band
  int32_t nb;               // bands per plane           (1 when a tile is k whole planes)
  int32_t upp;              // units per plane (= nb)
  int32_t k;                // planes per tile           (1 when nb > 1)
  int32_t slot_words;       // capacity of one ring slot
  int32_t slots;
  int32_t ring_off;         // byte offset of the rings inside dynamic shared memory
  int32_t order;            // 0: units dealt round-robin to the grid's warps (static); 2: dealt round-robin to
                            //    CTAs in groups of 8, warps of a CTA draw from the CTA's list; 3: every warp
                            //    draws `draw` consecutive units at a time from one grid-wide counter; 4: CTAs
                            //    draw groups of 8 consecutive units from the grid-wide counter, their warps draw
                            //    from the CTA's current group
  int32_t draw;
  int32_t n_static;         // order 5: units below this index are dealt like order 2, the rest (the tail of the
                            //    launch) are drawn one at a time from the grid-wide counter by whichever warp is free
  // output tensor may have more channels than the input (CubePad of a channel-concatenation, written
  // one source at a time: model/clstm.py:57-58): output plane of (face-in-batch nf, channel c) is
  // nf * out_C + out_coff + c
  int32_t out_C, out_coff;
  const float* scale;       // EPI kernels: [C] device, may be nullptr (1) / shift nullptr (0)
  const float* shift;
  int32_t relu;
  uint32_t* work;           // order 3: {next unit, finished CTAs}, zero at launch, zeroed again by the last CTA
  FastDiv d_upp, d_C;
};

// A unit is one tile: band `band` of plane `plane` (nb > 1), or k whole planes starting at `plane`.
__device__ __forceinline__ int2 unit_decode(const RowArgs& a, int u) {
  if (a.nb > 1) {
    const int plane = fdiv(u, a.d_upp);
    int band = u - plane * a.upp;
    if (a.order == 0) {
      // static dealing: a warp's stride is often a multiple of upp; rotate the band order by the
      // plane index so that no warp sits on the first / last bands (plate-row pushes) all launch
      band += plane - fdiv(plane, a.d_upp) * a.upp;
      if (band >= a.upp) band -= a.upp;
    }
    return make_int2(plane, band);
  }
  return make_int2(u * a.k, 0);
}

// Destination sub-rectangle of push entry `pe` fed by source rows [ya, yb): extent `wd` of the
// driving index starting at `i0` (the other index spans its full range). Packed i0 | wd << 16.
__device__ __forceinline__ uint32_t push_range(const PushEntry& pe, int ya, int yb) {
  const int sgn = pe.drive == 0 ? pe.yr : pe.yc;
  int wa = sgn > 0 ? ya - pe.y0 : pe.y0 - yb + 1;
  int wb = sgn > 0 ? yb - pe.y0 : pe.y0 - ya + 1;
  wa = max(wa, 0);
  wb = min(wb, pe.L);
  const int i0 = wa == 0 ? 0 : wa - pe.doff;
  const int i1 = wb == pe.L ? pe.I : wb - pe.doff;
  const int wd = wa < wb ? i1 - i0 : 0;
  return (uint32_t)i0 | ((uint32_t)wd << 16);
}

// nr rows of W words: sp (shared, row pitch W) -> dp (global, row pitch Wo); both already
// offset by the lane. NJ = ceil(W / 32) at compile time (0: run-time loop); FULL: W % 32 == 0.
template <int NJ, bool FULL, bool EPI>
__device__ __forceinline__ void row_copy(const uint32_t* __restrict__ sp, uint32_t* __restrict__ dp,
                                         int nr, int W, int Wo, int lane, const Epi& ep) {
  if (NJ == 0) {
    for (int r = 0; r < nr; ++r, sp += W, dp += Wo)
      for (int jj = 0; jj + lane < W; jj += 32) CP360_ROW_ST(dp + jj, epi_apply<EPI>(sp[jj], ep));
    return;
  }
  constexpr int NJc = NJ > 0 ? NJ : 1;
  const bool tail_ok = FULL || (NJc - 1) * 32 + lane < W;
  if (NJc >= 4) {
    // wide rows: one row per step, NJ loads in flight — two rows per step when an epilogue sits between the
    // load and the store (its dependent ALU chain needs more independent work to hide behind)
    int r = 0;
#ifndef CP360_ROW_EPI_ROWS
#define CP360_ROW_EPI_ROWS 2
#endif
#ifndef CP360_ROW_WIDE_ROWS
#define CP360_ROW_WIDE_ROWS 1
#endif
    if ((EPI && NJc <= 4) || (!EPI && NJc <= 4 && CP360_ROW_WIDE_ROWS > 1)) {
      constexpr int KR = EPI ? CP360_ROW_EPI_ROWS : CP360_ROW_WIDE_ROWS;
#pragma unroll 1
      for (; r + KR <= nr; r += KR, sp += KR * W, dp += KR * Wo) {
        uint32_t v[KR][NJc];
#pragma unroll
        for (int k = 0; k < KR; ++k)
#pragma unroll
          for (int j = 0; j < NJc; ++j)
            if (j < NJc - 1 || tail_ok) v[k][j] = sp[k * W + j * 32];
#pragma unroll
        for (int k = 0; k < KR; ++k)
#pragma unroll
          for (int j = 0; j < NJc; ++j)
            if (j < NJc - 1 || tail_ok) CP360_ROW_ST(dp + k * Wo + j * 32, epi_apply<EPI>(v[k][j], ep));
      }
    }
#pragma unroll 1
    for (; r < nr; ++r, sp += W, dp += Wo) {
      uint32_t v[NJc];
#pragma unroll
      for (int j = 0; j < NJc; ++j)
        if (j < NJc - 1 || tail_ok) v[j] = sp[j * 32];
#pragma unroll
      for (int j = 0; j < NJc; ++j)
        if (j < NJc - 1 || tail_ok) CP360_ROW_ST(dp + j * 32, epi_apply<EPI>(v[j], ep));
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
          if (j < NJc - 1 || tail_ok) CP360_ROW_ST(dp + k * Wo + j * 32, epi_apply<EPI>(v[k][j], ep));
    }
#pragma unroll 1
    for (; r < nr; ++r, sp += W, dp += Wo) {
#pragma unroll
      for (int j = 0; j < NJc; ++j)
        if (j < NJc - 1 || tail_ok) CP360_ROW_ST(dp + j * 32, epi_apply<EPI>(sp[j * 32], ep));
    }
  }
}

template <int NJ, bool FULL, bool EPI>
// one 8-warp CTA per SM (two at most): registers are plentiful — tell ptxas, which otherwise holds the kernel at 64
// registers for occupancy it cannot use and spills inside the copy loops
__global__ void __launch_bounds__(kRowThreads, 1)
cubepad_row_kernel(const RowArgs a, const __grid_constant__ CubePadGeom g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw) + warp * kRowMaxSlots;
  uint4* ptab = reinterpret_cast<uint4*>(smem_raw + kRowBarBytes);          // [6][nb] x 4 entries
  uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw + a.ring_off) +
                   (size_t)warp * a.slots * a.slot_words;
  const int H = g.H, W = g.W, Wo = g.Wo, HW = H * W, HoWo = g.Ho * Wo;
  const int slots = a.slots;
  const int gw = blockIdx.x * kRowWarps + warp, GW = gridDim.x * kRowWarps;
  int2* meta = reinterpret_cast<int2*>(smem_raw + kRowMetaOff) + warp * kRowMaxSlots;
  int* ctr = reinterpret_cast<int*>(smem_raw + kRowCtrOff);
  volatile int* gbase = reinterpret_cast<volatile int*>(smem_raw + kRowGrpOff);   // [8]
  volatile int* gtag = gbase + 8;                                                 // [8] group number held, -1: none

  // lane 0: bulk-load the tile (plane, band) into slot s
  auto issue = [&](int plane, int band, int s) {
    int64_t s0, s1;
    if (a.nb > 1) {
      const int ya = band * a.Rb, yb = min(ya + a.Rb, H);
      s0 = (int64_t)plane * HW + ya * W;
      s1 = (int64_t)plane * HW + yb * W;
    } else {
      s0 = (int64_t)plane * HW;
      s1 = (int64_t)min(plane + a.k, a.n_planes) * HW;
    }
    const int64_t w0 = s0 & ~(int64_t)3;
    const int64_t w1 = min((s1 + 3) & ~(int64_t)3, a.total_in_words);
    const uint32_t bytes = (uint32_t)(w1 - w0) * 4u;
    tma::mbar_expect_tx(&bar[s], bytes);
    tma::bulk_load(ring + s * a.slot_words, a.x + w0, bytes, &bar[s]);
  };
  // lane 0: take the warp's next unit, publish it as slot s's tile and start its load
  int my_k = 0, u_next = 0, u_left = 0;
  uint32_t ticket = 0;                                 // order 3: drawn one step ahead, so that the
  auto arm = [&](int s) {                              // atomic's round trip never sits on the re-arm path
    int u;
    if (a.order == 3) {
      if (u_left == 0) {
        // clamp: the counter keeps growing by `draw` per probe of an exhausted warp
        u_next = (int)min(ticket, (uint32_t)a.n_units);
        u_left = a.draw;
        ticket = atomicAdd(a.work, (uint32_t)a.draw);
      }
      u = u_next++;
      --u_left;
    } else if (a.order == 4) {
      const int k = atomicAdd(ctr, 1);
      const int gl = k >> 3;
      while (gtag[gl & 7] != gl) {}                    // published two groups ahead; practically never spins
      u = gbase[gl & 7] + (k & 7);
      if ((k & 7) == 0) ticket = 1;                    // this warp fetches group gl + 2 once its own load is on its way
      u_next = gl + 2;
    } else if (a.order == 2) {
      const int k = atomicAdd(ctr, 1);
      u = (k / kRowWarps) * GW + blockIdx.x * kRowWarps + (k % kRowWarps);
    } else if (a.order == 5) {
      u = a.n_units;
      if (!u_left) {                                   // u_left doubles as "this warp is in the tail"
        const int k = atomicAdd(ctr, 1);
        u = (k / kRowWarps) * GW + blockIdx.x * kRowWarps + (k % kRowWarps);
        if (u >= a.n_static) u_left = 1;
      }
      if (u_left) u = a.n_static + (int)min(atomicAdd(a.work, 1u), (uint32_t)(a.n_units - a.n_static));
    } else {
      u = gw + my_k * GW;
      ++my_k;
    }
    if (u < a.n_units) {
      const int2 m = unit_decode(a, u);
      meta[s] = m;
      issue(m.x, m.y, s);
    } else {
      meta[s] = make_int2(-1, 0);
    }
    if (a.order == 4 && ticket) {
      ticket = 0;
      const int gb = (int)min(atomicAdd(a.work, 8u), (uint32_t)a.n_units);
      gbase[u_next & 7] = gb;
      __threadfence_block();
      gtag[u_next & 7] = u_next;
    }
  };

  CP360_TRACE_BEGIN(1)
  pdl_trigger();
  CP360_TRACE_INIT_MIN(2);
  if (threadIdx.x == 0) *ctr = 0;
  // push ranges of every (face, band): computed once per CTA, one LDS.128 per tile afterwards
  for (int i = threadIdx.x; i < 6 * a.nb * 4; i += kRowThreads) {
    const int e = i & 3, fb = i >> 2, ff = fb / a.nb, b = fb - ff * a.nb;
    const int ya = a.nb > 1 ? b * a.Rb : 0, yb = a.nb > 1 ? min(ya + a.Rb, H) : H;
    reinterpret_cast<uint32_t*>(ptab)[i] = push_range(g.push[ff][e], ya, yb);
  }
  if (threadIdx.x < 8) gtag[threadIdx.x] = -1;
  __syncthreads();                                     // the only block-wide sync before the end of the kernel
  pdl_wait();
  CP360_TRACE_T0(1);
  if (a.order == 4 && threadIdx.x == 0) {              // groups 0 and 1 of this CTA
    const uint32_t g0 = atomicAdd(a.work, 8u), g1 = atomicAdd(a.work, 8u);
    gbase[0] = (int)min(g0, (uint32_t)a.n_units);
    gbase[1] = (int)min(g1, (uint32_t)a.n_units);
    __threadfence_block();
    gtag[0] = 0;
    gtag[1] = 1;
  }

  if (lane == 0) {
    for (int s = 0; s < slots; ++s) tma::mbar_init(&bar[s], 1);
    tma::fence_mbar_init();
    if (a.order == 3) ticket = atomicAdd(a.work, (uint32_t)a.draw);
    if (a.order == 4) ticket = 0;
    for (int s = 0; s < slots; ++s) arm(s);
  }
  __syncwarp();

  int s = 0;
  uint32_t phase = 0;
#ifdef CP360_TRACE
  bool first_tile = true;
#endif
#pragma unroll 1
  while (true) {
    const int2 m = meta[s];
    if (m.x < 0) break;                                // units are drawn in increasing order: nothing behind it
    const int plane0 = m.x, band_i = m.y;
    int nf = fdiv(plane0, a.d_C);                      // face / channel of the (first) plane
    int c = plane0 - nf * a.C;
    int f = nf % 6;
    int ya, yb, np;
    if (a.nb > 1) { ya = band_i * a.Rb; yb = min(ya + a.Rb, H); np = 1; }
    else { ya = 0; yb = H; np = min(a.k, a.n_planes - plane0); }
    const uint32_t* in_s = ring + s * a.slot_words + (int)(((int64_t)plane0 * HW + ya * W) & 3);
    // epilogue constants of the tile's first plane: fetched BEFORE the wait, so the global-load latency hides
    // behind the bulk copy instead of stalling the first store of every tile
    Epi ep0 = {1.0f, 0.0f, a.relu};
    if (EPI) {
      if (a.scale) ep0.sc = __ldg(a.scale + c);
      if (a.shift) ep0.sh = __ldg(a.shift + c);
    }
    tma::mbar_wait(&bar[s], phase);
#ifdef CP360_TRACE
    if (lane == 0 && first_tile) CP360_TRACE_MIN(2);
    first_tile = false;
#endif
#pragma unroll 1
    for (int j = 0; j < np; ++j) {
      const uint32_t* band = in_s + j * HW;                           // row ya of this plane
      uint32_t* __restrict__ outp = a.y + ((int64_t)nf * a.out_C + a.out_coff + c) * HoWo;
      Epi ep = ep0;
      if (EPI && j > 0) {                                  // further planes of a multi-plane tile
        if (a.scale) ep.sc = __ldg(a.scale + c);
        if (a.shift) ep.sh = __ldg(a.shift + c);
      }

      // ---- A. interior rows: shifted copy shared -> global
      row_copy<NJ, FULL, EPI>(band + lane, outp + (ya + g.pt) * Wo + g.pl + lane, yb - ya, W, Wo, lane, ep);

      // ---- B. push: halo elements of other faces' planes (same cube, same channel) whose source
      //         pixel lies in rows [ya, yb) of this plane
      int ia[4], wd[4], cnt[4];
      const uint4 pr4 = ptab[f * a.nb + band_i];
      const uint32_t prs[4] = {pr4.x, pr4.y, pr4.z, pr4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const PushEntry& pe = g.push[f][e];
        const int i0 = (int)(prs[e] & 0xffffu);
        ia[e] = i0;
        wd[e] = (int)(prs[e] >> 16);
        cnt[e] = wd[e] * (pe.drive == 0 ? pe.V : pe.U);
        if (pe.drive == 0 && pe.V >= 16 && wd[e] > 0) {
          // whole plate rows (top / down plates fed by the first / last rows of a face): a row
          // copy with an optional reversal and clamped ends, lanes along the row
          uint32_t* __restrict__ dface = a.y + ((int64_t)(nf - f + pe.dface) * a.out_C + a.out_coff + c) * HoWo;
#pragma unroll 1
          for (int u = i0; u < i0 + wd[e]; ++u) {
            const int r = pe.cmode ? min(max(u + pe.off, 0), H - 1) : u;
            const uint32_t* srow = band + (pe.yr * r + pe.y0 - ya) * W + pe.xr * r + pe.x0;
            uint32_t* __restrict__ drow = dface + (pe.oy0 + u) * Wo + pe.ox0;
#pragma unroll 1
            for (int v = lane; v < pe.V; v += 32) {
              const int cc = pe.cmode ? v : min(max(v + pe.off, 0), W - 1);
              CP360_ROW_ST(drow + v, epi_apply<EPI>(srow[pe.xc * cc], ep));
            }
          }
          cnt[e] = 0;
        }
      }
      const int c1 = cnt[0], c2 = c1 + cnt[1], c3 = c2 + cnt[2], n_push = c3 + cnt[3];
      uint32_t* __restrict__ cube_out = a.y + ((int64_t)(nf - f) * a.out_C + a.out_coff + c) * HoWo;
#pragma unroll 1
      for (int q = lane; q < n_push; q += 32) {
        const int e = (q >= c1) + (q >= c2) + (q >= c3);
        const int ql = q - (e == 0 ? 0 : e == 1 ? c1 : e == 2 ? c2 : c3);
        const int i0 = e == 0 ? ia[0] : e == 1 ? ia[1] : e == 2 ? ia[2] : ia[3];
        const int w = e == 0 ? wd[0] : e == 1 ? wd[1] : e == 2 ? wd[2] : wd[3];
        const PushEntry& pe = g.push[f][e];
        // ql = t2 * den + rem; all values are far below 2^22, so the float quotient is exact
        const int den = pe.drive == 0 ? pe.V : w;
        const int t2 = (int)__fdividef((float)ql + 0.5f, (float)den);
        const int rem = ql - t2 * den;
        const int u = pe.drive == 0 ? i0 + t2 : t2;
        const int v = pe.drive == 0 ? rem : i0 + rem;
        const int r = pe.cmode ? min(max(u + pe.off, 0), H - 1) : u;
        const int cc = pe.cmode ? v : min(max(v + pe.off, 0), W - 1);
        const int yy = pe.yr * r + pe.yc * cc + pe.y0, xx = pe.xr * r + pe.xc * cc + pe.x0;
        CP360_ROW_ST(cube_out + (int64_t)pe.dface * a.out_C * HoWo + (pe.oy0 + u) * Wo + pe.ox0 + v,
                     epi_apply<EPI>(band[(yy - ya) * W + xx], ep));
      }
      if (np > 1 && ++c == a.C) { c = 0; ++nf; if (++f == 6) f = 0; }
    }
    __syncwarp();                                  // every lane is done reading slot s and its meta:
    if (lane == 0) arm(s);                         // re-arm it with the warp's next tile
    __syncwarp();
    if (++s == slots) { s = 0; phase ^= 1u; }
  }
#ifdef CP360_TRACE
  if (lane == 0) CP360_TRACE_MAX(3);
#endif
  if (a.order >= 3) {
    __syncthreads();                                   // every warp of the CTA has drawn its last unit
    if (threadIdx.x == 0) {
      if (atomicAdd(a.work + 1, 1u) == gridDim.x - 1) {   // last CTA of the launch: hand the pair back zeroed
        a.work[0] = 0;
        a.work[1] = 0;
        __threadfence();
      }
    }
  }
}

}  // namespace cp360
