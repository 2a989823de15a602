// CubePad cube-tile kernel (32-bit elements, small planes: H <= ~45) — included by cubepad.cu.
//
// A tile is ALL SIX faces of a run of channels of one cube (the unit inside which CubePad is
// closed: every halo pixel of a face is an interior pixel of another face of the same cube and
// channel, model/cube_pad.py:114-176). Six TMA bulk loads (one contiguous k*H*W chunk per face)
// stage the tile in shared memory; every input element is read from DRAM exactly once.
//
// Warp roles: warp 0 is the producer (one lane issues the bulk loads, `stages` tiles ahead, gated
// by per-stage empty barriers); the other warps are consumers. A consumer thread owns fixed
// output positions e = (face, oy, ox) of the padded cube; a per-geometry table in shared memory
// (built once per CTA from cubepad_geom.h) gives the staged source word and the destination
// offset of each position. Per tile and position the thread then walks the channels:
// LDS -> STG, consecutive lanes writing consecutive words of the padded plane (coalesced 128 B
// streaming stores; the padded output never passes through shared memory).
//
// Work partition: the (cube, channel-quantum) index space is cut into one contiguous, balanced
// range per CTA (no round-robin tail), walked in chunks of at most kmax channels that do not
// cross a cube boundary.
#pragma once
#include "common.cuh"
#include "cubepad_geom.h"
#include "tma.cuh"

namespace cp360 {

constexpr int kCubeMaxStages = 8;

struct Cube2Args {
  const uint32_t* x;
  uint32_t* y;
  int64_t n_quanta;     // N * (C / kq)
  int32_t C;
  int32_t kq;           // channel quantum (bulk-copy alignment)
  int32_t kmax;         // channels per chunk (multiple of kq)
  int32_t qpc;          // quanta per cube = C / kq
  int32_t stages;
  int32_t stage_words;  // 6 * kmax * H * W
  int32_t lut_off;      // byte offset of the position table in dynamic shared memory
  int32_t ring_off;     // byte offset of the input ring
};

struct CubeChunk {
  int64_t q, q_end;     // next quantum / end of this CTA's range
  int64_t n;            // cube index of the current chunk
  int32_t c0, kl;       // first channel / number of channels (0: range exhausted)
};

__device__ __forceinline__ void chunk_next(CubeChunk& ck, const Cube2Args& a) {
  if (ck.q >= ck.q_end) { ck.kl = 0; return; }
  ck.n = ck.q / a.qpc;
  const int cq = (int)(ck.q - ck.n * a.qpc);
  const int take = (int)min((int64_t)min(a.kmax / a.kq, a.qpc - cq), ck.q_end - ck.q);
  ck.c0 = cq * a.kq;
  ck.kl = take * a.kq;
  ck.q += take;
}

__global__ void __launch_bounds__(1024)
cubepad_cube2_kernel(const Cube2Args a, const __grid_constant__ CubePadGeom g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                 // [stages]
  uint64_t* empty = full + kCubeMaxStages;                                // [stages]
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + a.lut_off);      // [6*Ho*Wo]
  const uint32_t* ring = reinterpret_cast<const uint32_t*>(smem_raw + a.ring_off);
  const int HW = g.H * g.W, HoWo = g.Ho * g.Wo, n_pos = 6 * HoWo;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_cons = (int)blockDim.x - 32, n_cons_warps = n_cons >> 5;
  const int fstride = a.kmax * HW;                                        // face stride in a stage

  CP360_TRACE_BEGIN(2)
  pdl_trigger();
  CP360_TRACE_INIT_MIN(2);
  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      tma::mbar_init(&full[s], 1);
      tma::mbar_init(&empty[s], n_cons_warps);
    }
    tma::fence_mbar_init();
  }
  // position table: lo 16 bits = staged source word of channel 0, hi 16 = face << 13 | plane offset
  for (int e = tid; e < n_pos; e += blockDim.x) {
    const int f = e / HoWo, r = e - f * HoWo;
    const int oy = r / g.Wo, ox = r - oy * g.Wo;
    int sf;
    const int pix = cubepad_src(g, f, oy, ox, &sf);
    lut[e] = (uint32_t)(sf * fstride + pix) | ((uint32_t)(f << 13 | r) << 16);
  }
  __syncthreads();
  pdl_wait();
  CP360_TRACE_T0(1);

  CubeChunk ck;
  ck.q = (a.n_quanta * blockIdx.x) / gridDim.x;
  ck.q_end = (a.n_quanta * (blockIdx.x + 1)) / gridDim.x;
  ck.kl = 0; ck.n = 0; ck.c0 = 0;
  chunk_next(ck, a);

  if (warp == 0) {
    // ---------------- producer
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int it = 0; ck.kl; ++it) {
        if (it >= a.stages) tma::mbar_wait(&empty[s], ph ^ 1u);
        const uint32_t bytes = (uint32_t)(ck.kl * HW) * 4u;
        tma::mbar_expect_tx(&full[s], 6u * bytes);
        uint32_t* dst = const_cast<uint32_t*>(ring) + (size_t)s * a.stage_words;
        const uint32_t* src = a.x + ((ck.n * 6) * a.C + ck.c0) * HW;
#pragma unroll
        for (int f = 0; f < 6; ++f)
          tma::bulk_load(dst + f * fstride, src + (int64_t)f * a.C * HW, bytes, &full[s]);
        chunk_next(ck, a);
        if (++s == a.stages) { s = 0; ph ^= 1u; }
      }
    }
    return;
  }

  // ---------------- consumers
  const int ctid = tid - 32;
  const int64_t CHoWo = (int64_t)a.C * HoWo;
  int s = 0;
  uint32_t ph = 0;
  while (ck.kl) {
    const uint32_t* in_s = ring + (size_t)s * a.stage_words;
    uint32_t* __restrict__ out = a.y + ((ck.n * 6) * a.C + ck.c0) * HoWo;
    const int kl = ck.kl;
    tma::mbar_wait(&full[s], ph);
#ifdef CP360_TRACE
    if (lane == 0) CP360_TRACE_MIN(2);
#endif
#pragma unroll 1
    for (int e = ctid; e < n_pos; e += n_cons) {
      const uint32_t l = lut[e];
      const uint32_t* sp = in_s + (l & 0xffffu);
      uint32_t* __restrict__ dp = out + (int64_t)(l >> 29) * CHoWo + ((l >> 16) & 0x1fffu);
      int cc = 0;
#pragma unroll 1
      for (; cc + 4 <= kl; cc += 4, sp += 4 * HW, dp += 4 * HoWo) {
        const uint32_t v0 = sp[0], v1 = sp[HW], v2 = sp[2 * HW], v3 = sp[3 * HW];
        __stcs(dp, v0);
        __stcs(dp + HoWo, v1);
        __stcs(dp + 2 * HoWo, v2);
        __stcs(dp + 3 * HoWo, v3);
      }
#pragma unroll 1
      for (; cc < kl; ++cc, sp += HW, dp += HoWo) __stcs(dp, *sp);
    }
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(&empty[s]);
    chunk_next(ck, a);
    if (++s == a.stages) { s = 0; ph ^= 1u; }
  }
#ifdef CP360_TRACE
  if (lane == 0) CP360_TRACE_MAX(3);
#endif
}

}  // namespace cp360
