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
// Work partition: the chunk list (cube, channel block of kmax channels) is dealt dynamically — the
// producer draws the next chunk id from a per-launch device counter (common.cuh:
// acquire_work_counter) — so SMs that stream faster take more chunks; without a counter the chunks
// are dealt round-robin. The chunk id staged in each ring slot travels with it in shared memory.
#pragma once
#include "common.cuh"
#include "cubepad_geom.h"
#include "tma.cuh"

namespace cp360 {

constexpr int kCubeMaxStages = 8;
constexpr int kCubeMaxConsWarps = 16;                        // + the producer warp: 544 threads, up to 120 registers each
constexpr int kCubeMaxThreads = 32 * (kCubeMaxConsWarps + 1);

struct Cube2Args {
  const uint32_t* x;
  uint32_t* y;
  uint32_t* work;       // {next chunk, finished CTAs} (zero at launch) or nullptr: chunks dealt round-robin
  int64_t n_chunks;     // N * cblocks
  int32_t C;
  int32_t out_C, out_coff;   // output plane of (face-in-batch nf, channel c): nf * out_C + out_coff + c
  const float* scale;        // EPI kernels: per-channel affine (+ ReLU) fused into the copy, see cubepad_row.cuh
  const float* shift;
  int32_t relu;
  int32_t kmax;         // channels per chunk (multiple of the bulk-copy channel quantum)
  int32_t cblocks;      // chunks per cube = ceil(C / kmax)
  int32_t stages;
  int32_t stage_words;  // 6 * kmax * H * W
  int32_t lut_off;      // byte offset of the position table in dynamic shared memory
  int32_t epi_off;      // EPI kernels: byte offset of the per-stage {scale[kmax], shift[kmax]} slots
  int32_t ring_off;     // byte offset of the input ring
};

// TH, TP > 0: face width and (symmetric) pad known at compile time; TK > 0: kmax known at compile
// time — plane strides become immediates and the channel walk of a full chunk is unrolled, one LDS
// and one STG per word. TH == 0 / TK == 0: run-time geometry / chunk depth.
template <int TH, int TP, int TK, bool EPI>
__global__ void __launch_bounds__(kCubeMaxThreads, 1)
cubepad_cube2_kernel(const Cube2Args a, const __grid_constant__ CubePadGeom g) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                 // [stages]
  uint64_t* empty = full + kCubeMaxStages;                                // [stages]
  int64_t* chunk_of = reinterpret_cast<int64_t*>(empty + kCubeMaxStages); // [stages] chunk id staged there, -1: end
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem_raw + a.lut_off);      // [6*Ho*Wo]
  const uint32_t* ring = reinterpret_cast<const uint32_t*>(smem_raw + a.ring_off);
  float* epi_s = reinterpret_cast<float*>(smem_raw + a.epi_off);          // [stages][2][kmax] (EPI kernels)
  const int HW = TH ? TH * TH : g.H * g.W;
  const int HoWo = TH ? (TH + 2 * TP) * (TH + 2 * TP) : g.Ho * g.Wo;
  const int n_pos = 6 * HoWo;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_cons = (int)blockDim.x - 32, n_cons_warps = n_cons >> 5;
  const int kmax = TK ? TK : a.kmax;
  const int fstride = kmax * HW;                                          // face stride in a stage

  CP360_TRACE_BEGIN(2)
  pdl_trigger();
  CP360_TRACE_INIT_MIN(2);
  if (tid == 0) {
    for (int s = 0; s < a.stages; ++s) {
      tma::mbar_init(&full[s], EPI ? 32 : 1);          // EPI: every producer lane releases its scale/shift writes itself
      // every consumer THREAD arrives on empty[s] (CP360_ARRIVE_PER_WARP: one lane per warp after __syncwarp —
      // equally correct under the PTX memory model and equally fast on B200 (profiles/README.md §racecheck),
      // but compute-sanitizer's racecheck only credits a barrier to the threads that arrive on it, so the
      // per-thread form is the one that verifies clean)
#ifdef CP360_ARRIVE_PER_WARP
      tma::mbar_init(&empty[s], n_cons_warps);
#else
      tma::mbar_init(&empty[s], n_cons);
#endif
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

  if (warp == 0) {
    // ---------------- producer: draw chunks, stage them `stages` deep. Plain copy: one lane does everything.
    // EPI kernels: the whole warp walks the loop — lane 0 waits / draws / issues the bulk loads, all lanes copy
    // the chunk's scale[] / shift[] slice into the stage's shared-memory slot, so the consumers never touch
    // global memory for the epilogue constants.
    if (!EPI && lane != 0) return;
    int s = 0;
    uint32_t ph = 0;
    // dynamic dealing: the first chunk of a CTA is its own index (no round trip to the counter in front of
    // the first load), every further one gridDim.x + a ticket drawn one step ahead of its use
    uint32_t ticket = blockIdx.x;
    for (int64_t it = 0;; ++it) {
      int64_t q = 0;
      if (EPI && it >= a.stages) tma::mbar_wait(&empty[s], ph ^ 1u);   // every lane acquires the slot it is about to write
      if (lane == 0) {
        if (!EPI && it >= a.stages) tma::mbar_wait(&empty[s], ph ^ 1u);
        q = (int64_t)blockIdx.x + it * gridDim.x;
        if (a.work) {
          q = (int64_t)ticket;
          if (q < a.n_chunks) ticket = gridDim.x + atomicAdd(a.work, 1u);
        }
      }
      if (EPI) q = __shfl_sync(0xffffffffu, q, 0);
      if (q >= a.n_chunks) {
        if (lane == 0) chunk_of[s] = -1;
        if (EPI || lane == 0) tma::mbar_arrive(&full[s]);   // completes the phase: consumers see the end mark
        break;
      }
      const int64_t n = q / a.cblocks;
      const int c0 = (int)(q - n * a.cblocks) * kmax;
      const int kl = min(kmax, a.C - c0);
      if (EPI) {
        float* es = epi_s + (size_t)s * 2 * kmax;
        for (int i = lane; i < kl; i += 32) {
          es[i] = a.scale ? __ldg(a.scale + c0 + i) : 1.0f;
          es[kmax + i] = a.shift ? __ldg(a.shift + c0 + i) : 0.0f;
        }
        if (lane != 0) tma::mbar_arrive(&full[s]);     // release of this lane's slot writes (lane 0: expect_tx below)
      }
      if (lane == 0) {
        chunk_of[s] = q;
        const uint32_t bytes = (uint32_t)(kl * HW) * 4u;
        tma::mbar_expect_tx(&full[s], 6u * bytes);
        uint32_t* dst = const_cast<uint32_t*>(ring) + (size_t)s * a.stage_words;
        const uint32_t* src = a.x + ((n * 6) * a.C + c0) * HW;
#pragma unroll
        for (int f = 0; f < 6; ++f)
          tma::bulk_load(dst + f * fstride, src + (int64_t)f * a.C * HW, bytes, &full[s]);
      }
      if (++s == a.stages) { s = 0; ph ^= 1u; }
    }
    if (lane == 0 && a.work && atomicAdd(a.work + 1, 1u) == gridDim.x - 1) {   // last CTA: hand the pair back zeroed
      a.work[0] = 0;
      a.work[1] = 0;
      __threadfence();
    }
    return;
  }

  // ---------------- consumers
  const int ctid = tid - 32;
  const int64_t CHoWo = (int64_t)a.out_C * HoWo;
  int s = 0;
  uint32_t ph = 0;
  while (true) {
    tma::mbar_wait(&full[s], ph);
    const int64_t q = chunk_of[s];
    if (q < 0) break;
#ifdef CP360_TRACE
    if (lane == 0) CP360_TRACE_MIN(2);
#endif
    const int64_t n = q / a.cblocks;
    const int c0 = (int)(q - n * a.cblocks) * kmax;
    const int kl = min(kmax, a.C - c0);
    const uint32_t* in_s = ring + (size_t)s * a.stage_words;
    uint32_t* __restrict__ out = a.y + ((n * 6) * a.out_C + a.out_coff + c0) * HoWo;
    const float* es = epi_s + (size_t)s * 2 * kmax;                       // this stage's scale[] / shift[] (EPI)
    auto epi = [&](uint32_t v, int cc) -> uint32_t {
      Epi ep = {es[cc], es[kmax + cc], a.relu};
      return epi_apply<EPI>(v, ep);
    };
    if (EPI && TK > 0 && kl == TK) {
      // full chunk with an epilogue: a batch of channels' constants lives in registers while the thread walks
      // all of its positions, so the epilogue costs three ALU instructions per word and no extra loads
      constexpr int KB = TK >= 8 ? 8 : (TK > 0 ? TK : 1);
#pragma unroll 1
      for (int c8 = 0; c8 < TK; c8 += KB) {
        float sc[KB], sh[KB];
#pragma unroll
        for (int j = 0; j < KB; ++j) { sc[j] = es[c8 + j]; sh[j] = es[kmax + c8 + j]; }
#pragma unroll 1
        for (int e = ctid; e < n_pos; e += n_cons) {
          const uint32_t l = lut[e];
          const uint32_t* sp = in_s + (l & 0xffffu) + c8 * HW;
          uint32_t* __restrict__ dp = out + (int64_t)(l >> 29) * CHoWo + ((l >> 16) & 0x1fffu) + c8 * HoWo;
          uint32_t v[KB];
#pragma unroll
          for (int j = 0; j < KB; ++j) v[j] = sp[j * HW];
#pragma unroll
          for (int j = 0; j < KB; ++j) {
            const Epi ep = {sc[j], sh[j], a.relu};
            __stcs(dp + j * HoWo, epi_apply<true>(v[j], ep));
          }
        }
      }
    } else if (TK > 0 && kl == TK) {
      // full chunk: channel walk unrolled in batches of 8, strides are immediates when TH > 0
      constexpr int KB = TK >= 8 ? 8 : (TK > 0 ? TK : 1);
#pragma unroll 1
      for (int e = ctid; e < n_pos; e += n_cons) {
        const uint32_t l = lut[e];
        const uint32_t* sp = in_s + (l & 0xffffu);
        uint32_t* __restrict__ dp = out + (int64_t)(l >> 29) * CHoWo + ((l >> 16) & 0x1fffu);
#pragma unroll
        for (int c8 = 0; c8 < TK; c8 += KB) {
          uint32_t v[KB];
#pragma unroll
          for (int j = 0; j < KB; ++j) v[j] = sp[(c8 + j) * HW];
#pragma unroll
          for (int j = 0; j < KB; ++j) __stcs(dp + (c8 + j) * HoWo, EPI ? epi(v[j], c8 + j) : v[j]);
        }
      }
    } else {
#pragma unroll 1
      for (int e = ctid; e < n_pos; e += n_cons) {
        const uint32_t l = lut[e];
        const uint32_t* sp = in_s + (l & 0xffffu);
        uint32_t* __restrict__ dp = out + (int64_t)(l >> 29) * CHoWo + ((l >> 16) & 0x1fffu);
        int cc = 0;
#pragma unroll 1
        for (; cc + 4 <= kl; cc += 4, sp += 4 * HW, dp += 4 * HoWo) {
          const uint32_t v0 = sp[0], v1 = sp[HW], v2 = sp[2 * HW], v3 = sp[3 * HW];
          __stcs(dp, EPI ? epi(v0, cc) : v0);
          __stcs(dp + HoWo, EPI ? epi(v1, cc + 1) : v1);
          __stcs(dp + 2 * HoWo, EPI ? epi(v2, cc + 2) : v2);
          __stcs(dp + 3 * HoWo, EPI ? epi(v3, cc + 3) : v3);
        }
#pragma unroll 1
        for (; cc < kl; ++cc, sp += HW, dp += HoWo) __stcs(dp, EPI ? epi(*sp, cc) : *sp);
      }
    }
#ifdef CP360_ARRIVE_PER_WARP
    __syncwarp();                                      // orders every lane's reads of the stage before lane 0's release
    if (lane == 0) tma::mbar_arrive(&empty[s]);
#else
    tma::mbar_arrive(&empty[s]);                       // release: this thread's reads of the stage are done
#endif
    if (++s == a.stages) { s = 0; ph ^= 1u; }
  }
#ifdef CP360_TRACE
  if (lane == 0) CP360_TRACE_MAX(3);
#endif
}

}  // namespace cp360
