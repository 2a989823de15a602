// Cube faces -> equirectangular for sm_100a — replaces Cube2Equi.to_equi_nn
// (utils/cube_to_equi.py:37-66: per call 2 map uploads, 6 full-grid grid_sample passes and 6
// boolean-mask scatters) with ONE pass over a per-resolution sampling plan built on the host
// (cp360_c2e_build_plan): per output pixel the face id, the north-west tap and the four fp32
// bilinear weights exactly as torch's grid_sample forms them. Out-of-face taps contribute 0
// (padding_mode='zeros'), which is what blends the output toward 0 along the cube seams.
//
//   c2e_kernel       equi[B,C,2w,4w]; thread = output pixel, loops over a channel chunk
//   c2e_max_kernel   sal[B,2w,4w] = max_c equi — the form every call site consumes
//                    (test_temporal.py:82-84); per-thread running max over the chunk, warp-free
//                    combine across chunks with an order-preserving integer atomic max
//   c2e_small_kernel w <= 16: the 6 faces of a channel group are staged in shared memory by TMA
//                    bulk copies; used by both variants
//   c2e_bwd_kernel   bilinear scatter-add of the output gradient (training path)
//   c2e_cubic_*      Cube2Equi.to_equi_cv2 (cube_to_equi.py:68-91): cv2.remap(INTER_CUBIC) arithmetic
#include <algorithm>
#include <stdlib.h>

#include <cooperative_groups.h>

#include "common.cuh"
#include "tma.cuh"

namespace cg = cooperative_groups;

namespace cp360 {

struct Tap {
  int face, y0, x0;
};

__device__ __forceinline__ Tap decode_tap(uint32_t t) {
  Tap r;
  r.face = (int)(t >> 28);
  r.y0 = (int)((t >> 14) & 0x3fffu) - 1;
  r.x0 = (int)(t & 0x3fffu) - 1;
  return r;
}

// float max through integer atomics, with torch.max's NaN semantics: a NaN beats every number. As an int the
// canonical quiet NaN 0x7fc00000 is above +inf, so atomicMax keeps it against every non-negative value, and as
// an unsigned word it is below every negative float's pattern, so the atomicMin of the negative branch keeps it too.
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v != v) { atomicMax(reinterpret_cast<int*>(addr), 0x7fc00000); return; }
  v += 0.0f;   // -0.0 -> +0.0
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// (value, channel) keys for the arg-max variant: order-preserving map of the float into the high word,
// ~channel in the low word, so a 64-bit atomicMax keeps the largest value and, among equal values,
// the LOWEST channel (torch.max(dim) returns the first maximal index). NaN is canonicalised to the quiet NaN
// above +inf, so the first NaN channel wins, as in torch. Every key is > 0.
__device__ __forceinline__ unsigned long long argmax_key(float v, int c) {
  const uint32_t low = 0xffffffffu - (uint32_t)c;
  if (v != v) return (0xffffffffull << 32) | low;            // above every number (no float arithmetic on the NaN)
  uint32_t b = __float_as_uint(v);
  if (b == 0x80000000u) b = 0u;                                // -0.0 -> +0.0
  const uint32_t o = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)o << 32) | low;
}

// running (max, arg-max) update with torch.max(dim) semantics: strictly greater replaces (first maximal index
// wins), a NaN replaces any number and is never replaced (first NaN wins)
__device__ __forceinline__ void max_update(float acc, int c, float& best, int& best_c) {
  if ((acc > best || acc != acc) && best == best) { best = acc; best_c = c; }
}
__device__ __forceinline__ float max_nan(float a, float b) {      // NaN-propagating max
  return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);
}
// is (av, ac) a better (max, index) pair than (bv, bc)?
__device__ __forceinline__ bool pair_better(float av, int ac, float bv, int bc) {
  const bool an = av != av, bn = bv != bv;
  if (an || bn) return an && (!bn || ac < bc);
  return av > bv || (av == bv && ac < bc);
}

__global__ void c2e_argmax_decode_kernel(const unsigned long long* __restrict__ keys, float* __restrict__ sal,
                                         int32_t* __restrict__ arg, int64_t n) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[i];
    const uint32_t o = (uint32_t)(k >> 32);
    sal[i] = __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
    arg[i] = (int32_t)(0xffffffffu - (uint32_t)k);
  }
}

__global__ void fill_kernel(float* __restrict__ p, int64_t n, float v) {
  pdl_trigger();
  pdl_wait();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}

constexpr int kC2eThreads = 256;

// MODE 0: write equi[B,C,P]; MODE 1: channel max into sal[B,P]; MODE 2: (max, arg-max channel) keys
// into a uint64 [B,P] scratch (decoded by c2e_argmax_decode_kernel)
template <int MODE>
__global__ void __launch_bounds__(kC2eThreads)
c2e_kernel(const float* __restrict__ cube, const uint32_t* __restrict__ taps,
           const float4* __restrict__ wts, float* __restrict__ out, int64_t B, int C, int w,
           int ch_per_block) {
  pdl_trigger();
  pdl_wait();
  const int P = 8 * w * w, ww = w * w;
  const int pix = blockIdx.x * kC2eThreads + threadIdx.x;
  if (pix >= P) return;
  const Tap t = decode_tap(__ldg(taps + pix));
  const float4 wt = __ldg(wts + pix);
  const bool xw_ok = (unsigned)t.x0 < (unsigned)w, xe_ok = (unsigned)(t.x0 + 1) < (unsigned)w;
  const bool yn_ok = (unsigned)t.y0 < (unsigned)w, ys_ok = (unsigned)(t.y0 + 1) < (unsigned)w;
  const bool nw_ok = xw_ok && yn_ok, ne_ok = xe_ok && yn_ok, sw_ok = xw_ok && ys_ok,
             se_ok = xe_ok && ys_ok;
  const int o_nw = t.y0 * w + t.x0;
  const int c_begin = blockIdx.y * ch_per_block, c_end = min(C, c_begin + ch_per_block);
  for (int64_t b = blockIdx.z; b < B; b += gridDim.z) {
    const float* src = cube + ((b * 6 + t.face) * C + c_begin) * (int64_t)ww + o_nw;
    float best = -INFINITY;
    int best_c = c_begin;
#pragma unroll 4
    for (int c = c_begin; c < c_end; ++c, src += ww) {
      float acc = 0.0f;                      // order of torch's grid_sampler CUDA kernel
      if (nw_ok) acc = fmaf(__ldg(src), wt.x, acc);
      if (ne_ok) acc = fmaf(__ldg(src + 1), wt.y, acc);
      if (sw_ok) acc = fmaf(__ldg(src + w), wt.z, acc);
      if (se_ok) acc = fmaf(__ldg(src + w + 1), wt.w, acc);
      if (MODE == 0) __stcs(out + (b * C + c) * (int64_t)P + pix, acc);
      else if (MODE == 1) best = max_nan(best, acc);
      else max_update(acc, c, best, best_c);
    }
    if (MODE == 1) atomic_max_float(out + b * (int64_t)P + pix, best);
    if (MODE == 2) atomicMax(reinterpret_cast<unsigned long long*>(out) + b * (int64_t)P + pix, argmax_key(best, best_c));
  }
}

// Small faces (w <= 16): stage the whole 6-face cube of a channel group in shared memory.
// Block = (b, channel group). Threads own output pixels (P = 8w^2 <= 2048, looped), channels
// are the inner loop, so the plan entry lives in registers and the cube is read from DRAM
// exactly once through six bulk copies.
constexpr int kC2eSmallThreads = 512;
// Bank skew of the staged faces: a warp covers 32 consecutive pixels of an equirect row, i.e. runs of
// ~w pixels on up to four faces (B L F R around the equator) that sample the SAME face row; with
// every face at a multiple of 32 words those runs collide in the same banks (4-way conflicts at
// w = 8). Offsetting the equatorial faces by 0 / 8 / 16 / 24 words (16 B multiples, as the bulk
// copies need) spreads them over all 32 banks.
__device__ __constant__ int kFaceSkew[6] = {0 /*B*/, 4 /*D*/, 8 /*F*/, 16 /*L*/, 24 /*R*/, 28 /*T*/};   // non-decreasing: regions stay disjoint
constexpr int kFaceSkewMax = 32;

// TWW > 0: w*w known at compile time (49 / 64): plane strides of the unrolled channel walk become immediates
template <int MODE, int TWW>
__global__ void __launch_bounds__(kC2eSmallThreads)
c2e_small_kernel(const float* __restrict__ cube, const uint32_t* __restrict__ taps,
                 const float4* __restrict__ wts, float* __restrict__ out, int C, int w, int kch,
                 int groups) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  float* cs = reinterpret_cast<float*>(smem_raw + 128);       // [6][kch][w*w]
  const int ww = TWW ? TWW : w * w, P = 8 * ww;
  const int b = blockIdx.x / groups, gidx = blockIdx.x - b * groups;
  const int c0 = gidx * kch, kl = min(kch, C - c0);
  const int tid = threadIdx.x;
  CP360_TRACE_BEGIN(4)
  pdl_trigger();
  pdl_wait();
  CP360_TRACE_T0(1);
  if (tid == 0) {
    tma::mbar_init(bar, 1);
    tma::fence_mbar_init();
    const uint32_t bytes = (uint32_t)(kl * ww) * 4u;
    tma::mbar_expect_tx(bar, 6u * bytes);
#pragma unroll
    for (int f = 0; f < 6; ++f)
      tma::bulk_load(cs + (size_t)f * kch * ww + kFaceSkew[f], cube + (((int64_t)b * 6 + f) * C + c0) * ww, bytes, bar);
  }
  __syncthreads();
  tma::mbar_wait(bar, 0);
  CP360_TRACE_T0(2);
  for (int pix = tid; pix < P; pix += kC2eSmallThreads) {
    const Tap t = decode_tap(__ldg(taps + pix));
    const float4 wt = __ldg(wts + pix);
    const bool xw_ok = (unsigned)t.x0 < (unsigned)w, xe_ok = (unsigned)(t.x0 + 1) < (unsigned)w;
    const bool yn_ok = (unsigned)t.y0 < (unsigned)w, ys_ok = (unsigned)(t.y0 + 1) < (unsigned)w;
    // clamp tap addresses into the face and zero the weight instead of branching per channel
    const int xw = xw_ok ? t.x0 : 0, xe = xe_ok ? t.x0 + 1 : 0, yn = yn_ok ? t.y0 : 0,
              ys = ys_ok ? t.y0 + 1 : 0;
    const bool nw_ok = xw_ok && yn_ok, ne_ok = xe_ok && yn_ok, sw_ok = xw_ok && ys_ok,
               se_ok = xe_ok && ys_ok;
    const float* src = cs + (size_t)t.face * kch * ww + kFaceSkew[t.face];
    const float* p_nw = src + yn * w + xw;
    const float* p_ne = src + yn * w + xe;
    const float* p_sw = src + ys * w + xw;
    const float* p_se = src + ys * w + xe;
    float best = -INFINITY;
    int best_c = 0;
    float* dst = out + ((int64_t)b * C + c0) * P + pix;
#pragma unroll 8
    for (int c = 0; c < kl; ++c) {
      float acc = 0.0f;
      if (nw_ok) acc = fmaf(p_nw[c * ww], wt.x, acc);
      if (ne_ok) acc = fmaf(p_ne[c * ww], wt.y, acc);
      if (sw_ok) acc = fmaf(p_sw[c * ww], wt.z, acc);
      if (se_ok) acc = fmaf(p_se[c * ww], wt.w, acc);
      if (MODE == 0) __stcs(dst + c * P, acc);
      else if (MODE == 1) best = max_nan(best, acc);
      else max_update(acc, c, best, best_c);
    }
    if (MODE == 1) atomic_max_float(out + (int64_t)b * P + pix, best);
    if (MODE == 2) atomicMax(reinterpret_cast<unsigned long long*>(out) + (int64_t)b * P + pix, argmax_key(best, c0 + best_c));
  }
  CP360_TRACE_T0(3);
}

// ---- K3m, small faces (w <= 16): channel max WITHOUT atomics, fill pass or scratch ----------------------------
// A cluster of G CTAs owns one frame; CTA r streams channels [r*Cg, (r+1)*Cg) of all six faces through a ring of
// TMA bulk-loaded stages (K channels each) while its threads — one per output pixel (up to kC2eMaxPix per thread
// at w = 16) — keep the running max (and arg-max channel) in registers. The G partial maps meet in distributed
// shared memory: after a cluster barrier CTA r combines its slice of the pixels, G lanes per pixel each reading one
// CTA's partial through DSMEM and a warp-shuffle max (north_star (c)) over those lanes; one lane stores the result.
// Deterministic, NaN-propagating like torch.max (test_temporal.py:83), and the output is written exactly once.
constexpr int kC2eMaxPix = 4;          // pixels per thread: 8 w^2 / 512 at w = 16
constexpr int kC2eMaxStages = 4;

struct C2eMaxArgs {
  const float* cube;
  const uint32_t* taps;
  const float4* wts;
  float* sal;
  int32_t* arg;          // MODE 2 only
  int C, w;
  int Cg;                // channels per CTA (multiple of the bulk-copy quantum)
  int K;                 // channels per stage
  int stages;
  int ring_off;          // byte offset of the ring in dynamic shared memory
  int stage_floats;      // 6 * K * w * w + skew
};

// TWW > 0: w*w known at compile time (49 / 64, the reference's map sizes): the channel walk's plane offsets become
// immediates. The kernel is instruction-issue bound (ncu: 62 % issue-active, 38 thread instructions per
// pixel-channel before this), so the inner loop carries nothing it does not need: four pointers and four
// predicates set up per stage, fmaxf / a compare for the running max, and the NaN rule of torch.max kept in a
// separate "first NaN channel" register that is merged once at the end.
template <int MODE, int NPIX, int TWW>
__global__ void __launch_bounds__(kC2eSmallThreads)
c2e_max_cluster_kernel(const C2eMaxArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int G = (int)cluster.num_blocks(), r = (int)cluster.block_rank();
  const int w = a.w, ww = TWW ? TWW : w * w, P = 8 * ww;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                    // [stages]
  float* part_val = reinterpret_cast<float*>(smem_raw + 64);                 // [P]
  int* part_arg = reinterpret_cast<int*>(part_val + P);                      // [P] (MODE 2)
  float* ring = reinterpret_cast<float*>(smem_raw + a.ring_off);
  const int b = blockIdx.x / G;
  const int c_begin = min(r * a.Cg, a.C), c_end = min(c_begin + a.Cg, a.C);
  const int n_st = (c_end - c_begin + a.K - 1) / a.K;
  const int tid = threadIdx.x;
  const int fstride = a.K * ww;                                             // face stride inside a stage

  pdl_trigger();
  auto issue = [&](int i, int s) {                                           // thread 0: step i of this CTA into ring slot s = i % stages
    const int c0 = c_begin + i * a.K, kl = min(a.K, c_end - c0);
    const uint32_t bytes = (uint32_t)(kl * ww) * 4u;
    tma::mbar_expect_tx(&full[s], 6u * bytes);
    float* dst = ring + (size_t)s * a.stage_floats;
#pragma unroll
    for (int f = 0; f < 6; ++f)
      tma::bulk_load(dst + (size_t)f * fstride + kFaceSkew[f], a.cube + (((int64_t)b * 6 + f) * a.C + c0) * ww, bytes, &full[s]);
  };
  if (tid < 32) {                                        // warp-uniform branch, predicated instruction: warp 0 never diverges here
    tma::mbar_init_if(tid < a.stages, &full[tid < a.stages ? tid : 0], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  if (tid == 0)
    for (int i = 0; i < min(a.stages, n_st); ++i) issue(i, i);
  __syncwarp();

  // this thread's pixels: plan entries live in registers for the whole kernel
  int off[NPIX][4];
  float wt[NPIX][4];
  int foff[NPIX];
  float best[NPIX];
  int best_c[NPIX], nan_c[NPIX];                           // nan_c: first channel whose value is NaN (INT_MAX: none)
#pragma unroll
  for (int k = 0; k < NPIX; ++k) {
    const int pix = tid + k * kC2eSmallThreads;
    best[k] = -INFINITY;
    best_c[k] = c_begin;
    nan_c[k] = 0x7fffffff;
    foff[k] = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { off[k][j] = 0; wt[k][j] = 0.0f; }
    if (pix < P) {
      const Tap t = decode_tap(__ldg(a.taps + pix));
      const float4 w4 = __ldg(a.wts + pix);
      const bool xw_ok = (unsigned)t.x0 < (unsigned)w, xe_ok = (unsigned)(t.x0 + 1) < (unsigned)w;
      const bool yn_ok = (unsigned)t.y0 < (unsigned)w, ys_ok = (unsigned)(t.y0 + 1) < (unsigned)w;
      // out-of-face taps (padding_mode='zeros') are skipped, not multiplied by zero: an Inf / NaN next to the face
      // edge must not leak in. Their addresses are clamped into the face so the pointer set-up stays uniform.
      const int xw = xw_ok ? t.x0 : 0, xe = xe_ok ? t.x0 + 1 : 0, yn = yn_ok ? t.y0 : 0, ys = ys_ok ? t.y0 + 1 : 0;
      off[k][0] = yn * w + xw; off[k][1] = yn * w + xe; off[k][2] = ys * w + xw; off[k][3] = ys * w + xe;
      wt[k][0] = w4.x; wt[k][1] = w4.y; wt[k][2] = w4.z; wt[k][3] = w4.w;
      const unsigned ok = (unsigned)(xw_ok && yn_ok) | (unsigned)(xe_ok && yn_ok) << 1 | (unsigned)(xw_ok && ys_ok) << 2 |
                          (unsigned)(xe_ok && ys_ok) << 3;
      foff[k] = (t.face * fstride + kFaceSkew[t.face]) | (int)(ok << 24);
    }
  }

  int s = 0;                                             // ring slot and phase of step i (no division in the loop)
  uint32_t ph = 0;
  for (int i = 0; i < n_st; ++i) {
    tma::mbar_wait(&full[s], ph);
    const int c0 = c_begin + i * a.K, kl = min(a.K, c_end - c0);
    const float* st = ring + (size_t)s * a.stage_floats;
#pragma unroll
    for (int k = 0; k < NPIX; ++k) {
      if (NPIX > 1 && tid + k * kC2eSmallThreads >= P) break;
      const unsigned ok = (unsigned)foff[k] >> 24;
      const bool ok0 = ok & 1u, ok1 = ok & 2u, ok2 = ok & 4u, ok3 = ok & 8u;
      const float* src = st + (foff[k] & 0xffffff);
      const float* p0 = src + off[k][0];
      const float* p1 = src + off[k][1];
      const float* p2 = src + off[k][2];
      const float* p3 = src + off[k][3];
      const float w0 = wt[k][0], w1 = wt[k][1], w2 = wt[k][2], w3 = wt[k][3];
      float bv = best[k];
      int bc = best_c[k], nc = nan_c[k];
#pragma unroll 8
      for (int c = 0; c < kl; ++c) {
        float acc = 0.0f;                      // order of torch's grid_sampler CUDA kernel
        if (ok0) acc = fmaf(p0[c * ww], w0, acc);
        if (ok1) acc = fmaf(p1[c * ww], w1, acc);
        if (ok2) acc = fmaf(p2[c * ww], w2, acc);
        if (ok3) acc = fmaf(p3[c * ww], w3, acc);
        if (acc != acc) nc = min(nc, c0 + c);
        if (MODE == 1) bv = fmaxf(bv, acc);    // ignores a NaN operand: NaNs are carried by nc
        else if (acc > bv) { bv = acc; bc = c0 + c; }
      }
      best[k] = bv;
      best_c[k] = bc;
      nan_c[k] = nc;
    }
    __syncwarp();                                        // reconverge lane 0 (bulk-load issue) before the block barrier
    __syncthreads();                                     // every thread is done with stage s
    if (tid == 0 && i + a.stages < n_st) issue(i + a.stages, s);
    if (++s == a.stages) { s = 0; ph ^= 1u; }
  }
#pragma unroll
  for (int k = 0; k < NPIX; ++k) {
    const int pix = tid + k * kC2eSmallThreads;
    if (pix < P) {
      const bool has_nan = nan_c[k] != 0x7fffffff;       // torch.max: a NaN wins, the first NaN channel is the arg-max
      part_val[pix] = has_nan ? __int_as_float(0x7fc00000) : best[k];
      if (MODE == 2) part_arg[pix] = has_nan ? nan_c[k] : best_c[k];
    }
  }
  cluster.sync();                                        // all G partial maps are in place (release / acquire)

  // combine: CTA r owns pixels [r*Pr, (r+1)*Pr); G consecutive lanes serve one pixel, lane j reads CTA j's partial
  const int Pr = (P + G - 1) / G;
  for (int q0 = 0; q0 < Pr * G; q0 += kC2eSmallThreads) {       // warp-uniform trip count: every lane shuffles
    const int q = q0 + tid;
    const int j = q % G, pix = r * Pr + q / G;
    const bool live = q < Pr * G && pix < P;
    float v = -INFINITY;
    int vc = 0x7fffffff;
    if (live) {
      v = *cluster.map_shared_rank(part_val + pix, j);
      if (MODE == 2) vc = *cluster.map_shared_rank(part_arg + pix, j);
    }
    for (int d = G >> 1; d > 0; d >>= 1) {               // G is a power of two <= 8: the group never straddles a warp
      const float ov = __shfl_xor_sync(0xffffffffu, v, d);
      if (MODE == 2) {
        const int oc = __shfl_xor_sync(0xffffffffu, vc, d);
        if (pair_better(ov, oc, v, vc)) { v = ov; vc = oc; }
      } else {
        v = max_nan(v, ov);
      }
    }
    if (j == 0 && live) {
      a.sal[(int64_t)b * P + pix] = v;
      if (MODE == 2) a.arg[(int64_t)b * P + pix] = vc;
    }
  }
  cluster.sync();                                        // nobody exits while a neighbour may still read its partials
}

// ---- backward of c2e as a GATHER over the transposed plan (cp360_c2e_build_bwd_plan): no atomics, fixed
// summation order -> bit-reproducible gradients. A thread owns cube pixels and walks its contributors once,
// accumulating KCH channels in registers.
//   c2e_bwd_small_kernel  w <= 16: CTA r of a frame walks the channel groups r, r + G, ...; the transposed plan is
//                         copied to shared memory once per CTA (16-bit offsets / pixel ids), the gradient planes of
//                         a group arrive by one TMA bulk copy each into a two-stage ring.
//                         Contributor counts are very uneven (a pole pixel of the top face collects 46 equi pixels
//                         at w = 8, 90 at w = 16; the mean is 5), so a cube pixel with more than kC2eBwdSplit
//                         contributors is split over a team of 2..32 adjacent lanes whose partial sums meet in a
//                         fixed shuffle tree: every warp walks at most ~kC2eBwdSplit contributors per pass and the
//                         summation order stays fixed. The slot table (one word per lane-slot: pixel, position in
//                         the team, log2 team size; teams of one size are contiguous and every size class starts on
//                         a warp boundary) is built per CTA by a block scan.
//   c2e_bwd_gather_kernel any w: contributors and gradients read through the read-only path.
constexpr int kC2eBwdSplit = 8;
constexpr int kC2eBwdClasses = 6;                        // team sizes 1, 2, 4, 8, 16, 32

struct C2eBwdArgs {
  const float* gequi;
  const int32_t* offs;
  const int32_t* pix;
  const float* wts;
  float* gcube;
  int C, w, G, n_entries, slot_cap, split, order;
  int slot_off, pix_off, wts_off, ring_off;    // byte offsets in dynamic shared memory (offs at 256)
};

__device__ __forceinline__ int c2e_bwd_class(int cnt, int split) {
  int ls = 0;
  while (ls < kC2eBwdClasses - 1 && ((cnt + (1 << ls) - 1) >> ls) > split) ++ls;
  return ls;
}

// walk order of the cube pixels when slots are handed out: 1 = row by row across the four equatorial faces (0, 2, 3, 4
// in this cube layout), so that the 32 lanes of a warp read 32 different equi columns (= shared-memory banks at w = 8)
__device__ __forceinline__ int c2e_bwd_walk(int q, int w, int ww, int order) {
  if (order == 0 || q >= 5 * ww) return q;
  if (q >= 4 * ww) return q - 3 * ww;                    // face 1 after the equatorial ring
  const int row = q / (4 * w), rem = q - row * 4 * w, fi = rem / w, col = rem - fi * w;
  return (fi == 0 ? 0 : fi + 1) * ww + row * w + col;
}

template <int KCH, int TW>                               // TW: compile-time face width (0 = a.w) -> immediate plane strides
__global__ void __launch_bounds__(kC2eSmallThreads)
c2e_bwd_small_kernel(const C2eBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);                  // [2]
  int* n_slots_s = reinterpret_cast<int*>(smem_raw + 16);                  // padded slot count
  unsigned long long* scan_lo = reinterpret_cast<unsigned long long*>(smem_raw + 64);   // [16] warp totals, classes 0..3
  uint32_t* scan_hi = reinterpret_cast<uint32_t*>(smem_raw + 192);         // [16] classes 4, 5
  uint16_t* offs_s = reinterpret_cast<uint16_t*>(smem_raw + 256);          // [NC + 1]
  uint32_t* slot_s = reinterpret_cast<uint32_t*>(smem_raw + a.slot_off);   // [slot_cap]
  uint16_t* pix_s = reinterpret_cast<uint16_t*>(smem_raw + a.pix_off);     // [n_entries]
  float* wts_s = reinterpret_cast<float*>(smem_raw + a.wts_off);           // [n_entries]
  float* ring = reinterpret_cast<float*>(smem_raw + a.ring_off);           // [2][KCH][P]
  const int w = TW ? TW : a.w, ww = w * w, P = 8 * ww, NC = 6 * ww;
  const int b = blockIdx.x / a.G, r = blockIdx.x - b * a.G;
  const int groups = (a.C + KCH - 1) / KCH;
  const int n_mine = r < groups ? (groups - r + a.G - 1) / a.G : 0;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int kWarps = kC2eSmallThreads / 32;
  pdl_trigger();
  for (int i = tid; i <= NC; i += kC2eSmallThreads) offs_s[i] = (uint16_t)__ldg(a.offs + i);
  for (int i = tid; i < a.n_entries; i += kC2eSmallThreads) {
    pix_s[i] = (uint16_t)__ldg(a.pix + i);
    wts_s[i] = __ldg(a.wts + i);
  }
  for (int i = tid; i < a.slot_cap; i += kC2eSmallThreads) slot_s[i] = 0xffffffffu;
  auto issue = [&](int i) {                                                 // thread 0: the CTA's i-th group
    const int c0 = (r + i * a.G) * KCH, kl = min(KCH, a.C - c0);
    const uint32_t bytes = (uint32_t)(kl * P) * 4u;
    tma::mbar_expect_tx(&full[i & 1], bytes);
    tma::bulk_load(ring + (size_t)(i & 1) * KCH * P, a.gequi + ((int64_t)b * a.C + c0) * P, bytes, &full[i & 1]);
  };
  if (tid < 32) {                                        // warp-uniform branch, predicated body
    if (tid < 2) tma::mbar_init(&full[tid], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  pdl_wait();
  if (tid == 0)
    for (int i = 0; i < min(2, n_mine); ++i) issue(i);
  __syncwarp();

  // ---- slot table: thread t owns the cube pixels [t * cpt, (t + 1) * cpt); an exclusive block scan of the per-class
  // pixel counts (16-bit fields: NC <= 1536) gives every pixel its rank inside its class in natural order
  const int cpt = (NC + kC2eSmallThreads - 1) / kC2eSmallThreads;
  unsigned long long lo = 0;
  uint32_t hi = 0;
  for (int i = 0; i < cpt; ++i) {
    const int q = tid * cpt + i;
    if (q < NC) {
      const int cell = c2e_bwd_walk(q, w, ww, a.order);
      const int ls = c2e_bwd_class((int)offs_s[cell + 1] - (int)offs_s[cell], a.split);
      if (ls < 4) lo += 1ull << (16 * ls); else hi += 1u << (16 * (ls - 4));
    }
  }
  unsigned long long inc_lo = lo;
  uint32_t inc_hi = hi;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long t_lo = __shfl_up_sync(0xffffffffu, inc_lo, d);
    const uint32_t t_hi = __shfl_up_sync(0xffffffffu, inc_hi, d);
    if (lane >= d) { inc_lo += t_lo; inc_hi += t_hi; }
  }
  if (lane == 31) { scan_lo[wid] = inc_lo; scan_hi[wid] = inc_hi; }
  __syncthreads();
  unsigned long long pre_lo = inc_lo - lo, tot_lo = 0;
  uint32_t pre_hi = inc_hi - hi, tot_hi = 0;
  for (int k = 0; k < kWarps; ++k) {
    if (k < wid) { pre_lo += scan_lo[k]; pre_hi += scan_hi[k]; }
    tot_lo += scan_lo[k]; tot_hi += scan_hi[k];
  }
  int base[kC2eBwdClasses + 1];
  base[0] = 0;
#pragma unroll
  for (int ls = 0; ls < kC2eBwdClasses; ++ls) {
    const int n = ls < 4 ? (int)((tot_lo >> (16 * ls)) & 0xffffu) : (int)((tot_hi >> (16 * (ls - 4))) & 0xffffu);
    base[ls + 1] = base[ls] + (((n << ls) + 31) & ~31);
  }
  if (tid == 0) *n_slots_s = base[kC2eBwdClasses];
  for (int i = 0; i < cpt; ++i) {
    const int q = tid * cpt + i;
    if (q < NC) {
      const int cell = c2e_bwd_walk(q, w, ww, a.order);
      const int ls = c2e_bwd_class((int)offs_s[cell + 1] - (int)offs_s[cell], a.split);
      const int rank = ls < 4 ? (int)((pre_lo >> (16 * ls)) & 0xffffu) : (int)((pre_hi >> (16 * (ls - 4))) & 0xffffu);
      if (ls < 4) pre_lo += 1ull << (16 * ls); else pre_hi += 1u << (16 * (ls - 4));
      int bs = 0;
#pragma unroll
      for (int k = 0; k < kC2eBwdClasses; ++k) if (k == ls) bs = base[k];
      const int s0 = bs + (rank << ls);
      for (int j = 0; j < (1 << ls); ++j) slot_s[s0 + j] = (uint32_t)cell | ((uint32_t)j << 11) | ((uint32_t)ls << 16);
    }
  }
  __syncthreads();
  const int n_slots = *n_slots_s;

  for (int i = 0; i < n_mine; ++i) {
    tma::mbar_wait(&full[i & 1], (uint32_t)((i >> 1) & 1));
    const int c0 = (r + i * a.G) * KCH, kl = min(KCH, a.C - c0);
    const float* gs = ring + (size_t)(i & 1) * KCH * P;
    for (int s0 = wid * 32; s0 < n_slots; s0 += kC2eSmallThreads) {
      const uint32_t info = slot_s[s0 + lane];
      const bool valid = info != 0xffffffffu;
      const int ls = (int)(__shfl_sync(0xffffffffu, info, 0) >> 16);       // a chunk holds one class; its lane 0 is never padding
      const int cell = valid ? (int)(info & 0x7ffu) : 0, j = (int)((info >> 11) & 31u);
      int e0 = 0, e1 = 0;                                // lane j of a team takes contributors j, j + s, j + 2s, ...
      if (valid) {
        e0 = (int)offs_s[cell] + j;
        e1 = offs_s[cell + 1];
      }
      const int es = 1 << ls;
      float acc[KCH];
#pragma unroll
      for (int c = 0; c < KCH; ++c) acc[c] = 0.0f;
      for (int e = e0; e < e1; e += es) {
        const float* g = gs + pix_s[e];
        const float wt = wts_s[e];
#pragma unroll
        for (int c = 0; c < KCH; ++c) acc[c] = fmaf(g[c * P], wt, acc[c]);   // rows beyond kl hold stale data; never stored
      }
      for (int d = (1 << ls) >> 1; d > 0; d >>= 1) {     // fixed tree over the team; lane j == 0 ends with the sum
#pragma unroll
        for (int c = 0; c < KCH; ++c) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], d);
      }
      if (valid && j == 0) {
        const int f = cell / ww, rr = cell - f * ww;
        float* dst = a.gcube + (((int64_t)b * 6 + f) * a.C + c0) * ww + rr;
        if (kl == KCH) {
#pragma unroll
          for (int c = 0; c < KCH; ++c) __stcs(dst + c * ww, acc[c]);
        } else {
#pragma unroll
          for (int c = 0; c < KCH; ++c)
            if (c < kl) __stcs(dst + c * ww, acc[c]);
        }
      }
    }
    __syncwarp();                                        // reconverge lane 0 (bulk-load issue) before the block barrier
    __syncthreads();                                     // every thread is done with ring[i & 1]
    if (tid == 0 && i + 2 < n_mine) issue(i + 2);
  }
}

template <int KCH>
__global__ void __launch_bounds__(kC2eSmallThreads)
c2e_bwd_gather_kernel(const float* __restrict__ gequi, const int32_t* __restrict__ offs, const int32_t* __restrict__ pix,
                      const float* __restrict__ wts, float* __restrict__ gcube, int C, int w, int groups) {
  const int ww = w * w, P = 8 * ww, NC = 6 * ww;
  const int b = blockIdx.x / groups, c0 = (blockIdx.x - b * groups) * KCH;
  const int kl = min(KCH, C - c0);
  const float* src = gequi + ((int64_t)b * C + c0) * P;
  for (int cell = threadIdx.x; cell < NC; cell += kC2eSmallThreads) {
    float acc[KCH];
#pragma unroll
    for (int c = 0; c < KCH; ++c) acc[c] = 0.0f;
    const int e0 = __ldg(offs + cell), e1 = __ldg(offs + cell + 1);
    for (int e = e0; e < e1; ++e) {
      const int p = __ldg(pix + e);
      const float wt = __ldg(wts + e);
#pragma unroll
      for (int c = 0; c < KCH; ++c)
        if (c < kl) acc[c] = fmaf(__ldg(src + (int64_t)c * P + p), wt, acc[c]);
    }
    const int f = cell / ww, r = cell - f * ww;
    float* dst = gcube + (((int64_t)b * 6 + f) * C + c0) * ww + r;
#pragma unroll
    for (int c = 0; c < KCH; ++c)
      if (c < kl) __stcs(dst + (int64_t)c * ww, acc[c]);
  }
}

// backward of the fused channel max, single-owner form: a thread owns one cube pixel of one frame and adds the
// routed gradients of its contributors (gsal[b,p] * weight into channel argmax[b,p]) one after the other — no two
// threads ever write the same element, so no atomics and a fixed summation order. gcube is zero-filled first.
// A pole pixel has 46 (w = 8) .. 90 (w = 16) contributors: one dependent read-modify-write per contributor would be
// a chain of that many memory round trips, so the contributors are taken eight at a time — their routing loads and
// the eight current values are independent loads, the in-order additions (a later contributor of the same channel
// continues from the earlier one's sum) happen in registers, then the eight stores go out in order.
constexpr int kMaxBwdBatch = 8;

__global__ void __launch_bounds__(kC2eThreads)
c2e_max_bwd_gather_kernel(const float* __restrict__ gsal, const int32_t* __restrict__ arg, const int32_t* __restrict__ offs,
                          const int32_t* __restrict__ pix, const float* __restrict__ wts, float* gcube,
                          int64_t B, int C, int w) {
  const int ww = w * w, P = 8 * ww, NC = 6 * ww;
  const int cell = blockIdx.x * kC2eThreads + threadIdx.x;
  if (cell >= NC) return;
  const int e0 = __ldg(offs + cell), e1 = __ldg(offs + cell + 1);
  const int f = cell / ww, r = cell - f * ww;
  for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
    float* base = gcube + ((b * 6 + f) * C) * (int64_t)ww + r;
    for (int e = e0; e < e1; e += kMaxBwdBatch) {
      int ch[kMaxBwdBatch];
      float val[kMaxBwdBatch], cur[kMaxBwdBatch];
#pragma unroll
      for (int i = 0; i < kMaxBwdBatch; ++i) {
        ch[i] = -1;
        val[i] = 0.0f;
        if (e + i < e1) {
          const int p = __ldg(pix + e + i);
          const int c = __ldg(arg + b * (int64_t)P + p);
          if ((unsigned)c < (unsigned)C) {               // anything else is never produced by the forward; guards foreign input
            ch[i] = c;
            val[i] = __ldg(gsal + b * (int64_t)P + p) * __ldg(wts + e + i);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kMaxBwdBatch; ++i) cur[i] = ch[i] >= 0 ? base[(int64_t)ch[i] * ww] : 0.0f;
#pragma unroll
      for (int i = 0; i < kMaxBwdBatch; ++i) {
        float acc = cur[i];
#pragma unroll
        for (int j = 0; j < i; ++j)
          if (ch[j] == ch[i]) acc = cur[j];              // the latest earlier contributor of this channel
        cur[i] = acc + val[i];
      }
#pragma unroll
      for (int i = 0; i < kMaxBwdBatch; ++i)
        if (ch[i] >= 0) base[(int64_t)ch[i] * ww] = cur[i];
    }
  }
}

// ---- bicubic variant: Cube2Equi.to_equi_cv2 (utils/cube_to_equi.py:68-91) ---------------------
// cv2.remap(INTER_CUBIC) semantics for float sources (OpenCV remapBicubic): 1/32-pixel fractions
// select rows of the A = -0.75 coefficient table, the 4x4 weight is fl(wy[i]*wx[j]); a window that
// lies fully inside the face is summed row by row ((S0*w0 + S1*w1) + S2*w2) + S3*w3, any other
// window tap by tap with the taps outside the face skipped (BORDER_CONSTANT 0). fp32, no FMA —
// the result is bit-identical to cv2's.
struct CubicTap {
  int face, y0, x0, fx, fy;
};

__device__ __forceinline__ CubicTap decode_cubic_tap(uint32_t t) {
  CubicTap r;
  r.face = (int)(t >> 28);
  r.fy = (int)((t >> 23) & 31u);
  r.fx = (int)((t >> 18) & 31u);
  r.y0 = (int)((t >> 9) & 0x1ffu) - 1;
  r.x0 = (int)(t & 0x1ffu) - 1;
  return r;
}

// OpenCV's interpolateCubic at x = k/32, same operation order, no contraction
__device__ __forceinline__ void cubic_coeffs(int k, float c[4]) {
  const float A = -0.75f;
  const float x = __fmul_rn((float)k, 1.0f / 32);
  const float x1 = __fadd_rn(x, 1.0f), xm = __fsub_rn(1.0f, x);
  c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, x1), 5.0f * A), x1), 8.0f * A), x1), 4.0f * A);
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.0f, x), A + 3.0f), x), x), 1.0f);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.0f, xm), A + 3.0f), xm), xm), 1.0f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, c[0]), c[1]), c[2]);
}

// One output pixel over `kl` channels. src = channel 0 of the pixel's face (plane stride ww),
// dst = channel 0 of the output pixel (plane stride P).
__device__ __forceinline__ void cubic_pixel(const float* __restrict__ src, float* __restrict__ dst,
                                            const CubicTap t, int w, int ww, int P, int kl) {
  float cx[4], cy[4], wt[16];
  cubic_coeffs(t.fx, cx);
  cubic_coeffs(t.fy, cy);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) wt[i * 4 + j] = __fmul_rn(cy[i], cx[j]);
  const int lim = max(w - 3, 0);
  const bool inside = (unsigned)t.x0 < (unsigned)lim && (unsigned)t.y0 < (unsigned)lim;
  if (inside) {
    src += t.y0 * w + t.x0;
#pragma unroll 2
    for (int c = 0; c < kl; ++c, src += ww) {
      float sum = 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float* r = src + i * w;
        float rs = __fadd_rn(__fmul_rn(r[0], wt[i * 4]), __fmul_rn(r[1], wt[i * 4 + 1]));
        rs = __fadd_rn(rs, __fmul_rn(r[2], wt[i * 4 + 2]));
        rs = __fadd_rn(rs, __fmul_rn(r[3], wt[i * 4 + 3]));
        sum = i == 0 ? rs : __fadd_rn(sum, rs);
      }
      __stcs(dst + (int64_t)c * P, sum);
    }
  } else {
    unsigned ok = 0;
    int off[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int yy = t.y0 + i, xx = t.x0 + j;
        const bool v = (unsigned)yy < (unsigned)w && (unsigned)xx < (unsigned)w;
        ok |= (unsigned)v << (i * 4 + j);
        off[i * 4 + j] = v ? yy * w + xx : 0;
      }
#pragma unroll 2
    for (int c = 0; c < kl; ++c, src += ww) {
      float sum = 0.0f;
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (ok & (1u << k)) sum = __fadd_rn(sum, __fmul_rn(src[off[k]], wt[k]));
      __stcs(dst + (int64_t)c * P, sum);
    }
  }
}

// w <= 16: same staging as c2e_small_kernel (block = (b, channel group), six bulk copies). A warp whose lanes mix
// windows inside the face (row-by-row sum) with windows on a border (tap-by-tap sum) executes both code paths for
// every channel — and at w = 8 every run of 32 consecutive equi pixels mixes them. The pixels are therefore handed to
// the lanes in a sorted order (inside windows first, then border windows, each in raster order; built per CTA with
// ballots while the bulk copies are in flight), so all but one warp run a single path. TW = compile-time face width
// (0: run-time), which turns the row / plane strides of the 16 taps into immediates.
__device__ __forceinline__ bool cubic_inside(const CubicTap t, int w) {
  const int lim = max(w - 3, 0);
  return (unsigned)t.x0 < (unsigned)lim && (unsigned)t.y0 < (unsigned)lim;
}

constexpr int kCubicMaxRounds = 4;                        // P = 8 w^2 <= 2048 pixels in rounds of kC2eSmallThreads

template <int TW>
__global__ void __launch_bounds__(kC2eSmallThreads)
c2e_cubic_small_kernel(const float* __restrict__ cube, const uint32_t* __restrict__ taps,
                       float* __restrict__ out, int C, int w_rt, int kch, int groups, int order_off) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  uint16_t* cnt_s = reinterpret_cast<uint16_t*>(smem_raw + 16);             // [kCubicMaxRounds][16 warps] inside windows
  float* cs = reinterpret_cast<float*>(smem_raw + 256);      // [6][kch][w*w]
  uint16_t* order_s = reinterpret_cast<uint16_t*>(smem_raw + order_off);    // [P]
  const int w = TW ? TW : w_rt;
  const int P = 8 * w * w, ww = w * w;
  const int b = blockIdx.x / groups, gidx = blockIdx.x - b * groups;
  const int c0 = gidx * kch, kl = min(kch, C - c0);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int kWarps = kC2eSmallThreads / 32;
  pdl_trigger();
  pdl_wait();
  if (tid == 0) {
    tma::mbar_init(bar, 1);
    tma::fence_mbar_init();
    const uint32_t bytes = (uint32_t)(kl * ww) * 4u;
    tma::mbar_expect_tx(bar, 6u * bytes);
#pragma unroll
    for (int f = 0; f < 6; ++f)
      tma::bulk_load(cs + (size_t)f * kch * ww + kFaceSkew[f], cube + (((int64_t)b * 6 + f) * C + c0) * ww, bytes, bar);
  }
  __syncwarp();
  const int rounds = (P + kC2eSmallThreads - 1) / kC2eSmallThreads;
  unsigned in_mask[kCubicMaxRounds];
#pragma unroll
  for (int r = 0; r < kCubicMaxRounds; ++r) {
    in_mask[r] = 0;
    if (r < rounds) {
      const int pix = r * kC2eSmallThreads + tid;
      const bool in = pix < P && cubic_inside(decode_cubic_tap(__ldg(taps + pix)), w);
      in_mask[r] = __ballot_sync(0xffffffffu, in);
      if (lane == 0) cnt_s[r * kWarps + wid] = (uint16_t)__popc(in_mask[r]);
    }
  }
  __syncthreads();
  int n_in = 0;
  for (int k = 0; k < rounds * kWarps; ++k) n_in += cnt_s[k];
#pragma unroll
  for (int r = 0; r < kCubicMaxRounds; ++r) {
    if (r < rounds) {
      const int first = r * kC2eSmallThreads + wid * 32, pix = first + lane;   // the warp's 32 pixels of this round
      int before_in = 0;                                 // inside windows among the pixels before `first`
      for (int k = 0; k < r * kWarps + wid; ++k) before_in += cnt_s[k];
      const unsigned lt = (1u << lane) - 1u;
      if (pix < P) {
        const int pos = ((in_mask[r] >> lane) & 1u) ? before_in + __popc(in_mask[r] & lt)
                                                    : n_in + (first - before_in) + __popc(~in_mask[r] & lt);
        order_s[pos] = (uint16_t)pix;
      }
    }
  }
  __syncthreads();
  tma::mbar_wait(bar, 0);
  for (int idx = tid; idx < P; idx += kC2eSmallThreads) {
    const int pix = order_s[idx];
    const CubicTap t = decode_cubic_tap(__ldg(taps + pix));
    cubic_pixel(cs + (size_t)t.face * kch * ww + kFaceSkew[t.face], out + ((int64_t)b * C + c0) * P + pix, t, w, ww, P, kl);
  }
}

// 8 < w <= 16, channel-interleaved staging: the kernel above reads one float per (tap, channel) and its 32 lanes
// collide in the banks of a single 49..256-word channel plane — 16 taps x ~2 wavefronts per pixel·channel, which is
// what bounds it (shared-memory bandwidth; ncu: 0.9-1.4 TB/s of DRAM-side bytes). Here the planes of a group of KS
// channels are TRANSPOSED on the way into shared memory ([face][pixel][KS + 4] floats: coalesced global loads,
// scattered 4-byte stores — 5 % of the kernel's shared-memory traffic), so a tap of four channels is ONE 16-byte
// load and lanes on different pixels fall into different bank groups (pixel stride = KS + 4 floats = odd multiple
// of 16 B): half the wavefronts per output value and a quarter of the load instructions. Same arithmetic, bit-exact.
constexpr int kCubicTThreads = 256;

template <int KS>
__global__ void __launch_bounds__(kCubicTThreads)
c2e_cubic_t_kernel(const float* __restrict__ cube, const uint32_t* __restrict__ taps,
                   float* __restrict__ out, int C, int w, int groups) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  constexpr int PS = KS + 4;                                  // pixel stride in floats
  float* in_t = reinterpret_cast<float*>(smem_raw);           // [6][ww][PS]
  const int ww = w * w, P = 8 * ww;
  const int b = blockIdx.x / groups, c0 = (blockIdx.x - b * groups) * KS;
  const int kl = min(KS, C - c0);
  const int tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  const float inv_ww = 1.0f / (float)ww;
  for (int f = 0; f < 6; ++f) {
    const float* src = cube + (((int64_t)b * 6 + f) * C + c0) * ww;
    float* dst = in_t + (size_t)f * ww * PS;
    for (int e = tid; e < kl * ww; e += kCubicTThreads) {
      const int c = (int)(((float)e + 0.5f) * inv_ww);        // e / ww (exact for these sizes)
      dst[(e - c * ww) * PS + c] = __ldg(src + e);
    }
  }
  __syncthreads();
  const int nq = (kl + 3) >> 2;
  const int lim = max(w - 3, 0);
  for (int pix = tid; pix < P; pix += kCubicTThreads) {
    const CubicTap t = decode_cubic_tap(__ldg(taps + pix));
    float cx[4], cy[4];
    cubic_coeffs(t.fx, cx);
    cubic_coeffs(t.fy, cy);
    const bool inside = (unsigned)t.x0 < (unsigned)lim && (unsigned)t.y0 < (unsigned)lim;
    unsigned ok = 0;
    int off[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int yy = t.y0 + i, xx = t.x0 + j;
        const bool v = (unsigned)yy < (unsigned)w && (unsigned)xx < (unsigned)w;
        ok |= (unsigned)v << (i * 4 + j);
        off[i * 4 + j] = v ? (yy * w + xx) * PS : 0;
      }
    const float* base = in_t + (size_t)t.face * ww * PS;
    float* dst = out + ((int64_t)b * C + c0) * P + pix;
    for (int q = 0; q < nq; ++q) {
      float sum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float rs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (inside || (ok & (1u << (i * 4 + j)))) {
            const float4 v = *reinterpret_cast<const float4*>(base + off[i * 4 + j] + 4 * q);
            const float wt = __fmul_rn(cy[i], cx[j]);
            const float s4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float pr = __fmul_rn(s4[k], wt);
              if (inside) rs[k] = j == 0 ? pr : __fadd_rn(rs[k], pr);      // row sum ((S0*w0 + S1*w1) + S2*w2) + S3*w3
              else sum[k] = __fadd_rn(sum[k], pr);                         // border: tap by tap, outside taps skipped
            }
          }
        }
        if (inside) {
#pragma unroll
          for (int k = 0; k < 4; ++k) sum[k] = i == 0 ? rs[k] : __fadd_rn(sum[k], rs[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (4 * q + k < kl) __stcs(dst + (int64_t)(4 * q + k) * P, sum[k]);
    }
  }
}

// any w: taps read through the read-only path (the cube of one frame is L2-resident)
__global__ void __launch_bounds__(kC2eThreads)
c2e_cubic_kernel(const float* __restrict__ cube, const uint32_t* __restrict__ taps,
                 float* __restrict__ out, int64_t B, int C, int w, int ch_per_block) {
  pdl_trigger();
  pdl_wait();
  const int P = 8 * w * w, ww = w * w;
  const int pix = blockIdx.x * kC2eThreads + threadIdx.x;
  if (pix >= P) return;
  const CubicTap t = decode_cubic_tap(__ldg(taps + pix));
  const int c_begin = blockIdx.y * ch_per_block, kl = min(C, c_begin + ch_per_block) - c_begin;
  for (int64_t b = blockIdx.z; b < B; b += gridDim.z)
    cubic_pixel(cube + ((b * 6 + t.face) * C + c_begin) * (int64_t)ww,
                out + (b * C + c_begin) * (int64_t)P + pix, t, w, ww, P, kl);
}

static int check_common(const void* a, const void* taps, const void* wts, const void* o, int64_t B,
                        int64_t C, int w) {
  CP360_CHECK_ARG(B >= 0 && C >= 0 && w > 0, CP360_ERR_BAD_ARG, "bad size");
  CP360_CHECK_ARG(w <= 8191 && C <= 0x7fffffff, CP360_ERR_RANGE, "face width / channels too large");
  if (B == 0 || C == 0) return CP360_OK;
  CP360_CHECK_ARG(a && taps && wts && o, CP360_ERR_BAD_ARG, "null pointer");
  CP360_CHECK_ARG(((uintptr_t)wts % 16) == 0 && ((uintptr_t)a % 4) == 0 && ((uintptr_t)o % 4) == 0 &&
                      ((uintptr_t)taps % 4) == 0, CP360_ERR_ALIGN,
                  "weights must be 16 B aligned, tensors 4 B aligned");
  return require_device();
}

// channel-group size for the shared-memory variant (0 = does not apply)
static int small_plan(const void* cube, int64_t C, int w, size_t* smem) {
  if (w > 16 || ((uintptr_t)cube % 16) != 0) return 0;
  const int ww = w * w;
  int q = 1;
  while ((q * ww) % 4) q <<= 1;                       // 16 B granularity of the bulk copies
  if (C % q) return 0;
  int k = (int)((96 * 1024) / (6 * ww * 4));          // <= 96 KB per block
  k = (int)std::min<int64_t>(k, C);
  k -= k % q;
  if (k < q) return 0;
  *smem = 128 + (size_t)6 * k * ww * 4 + kFaceSkewMax * 4;
  return k;
}

template <int MODE>
static int launch_c2e(const float* cube, const uint32_t* taps, const float* wts, float* out,
                      int64_t B, int64_t C, int w, cudaStream_t st) {
  const int P = 8 * w * w;
  size_t smem = 0;
  int k = small_plan(cube, C, w, &smem);
  if (k > 0) {
    // keep >= ~2 blocks per SM in flight: shrink the group if the grid would be too small
    const int ww = w * w;
    int q = 1;
    while ((q * ww) % 4) q <<= 1;
    while (k > 4 * q && B * ((C + k - 1) / k) < 2 * (int64_t)sm_count()) {
      k = std::max(q, (k / 2) - ((k / 2) % q));
    }
    smem = 128 + (size_t)6 * k * ww * 4 + kFaceSkewMax * 4;
    const int groups = (int)((C + k - 1) / k);
    CP360_CHECK_ARG(B * groups < 0x7fffffff, CP360_ERR_RANGE, "grid too large");
    auto kern = ww == 64 ? c2e_small_kernel<MODE, 64> : ww == 49 ? c2e_small_kernel<MODE, 49> : c2e_small_kernel<MODE, 0>;
    CP360_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_kernel(kern, (unsigned)(B * groups), kC2eSmallThreads, smem, st, cube, taps, reinterpret_cast<const float4*>(wts), out, (int)C, w, k, groups);
    CP360_LAUNCHED();
    return CP360_OK;
  }
  int chb = MODE == 0 ? 8 : 32;   // MODE 1 / 2: fewer atomics per pixel
  chb = (int)std::min<int64_t>(chb, C);
  dim3 grid((P + kC2eThreads - 1) / kC2eThreads, (unsigned)((C + chb - 1) / chb),
            (unsigned)std::min<int64_t>(B, 65535));
  CP360_CHECK_ARG(grid.y <= 65535, CP360_ERR_RANGE, "too many channel chunks");
  launch_kernel(c2e_kernel<MODE>, grid, kC2eThreads, 0, st, cube, taps, reinterpret_cast<const float4*>(wts),
                                                 out, B, (int)C, w, chb);
  CP360_LAUNCHED();
  return CP360_OK;
}

// K3m through the cluster kernel; returns false (nothing launched) when it does not apply.
template <int MODE>
static bool try_c2e_max_cluster(const float* cube, const uint32_t* taps, const float* wts, float* sal, int32_t* arg,
                                int64_t B, int64_t C, int w, cudaStream_t st, int* rc_out) {
  static const bool enabled = [] { const char* v = getenv("CP360_C2E_CLUSTER"); return !(v && *v == '0'); }();
  if (!enabled || w > 16 || B > 0x3fffffff) return false;
  const int ww = w * w, P = 8 * ww;
  if (((uintptr_t)cube % 16) != 0) return false;
  int q = 1;
  while ((q * ww) % 4) q <<= 1;                       // 16 B granularity of the bulk copies
  if (C % q) return false;
  C2eMaxArgs a;
  a.cube = cube; a.taps = taps; a.wts = reinterpret_cast<const float4*>(wts); a.sal = sal; a.arg = arg;
  a.C = (int)C; a.w = w;
  a.K = std::max(q, (int)((24 * 1024) / (6 * ww * 4)) / q * q);
  int G = 8;
  if (const char* v = getenv("CP360_C2E_CLUSTER_SIZE")) G = std::max(1, std::min(8, atoi(v)));   // experiments (power of two)
  while (G > 1 && (B * G > 4 * (int64_t)sm_count() || (C + G - 1) / G < a.K)) G >>= 1;   // enough CTAs, >= one stage each
  a.Cg = (int)(((C + G - 1) / G + q - 1) / q * q);
  a.K = std::min(a.K, a.Cg);
  a.stages = 3;
  a.stage_floats = 6 * a.K * ww + kFaceSkewMax;
  a.ring_off = (64 + 8 * P + 127) & ~127;
  const size_t smem = (size_t)a.ring_off + (size_t)a.stages * a.stage_floats * 4;
  if (smem > 200 * 1024) return false;
  void (*kern)(const C2eMaxArgs) = ww == 64   ? c2e_max_cluster_kernel<MODE, 1, 64>
                                   : ww == 49 ? c2e_max_cluster_kernel<MODE, 1, 49>
                                   : P <= kC2eSmallThreads ? c2e_max_cluster_kernel<MODE, 1, 0>
                                                           : c2e_max_cluster_kernel<MODE, kC2eMaxPix, 0>;
  if (P > kC2eSmallThreads * kC2eMaxPix) return false;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  cudaError_t e = launch_kernel_cluster(kern, dim3((unsigned)(B * G)), dim3(kC2eSmallThreads), smem, st, (unsigned)G, a);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("c2e cluster launch failed: %s", cudaGetErrorString(e));
    *rc_out = CP360_ERR_CUDA;
    return true;
  }
  count_launch();
  *rc_out = CP360_OK;
  return true;
}

static int check_bwd_plan(const void* offs, const void* pix, const void* wts) {
  CP360_CHECK_ARG(offs && pix && wts, CP360_ERR_BAD_ARG, "null backward plan");
  CP360_CHECK_ARG(((uintptr_t)offs % 4) == 0 && ((uintptr_t)pix % 4) == 0 && ((uintptr_t)wts % 4) == 0, CP360_ERR_ALIGN,
                  "backward plan must be 4 B aligned");
  return CP360_OK;
}

}  // namespace cp360

using namespace cp360;

extern "C" {

int cp360_c2e_fwd(const float* cube, const uint32_t* taps, const float* wts, float* equi, int64_t B,
                  int64_t C, int w, void* stream) {
  int rc = check_common(cube, taps, wts, equi, B, C, w);
  if (rc != CP360_OK || B == 0 || C == 0) return rc;
  return launch_c2e<0>(cube, taps, wts, equi, B, C, w, (cudaStream_t)stream);
}

int cp360_c2e_max_fwd(const float* cube, const uint32_t* taps, const float* wts, float* sal,
                      int64_t B, int64_t C, int w, void* stream) {
  int rc = check_common(cube, taps, wts, sal, B, C, w);
  if (rc != CP360_OK || B == 0) return rc;
  CP360_CHECK_ARG(C > 0, CP360_ERR_BAD_ARG, "channel max over zero channels");
  cudaStream_t st = (cudaStream_t)stream;
  if (try_c2e_max_cluster<1>(cube, taps, wts, sal, nullptr, B, C, w, st, &rc)) return rc;
  // large faces: -inf fill, then per-chunk running maxima combined with an order-preserving atomic max
  const int64_t n = B * 8 * (int64_t)w * w;
  launch_kernel(fill_kernel, (unsigned)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, st, sal, n, -INFINITY);
  CP360_LAUNCHED();
  return launch_c2e<1>(cube, taps, wts, sal, B, C, w, st);
}

int cp360_c2e_max_arg_fwd(const float* cube, const uint32_t* taps, const float* wts, float* sal,
                          int32_t* argmax, uint64_t* scratch, int64_t B, int64_t C, int w, void* stream) {
  int rc = check_common(cube, taps, wts, sal, B, C, w);
  if (rc != CP360_OK || B == 0) return rc;
  CP360_CHECK_ARG(C > 0, CP360_ERR_BAD_ARG, "channel max over zero channels");
  CP360_CHECK_ARG(argmax && ((uintptr_t)argmax % 4) == 0, CP360_ERR_BAD_ARG, "argmax null or misaligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (try_c2e_max_cluster<2>(cube, taps, wts, sal, argmax, B, C, w, st, &rc)) return rc;   // w <= 16: no scratch needed
  CP360_CHECK_ARG(scratch && ((uintptr_t)scratch % 8) == 0, CP360_ERR_BAD_ARG,
                  "faces wider than 16 need an 8 B-aligned scratch buffer of B*2w*4w uint64");
  const int64_t n = B * 8 * (int64_t)w * w;
  CP360_CUDA_OK(cudaMemsetAsync(scratch, 0, (size_t)n * sizeof(uint64_t), st));   // below every key
  rc = launch_c2e<2>(cube, taps, wts, reinterpret_cast<float*>(scratch), B, C, w, st);
  if (rc != CP360_OK) return rc;
  launch_kernel(c2e_argmax_decode_kernel, (unsigned)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, st,
                reinterpret_cast<const unsigned long long*>(scratch), sal, argmax, n);
  CP360_LAUNCHED();
  return CP360_OK;
}

int cp360_c2e_max_bwd(const float* gsal, const int32_t* argmax, const int32_t* offs, const int32_t* pix, const float* bwts,
                      float* gcube, int64_t B, int64_t C, int w, void* stream) {
  CP360_CHECK_ARG(B >= 0 && C >= 0 && w > 0 && w <= 8191 && C <= 0x7fffffff, CP360_ERR_BAD_ARG, "bad size");
  if (B == 0 || C == 0) return CP360_OK;
  CP360_CHECK_ARG(gsal && gcube && argmax, CP360_ERR_BAD_ARG, "null pointer");
  CP360_CHECK_ARG(((uintptr_t)gsal % 4) == 0 && ((uintptr_t)gcube % 4) == 0 && ((uintptr_t)argmax % 4) == 0, CP360_ERR_ALIGN,
                  "tensors must be 4 B aligned");
  int rc = check_bwd_plan(offs, pix, bwts);
  if (rc != CP360_OK) return rc;
  rc = require_device();
  if (rc != CP360_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CP360_CUDA_OK(cudaMemsetAsync(gcube, 0, (size_t)B * 6 * C * w * w * sizeof(float), st));
  dim3 grid((6 * w * w + kC2eThreads - 1) / kC2eThreads, (unsigned)std::min<int64_t>(B, 65535));
  c2e_max_bwd_gather_kernel<<<grid, kC2eThreads, 0, st>>>(gsal, argmax, offs, pix, bwts, gcube, B, (int)C, w);
  CP360_LAUNCHED();
  return CP360_OK;
}

int cp360_c2e_cubic_fwd(const float* cube, const uint32_t* taps, float* equi, int64_t B, int64_t C,
                        int w, void* stream) {
  CP360_CHECK_ARG(B >= 0 && C >= 0 && w > 0, CP360_ERR_BAD_ARG, "bad size");
  CP360_CHECK_ARG(w <= 512 && C <= 0x7fffffff, CP360_ERR_RANGE, "face width > 512 / too many channels");
  if (B == 0 || C == 0) return CP360_OK;
  CP360_CHECK_ARG(cube && taps && equi, CP360_ERR_BAD_ARG, "null pointer");
  CP360_CHECK_ARG(((uintptr_t)cube % 4) == 0 && ((uintptr_t)equi % 4) == 0 && ((uintptr_t)taps % 4) == 0,
                  CP360_ERR_ALIGN, "tensors must be 4 B aligned");
  int rc = require_device();
  if (rc != CP360_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = 8 * w * w;
  static const bool transposed = [] { const char* v = getenv("CP360_CUBIC_T"); return !(v && *v == '0'); }();
  // channel-interleaved staging (any C, any alignment). Measured on B200 at B = 32: [192,256,16,16] 111 us vs 129 us for
  // the plane-major kernel, but 104 vs 95 us at [192,1000,8,8] — so it is the default only above w = 8
  // (CP360_CUBIC_T=2 forces it for every w <= 16, =0 disables it)
  static const bool transposed_all = [] { const char* v = getenv("CP360_CUBIC_T"); return v && *v == '2'; }();
  if (w <= 16 && transposed && (w > 8 || transposed_all)) {
    const int ks = w <= 8 ? 32 : 8;
    const int64_t groups = (C + ks - 1) / ks;
    CP360_CHECK_ARG(B * groups < 0x7fffffff, CP360_ERR_RANGE, "grid too large");
    const size_t smem_t = (size_t)6 * w * w * (ks + 4) * 4;
    void (*kern)(const float*, const uint32_t*, float*, int, int, int) = ks == 32 ? c2e_cubic_t_kernel<32> : c2e_cubic_t_kernel<8>;
    CP360_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    launch_kernel(kern, (unsigned)(B * groups), kCubicTThreads, smem_t, st, cube, taps, equi, (int)C, w, (int)groups);
    CP360_LAUNCHED();
    return CP360_OK;
  }
  size_t smem = 0;
  int k = small_plan(cube, C, w, &smem);
  if (k > 0) {
    const int ww = w * w;
    int q = 1;
    while ((q * ww) % 4) q <<= 1;
    while (k > 4 * q && B * ((C + k - 1) / k) < 2 * (int64_t)sm_count()) k = std::max(q, (k / 2) - ((k / 2) % q));
    if (const char* e = std::getenv("CP360_CUBIC_K")) k = std::max(q, std::min(k, std::atoi(e) / q * q));   // experiments
    const int order_off = (int)((256 + ((size_t)6 * k * ww + kFaceSkewMax) * 4 + 15) & ~(size_t)15);
    smem = (size_t)order_off + (size_t)P * 2;
    const int groups = (int)((C + k - 1) / k);
    CP360_CHECK_ARG(B * groups < 0x7fffffff, CP360_ERR_RANGE, "grid too large");
    void (*kern)(const float*, const uint32_t*, float*, int, int, int, int, int) =
        w == 8 ? c2e_cubic_small_kernel<8> : w == 7 ? c2e_cubic_small_kernel<7> : c2e_cubic_small_kernel<0>;
    CP360_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_kernel(kern, (unsigned)(B * groups), kC2eSmallThreads, smem, st, cube, taps, equi, (int)C, w, k, groups, order_off);
    CP360_LAUNCHED();
    return CP360_OK;
  }
  const int chb = (int)std::min<int64_t>(8, C);
  dim3 grid((P + kC2eThreads - 1) / kC2eThreads, (unsigned)((C + chb - 1) / chb),
            (unsigned)std::min<int64_t>(B, 65535));
  CP360_CHECK_ARG(grid.y <= 65535, CP360_ERR_RANGE, "too many channel chunks");
  launch_kernel(c2e_cubic_kernel, grid, kC2eThreads, 0, st, cube, taps, equi, B, (int)C, w, chb);
  CP360_LAUNCHED();
  return CP360_OK;
}

int cp360_c2e_bwd(const float* gequi, const int32_t* offs, const int32_t* pix, const float* bwts, int n_entries,
                  float* gcube, int64_t B, int64_t C, int w, void* stream) {
  CP360_CHECK_ARG(B >= 0 && C >= 0 && w > 0 && w <= 8191 && C <= 0x7fffffff, CP360_ERR_BAD_ARG, "bad size");
  CP360_CHECK_ARG(n_entries >= 0 && (int64_t)n_entries <= (int64_t)32 * w * w, CP360_ERR_BAD_ARG,
                  "n_entries must be offsets[6*w*w] of cp360_c2e_build_bwd_plan (<= 32*w*w)");
  if (B == 0 || C == 0) return CP360_OK;
  CP360_CHECK_ARG(gequi && gcube, CP360_ERR_BAD_ARG, "null pointer");
  CP360_CHECK_ARG(((uintptr_t)gequi % 4) == 0 && ((uintptr_t)gcube % 4) == 0, CP360_ERR_ALIGN, "tensors must be 4 B aligned");
  int rc = check_bwd_plan(offs, pix, bwts);
  if (rc != CP360_OK) return rc;
  rc = require_device();
  if (rc != CP360_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int P = 8 * w * w, NC = 6 * w * w;
  const char* sw = getenv("CP360_C2E_BWD_SMALL");        // "0": the read-only-path gather for every w (A/B, racecheck_probe.py)
  const bool small = w <= 16 && ((uintptr_t)gequi % 16) == 0 && n_entries <= 65535 && !(sw && *sw == '0');   // 16-bit ids in shared memory
  if (small) {
    C2eBwdArgs a;
    a.gequi = gequi; a.offs = offs; a.pix = pix; a.wts = bwts; a.gcube = gcube; a.C = (int)C; a.w = w;
    a.n_entries = n_entries;
    const int kch = w <= 8 ? 16 : 8;
    const int64_t groups = (C + kch - 1) / kch;
    a.split = kC2eBwdSplit;
    a.order = 1;
    if (const char* e = std::getenv("CP360_C2E_BWD_SPLIT")) a.split = std::max(2, std::min(64, std::atoi(e)));   // experiments
    if (const char* e = std::getenv("CP360_C2E_BWD_ORDER")) a.order = std::atoi(e) != 0;
    a.slot_cap = NC + 2 * n_entries / a.split + 32 * kC2eBwdClasses;   // teams: sum of sizes < NC + 2 * entries / split
    a.slot_off = (256 + 2 * (NC + 1) + 15) & ~15;
    a.pix_off = a.slot_off + 4 * a.slot_cap;
    a.wts_off = (a.pix_off + 2 * n_entries + 15) & ~15;
    a.ring_off = (a.wts_off + 4 * n_entries + 127) & ~127;
    const size_t smem = (size_t)a.ring_off + (size_t)2 * kch * P * 4;
    if (smem <= 200 * 1024) {
      void (*kern)(const C2eBwdArgs) = w == 8    ? c2e_bwd_small_kernel<16, 8>       // the reference's 224 / 256 px score maps
                                       : w == 7  ? c2e_bwd_small_kernel<16, 7>
                                       : w == 16 ? c2e_bwd_small_kernel<8, 16>
                                       : w == 14 ? c2e_bwd_small_kernel<8, 14>
                                       : kch == 16 ? c2e_bwd_small_kernel<16, 0> : c2e_bwd_small_kernel<8, 0>;
      CP360_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 1;                                     // one resident wave: CTA r of a frame walks groups r, r + G, ...
      CP360_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kC2eSmallThreads, smem));
      const int64_t slots = (int64_t)sm_count() * std::max(1, per_sm);
      a.G = (int)std::max<int64_t>(1, std::min<int64_t>(groups, slots / B));
      if (const char* e = std::getenv("CP360_C2E_BWD_G")) a.G = (int)std::max<int64_t>(1, std::min<int64_t>(groups, std::atoi(e)));
      CP360_CHECK_ARG(B * a.G < 0x7fffffff, CP360_ERR_RANGE, "grid too large");
      launch_kernel(kern, (unsigned)(B * a.G), kC2eSmallThreads, smem, st, a);
      CP360_LAUNCHED();
      return CP360_OK;
    }
  }
  const int kch = 8;
  const int64_t groups = (C + kch - 1) / kch;
  CP360_CHECK_ARG(B * groups < 0x7fffffff, CP360_ERR_RANGE, "grid too large");
  c2e_bwd_gather_kernel<8><<<(unsigned)(B * groups), kC2eSmallThreads, 0, st>>>(gequi, offs, pix, bwts, gcube, (int)C, w, (int)groups);
  CP360_LAUNCHED();
  return CP360_OK;
}

}  // extern "C"

#ifdef CP360_TRACE
extern "C" __attribute__((visibility("default"))) int cp360_trace_bind_c2e(void* rec, unsigned cap, void* n) {
  cp360::TraceBuf tb = {(cp360::TraceRec*)rec, cap, (unsigned*)n};
  return cudaMemcpyToSymbol(cp360::g_tb, &tb, sizeof(tb)) == cudaSuccess ? 0 : CP360_ERR_CUDA;
}
#endif
