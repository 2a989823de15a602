// Equirectangular -> cube faces for sm_100a — replaces Equi2Cube.to_cube
// (utils/equi_to_cube.py:112-129: 18 cv2.remap calls per frame on the CPU).
//
// One thread per output pixel. The per-resolution packed map (x0:11|y0:10|fx:5|fy:5, built on
// the host by cp360_e2c_build_map) is read coalesced; the 2x2 taps of all C channels are
// contiguous in the HWC frame (2*C floats per row) and neighbouring output pixels sample
// neighbouring input pixels, so a warp's gathers fall into a few 128 B lines per tap row.
// The arithmetic is cv2's: weights (1-fy)(1-fx).. with f = k/32 exact in fp32, the four
// products accumulated left to right with separate multiply and add (no FMA) -> bit-identical
// to cv2.remap(INTER_LINEAR) on fp32 input.
#include <algorithm>

#include "common.cuh"
#include "cubepad_geom.h"

namespace cp360 {

struct NormParams { float mean[4]; float std[4]; };

constexpr int kE2cTileX = 32, kE2cTileY = 8;

// Fused CubePad (e2c + im_norm + HWC->NCHW + CubePad(3) in one kernel — the input
// of conv1, model/resnet_cubic.py:116-117). With WithPad the grid covers the PADDED faces; a thread
// resolves its output pixel to the (face, pixel) CubePad copies it from (cubepad_geom.h, the table
// the CubePad kernels use) and samples that pixel's map entry, so the halo is resampled from the
// frame (+4.7 % gathers at 256/p3, all cache hits) instead of being copied from a faces tensor that
// is never written. Values are bit-identical to CubePad(to_cube(frame)).
struct NoPad { static constexpr bool kPad = false; };
struct WithPad { static constexpr bool kPad = true; CubePadGeom g; };

struct E2cPix { int f, map_idx, out_pix, out_plane; };

// One map entry. Frames up to 2047 x 1023 use the packed word x0:11 | y0:10 | fx:5 | fy:5; larger frames
// (4K / 8K equirects) the wide form of two words per pixel, x0:16 | y0:16 and fx:5 | fy:5, read as one
// 8-byte load (cp360_e2c_build_map writes whichever form the frame size selects, cp360_e2c_map_words).
struct MapEntry { int x0, y0; float fx, fy; };

__device__ __forceinline__ MapEntry load_map(const uint32_t* __restrict__ packed, int idx, int Hin, int Win) {
  MapEntry m;
  if (Win > 2047 || Hin > 1023) {                      // uniform across the grid
    const uint2 p = __ldg(reinterpret_cast<const uint2*>(packed) + idx);
    m.x0 = (int)(p.x >> 16); m.y0 = (int)(p.x & 0xffffu);
    m.fx = (float)((p.y >> 5) & 31u) * 0.03125f; m.fy = (float)(p.y & 31u) * 0.03125f;
  } else {
    const uint32_t p = __ldg(packed + idx);
    m.x0 = (int)(p >> 20); m.y0 = (int)((p >> 10) & 1023u);
    m.fx = (float)((p >> 5) & 31u) * 0.03125f; m.fy = (float)(p & 31u) * 0.03125f;
  }
  return m;
}

template <class GEOM>
__device__ __forceinline__ bool e2c_locate(const GEOM& geom, int w, E2cPix* q) {
  const int ox = blockIdx.x * kE2cTileX + threadIdx.x;
  if constexpr (GEOM::kPad) {
    const int Ho = geom.g.Ho, Wo = geom.g.Wo;
    const int tiles_y = (Ho + kE2cTileY - 1) / kE2cTileY;
    q->f = blockIdx.y / tiles_y;
    const int oy = (blockIdx.y - q->f * tiles_y) * kE2cTileY + threadIdx.y;
    if (ox >= Wo || oy >= Ho) return false;
    int sf;
    const int sp = cubepad_src(geom.g, q->f, oy, ox, &sf);
    q->map_idx = sf * w * w + sp;
    q->out_pix = oy * Wo + ox;
    q->out_plane = Ho * Wo;
  } else {
    const int tiles_y = (w + kE2cTileY - 1) / kE2cTileY;
    q->f = blockIdx.y / tiles_y;
    const int oy = (blockIdx.y - q->f * tiles_y) * kE2cTileY + threadIdx.y;
    if (ox >= w || oy >= w) return false;
    q->map_idx = q->f * w * w + oy * w + ox;
    q->out_pix = oy * w + ox;
    q->out_plane = w * w;
  }
  return true;
}

template <int C>
__device__ __forceinline__ void load_px(const float* __restrict__ p, float (&v)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = __ldg(p + c);
}

template <int C, int LAYOUT, bool NORM, class GEOM>
__global__ void __launch_bounds__(kE2cTileX * kE2cTileY)
e2c_kernel(const float* __restrict__ frames, const uint32_t* __restrict__ packed,
           float* __restrict__ faces, int64_t B, int Hin, int Win, int w, NormParams nrm,
           const __grid_constant__ GEOM geom) {
  pdl_trigger();
  pdl_wait();
  E2cPix q;
  if (!e2c_locate(geom, w, &q)) return;
  const int f = q.f, ww = q.out_plane, pix = q.out_pix;     // output plane size / pixel (padded or not)
  const MapEntry me = load_map(packed, q.map_idx, Hin, Win);
  const int x0 = me.x0, y0 = me.y0;
  const float fx = me.fx, fy = me.fy;
  const float w00 = __fmul_rn(1.0f - fy, 1.0f - fx), w01 = __fmul_rn(1.0f - fy, fx);
  const float w10 = __fmul_rn(fy, 1.0f - fx), w11 = __fmul_rn(fy, fx);
  const bool x1_ok = x0 + 1 < Win, y1_ok = y0 + 1 < Hin;   // BORDER_CONSTANT 0

  for (int64_t b = blockIdx.z; b < B; b += gridDim.z) {
    const float* row0 = frames + ((b * Hin + y0) * (int64_t)Win + x0) * C;
    const float* row1 = row0 + (int64_t)Win * C;
    float s00[C], s01[C], s10[C], s11[C];
    load_px<C>(row0, s00);
    if (x1_ok) load_px<C>(row0 + C, s01);
    if (y1_ok) load_px<C>(row1, s10);
    if (x1_ok && y1_ok) load_px<C>(row1 + C, s11);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float a01 = x1_ok ? s01[c] : 0.0f, a10 = y1_ok ? s10[c] : 0.0f;
      const float a11 = (x1_ok && y1_ok) ? s11[c] : 0.0f;
      float v = __fmul_rn(s00[c], w00);
      v = __fadd_rn(v, __fmul_rn(a01, w01));
      v = __fadd_rn(v, __fmul_rn(a10, w10));
      v = __fadd_rn(v, __fmul_rn(a11, w11));
      if (NORM) v = __fdiv_rn(__fsub_rn(v, nrm.mean[c]), nrm.std[c]);   // utils/utils.py:28-33
      if (LAYOUT == CP360_LAYOUT_NCHW)
        __stcs(faces + ((b * 6 + f) * C + c) * (int64_t)ww + pix, v);
      else
        faces[((b * 6 + f) * (int64_t)ww + pix) * C + c] = v;
    }
  }
}

// C == 3, 16 B-aligned rows and frames: the six floats of a tap pair (pixels x0, x0+1) are 24
// contiguous bytes at a 4 B-aligned address. Six scalar loads per row make the L1 serve the same
// ~25 sectors six times per warp (ncu: 25 sectors/request, lg_throttle); instead each lane fetches
// the 16 B-aligned window that holds them with two (three when it starts 12 B in) 128-bit loads and
// shifts its floats into place with a two-stage select (by 2 words, then by 1). A thread keeps
// its pixel (map entry, weights, window offsets) for kE2cFramesPerThread frames. Same arithmetic
// as e2c_kernel, bit-identical results.
constexpr int kE2cFramesPerThread = 4;

// v[0..5] = w[o..o+5]; hi = w[8] (needed for o == 3 only)
__device__ __forceinline__ void realign6(const float4& a, const float4& b, float hi, int o, float (&v)[6]) {
  const float w[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, hi};
  const bool s2 = (o & 2) != 0, s1 = (o & 1) != 0;
  float t[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) t[k] = s2 ? w[k + 2] : w[k];
#pragma unroll
  for (int k = 0; k < 6; ++k) v[k] = s1 ? t[k + 1] : t[k];
}

template <int LAYOUT, bool NORM, class GEOM>
__global__ void __launch_bounds__(kE2cTileX * kE2cTileY)
e2c_kernel_c3v(const float* __restrict__ frames, const uint32_t* __restrict__ packed,
               float* __restrict__ faces, int64_t B, int Hin, int Win, int w, NormParams nrm,
               const __grid_constant__ GEOM geom) {
  constexpr int C = 3;
  CP360_TRACE_BEGIN(3)
  pdl_trigger();
  pdl_wait();
  CP360_TRACE_T0(1);
  E2cPix q;
  if (!e2c_locate(geom, w, &q)) return;
  const int f = q.f, ww = q.out_plane, pix = q.out_pix;     // output plane size / pixel (padded or not)
  const MapEntry me = load_map(packed, q.map_idx, Hin, Win);
  const int x0 = me.x0, y0 = me.y0;
  const float fx = me.fx, fy = me.fy;
  const float w00 = __fmul_rn(1.0f - fy, 1.0f - fx), w01 = __fmul_rn(1.0f - fy, fx);
  const float w10 = __fmul_rn(fy, 1.0f - fx), w11 = __fmul_rn(fy, fx);
  const bool x1_ok = x0 + 1 < Win, y1_ok = y0 + 1 < Hin;   // BORDER_CONSTANT 0
  // window of row y0 inside a frame, in floats; frames and rows are 16 B aligned, so the window
  // alignment o is the same for both rows and every frame
  const int frame_floats = Hin * Win * C, row_floats = Win * C;
  const int e0 = (y0 * Win + x0) * C;
  const int o = e0 & 3;
  const int a0 = e0 - o;                                   // aligned start of the row-y0 window
  const int need = o == 3 ? 12 : 8;                        // floats fetched per window
  // only the last frame can run off the end of the buffer (elsewhere the overrun lands in the next frame)
  const bool risky = (y1_ok ? a0 + row_floats : a0) + need > frame_floats;

  for (int64_t b_begin = (int64_t)blockIdx.z * kE2cFramesPerThread; b_begin < B;
       b_begin += (int64_t)gridDim.z * kE2cFramesPerThread) {
  const int64_t b_end = min(B, b_begin + kE2cFramesPerThread);
#pragma unroll 2
  for (int64_t b = b_begin; b < b_end; ++b) {
    const float* fr = frames + b * frame_floats;
    const float4* q0 = reinterpret_cast<const float4*>(fr + a0);
    const float4* q1 = reinterpret_cast<const float4*>(fr + a0 + row_floats);
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 r0a, r0b, r1a = z, r1b = z;
    float r0c = 0.f, r1c = 0.f;
    if (!(risky && b == B - 1)) {
      r0a = __ldg(q0); r0b = __ldg(q0 + 1);
      if (o == 3) r0c = __ldg(reinterpret_cast<const float*>(q0 + 2));
      if (y1_ok) {
        r1a = __ldg(q1); r1b = __ldg(q1 + 1);
        if (o == 3) r1c = __ldg(reinterpret_cast<const float*>(q1 + 2));
      }
    } else {                                               // tail of the buffer: guarded element loads
      float t0[9], t1[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        t0[k] = a0 + k < frame_floats ? __ldg(fr + a0 + k) : 0.f;
        t1[k] = (y1_ok && a0 + row_floats + k < frame_floats) ? __ldg(fr + a0 + row_floats + k) : 0.f;
      }
      r0a = make_float4(t0[0], t0[1], t0[2], t0[3]); r0b = make_float4(t0[4], t0[5], t0[6], t0[7]); r0c = t0[8];
      r1a = make_float4(t1[0], t1[1], t1[2], t1[3]); r1b = make_float4(t1[4], t1[5], t1[6], t1[7]); r1c = t1[8];
    }
    float v0[6], v1[6];
    realign6(r0a, r0b, r0c, o, v0);
    realign6(r1a, r1b, r1c, o, v1);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float s00 = v0[c];
      const float a01 = x1_ok ? v0[C + c] : 0.0f;
      const float a10 = y1_ok ? v1[c] : 0.0f;
      const float a11 = (x1_ok && y1_ok) ? v1[C + c] : 0.0f;
      float v = __fmul_rn(s00, w00);
      v = __fadd_rn(v, __fmul_rn(a01, w01));
      v = __fadd_rn(v, __fmul_rn(a10, w10));
      v = __fadd_rn(v, __fmul_rn(a11, w11));
      if (NORM) v = __fdiv_rn(__fsub_rn(v, nrm.mean[c]), nrm.std[c]);   // utils/utils.py:28-33
      if (LAYOUT == CP360_LAYOUT_NCHW)
        __stcs(faces + ((b * 6 + f) * C + c) * (int64_t)ww + pix, v);
      else
        faces[((b * 6 + f) * (int64_t)ww + pix) * C + c] = v;
    }
  }
  }
  CP360_TRACE_T0(3);
}

// any channel count (no normalisation)
template <int LAYOUT, class GEOM>
__global__ void __launch_bounds__(kE2cTileX * kE2cTileY)
e2c_kernel_anyc(const float* __restrict__ frames, const uint32_t* __restrict__ packed,
                float* __restrict__ faces, int64_t B, int Hin, int Win, int C, int w,
                const __grid_constant__ GEOM geom) {
  pdl_trigger();
  pdl_wait();
  E2cPix q;
  if (!e2c_locate(geom, w, &q)) return;
  const int f = q.f, ww = q.out_plane, pix = q.out_pix;     // output plane size / pixel (padded or not)
  const MapEntry me = load_map(packed, q.map_idx, Hin, Win);
  const int x0 = me.x0, y0 = me.y0;
  const float fx = me.fx, fy = me.fy;
  const float w00 = __fmul_rn(1.0f - fy, 1.0f - fx), w01 = __fmul_rn(1.0f - fy, fx);
  const float w10 = __fmul_rn(fy, 1.0f - fx), w11 = __fmul_rn(fy, fx);
  const bool x1_ok = x0 + 1 < Win, y1_ok = y0 + 1 < Hin;
  for (int64_t b = blockIdx.z; b < B; b += gridDim.z) {
    const float* row0 = frames + ((b * Hin + y0) * (int64_t)Win + x0) * C;
    const float* row1 = row0 + (int64_t)Win * C;
    for (int c = 0; c < C; ++c) {
      const float a00 = __ldg(row0 + c);
      const float a01 = x1_ok ? __ldg(row0 + C + c) : 0.0f;
      const float a10 = y1_ok ? __ldg(row1 + c) : 0.0f;
      const float a11 = (x1_ok && y1_ok) ? __ldg(row1 + C + c) : 0.0f;
      float v = __fmul_rn(a00, w00);
      v = __fadd_rn(v, __fmul_rn(a01, w01));
      v = __fadd_rn(v, __fmul_rn(a10, w10));
      v = __fadd_rn(v, __fmul_rn(a11, w11));
      if (LAYOUT == CP360_LAYOUT_NCHW)
        faces[((b * 6 + f) * C + c) * (int64_t)ww + pix] = v;
      else
        faces[((b * 6 + f) * (int64_t)ww + pix) * C + c] = v;
    }
  }
}

// uint8 frames (what a video decoder / PIL resize hands over, dataset_feat_extractor.py:126-142):
// pixel value = float32(u8) / denom — for denom = 255 this equals the reference's
// float32(u8 / 255.0) for all 256 codes (tests/test_gpu_parity.py) — then the same fixed-point
// bilinear arithmetic. The 256 quotients live in shared memory; a tap pair of a C = 3 row is six
// contiguous bytes, fetched as three aligned words and realigned with funnel shifts. H2D traffic
// and DRAM reads are a quarter of the fp32 path.
template <int LAYOUT, bool NORM, class GEOM>
__global__ void __launch_bounds__(kE2cTileX * kE2cTileY)
e2c_kernel_u8c3(const uint8_t* __restrict__ frames, const uint32_t* __restrict__ packed,
                float* __restrict__ faces, int64_t B, int Hin, int Win, int w, float denom,
                NormParams nrm, const __grid_constant__ GEOM geom) {
  constexpr int C = 3;
  __shared__ float lut[256];
  const int tid = threadIdx.y * kE2cTileX + threadIdx.x;
  pdl_trigger();
  lut[tid] = __fdiv_rn((float)tid, denom);                 // block is 256 threads
  __syncthreads();
  pdl_wait();
  E2cPix q;
  if (!e2c_locate(geom, w, &q)) return;
  const int f = q.f, ww = q.out_plane, pix = q.out_pix;     // output plane size / pixel (padded or not)
  const MapEntry me = load_map(packed, q.map_idx, Hin, Win);
  const int x0 = me.x0, y0 = me.y0;
  const float fx = me.fx, fy = me.fy;
  const float w00 = __fmul_rn(1.0f - fy, 1.0f - fx), w01 = __fmul_rn(1.0f - fy, fx);
  const float w10 = __fmul_rn(fy, 1.0f - fx), w11 = __fmul_rn(fy, fx);
  const bool x1_ok = x0 + 1 < Win, y1_ok = y0 + 1 < Hin;   // BORDER_CONSTANT 0
  const int64_t total_words = (B * (int64_t)Hin * Win * C + 3) >> 2;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(frames);

  for (int64_t b = blockIdx.z; b < B; b += gridDim.z) {
    const int64_t byte0 = ((b * Hin + y0) * (int64_t)Win + x0) * C;
    uint32_t px[2][2];                                      // [row][0: bytes 0-3, 1: bytes 4-5]
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int64_t bb = byte0 + (int64_t)r * Win * C;
      const int64_t wi = bb >> 2;
      const int sh = (int)(bb & 3) * 8;
      uint32_t a0 = 0, a1 = 0, a2 = 0;
      if (r == 0 || y1_ok) {
        a0 = __ldg(words + wi);
        a1 = wi + 1 < total_words ? __ldg(words + wi + 1) : 0u;
        a2 = (sh == 24 && wi + 2 < total_words) ? __ldg(words + wi + 2) : 0u;
      }
      px[r][0] = __funnelshift_r(a0, a1, sh);
      px[r][1] = __funnelshift_r(a1, a2, sh);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const uint32_t t00 = (px[0][0] >> (8 * c)) & 0xffu;
      const uint32_t t01 = c == 0 ? (px[0][0] >> 24) : ((px[0][1] >> (8 * (c - 1))) & 0xffu);
      const uint32_t t10 = (px[1][0] >> (8 * c)) & 0xffu;
      const uint32_t t11 = c == 0 ? (px[1][0] >> 24) : ((px[1][1] >> (8 * (c - 1))) & 0xffu);
      const float s00 = lut[t00];
      const float a01 = x1_ok ? lut[t01] : 0.0f;
      const float a10 = y1_ok ? lut[t10] : 0.0f;
      const float a11 = (x1_ok && y1_ok) ? lut[t11] : 0.0f;
      float v = __fmul_rn(s00, w00);
      v = __fadd_rn(v, __fmul_rn(a01, w01));
      v = __fadd_rn(v, __fmul_rn(a10, w10));
      v = __fadd_rn(v, __fmul_rn(a11, w11));
      if (NORM) v = __fdiv_rn(__fsub_rn(v, nrm.mean[c]), nrm.std[c]);
      if (LAYOUT == CP360_LAYOUT_NCHW)
        __stcs(faces + ((b * 6 + f) * C + c) * (int64_t)ww + pix, v);
      else
        faces[((b * 6 + f) * (int64_t)ww + pix) * C + c] = v;
    }
  }
}

// uint8, any channel count / alignment (byte loads)
template <int LAYOUT, class GEOM>
__global__ void __launch_bounds__(kE2cTileX * kE2cTileY)
e2c_kernel_u8_anyc(const uint8_t* __restrict__ frames, const uint32_t* __restrict__ packed,
                   float* __restrict__ faces, int64_t B, int Hin, int Win, int C, int w, float denom,
                   const __grid_constant__ GEOM geom) {
  __shared__ float lut[256];
  const int tid = threadIdx.y * kE2cTileX + threadIdx.x;
  pdl_trigger();
  lut[tid] = __fdiv_rn((float)tid, denom);
  __syncthreads();
  pdl_wait();
  E2cPix q;
  if (!e2c_locate(geom, w, &q)) return;
  const int f = q.f, ww = q.out_plane, pix = q.out_pix;     // output plane size / pixel (padded or not)
  const MapEntry me = load_map(packed, q.map_idx, Hin, Win);
  const int x0 = me.x0, y0 = me.y0;
  const float fx = me.fx, fy = me.fy;
  const float w00 = __fmul_rn(1.0f - fy, 1.0f - fx), w01 = __fmul_rn(1.0f - fy, fx);
  const float w10 = __fmul_rn(fy, 1.0f - fx), w11 = __fmul_rn(fy, fx);
  const bool x1_ok = x0 + 1 < Win, y1_ok = y0 + 1 < Hin;
  for (int64_t b = blockIdx.z; b < B; b += gridDim.z) {
    const uint8_t* row0 = frames + ((b * Hin + y0) * (int64_t)Win + x0) * C;
    const uint8_t* row1 = row0 + (int64_t)Win * C;
    for (int c = 0; c < C; ++c) {
      const float a00 = lut[__ldg(row0 + c)];
      const float a01 = x1_ok ? lut[__ldg(row0 + C + c)] : 0.0f;
      const float a10 = y1_ok ? lut[__ldg(row1 + c)] : 0.0f;
      const float a11 = (x1_ok && y1_ok) ? lut[__ldg(row1 + C + c)] : 0.0f;
      float v = __fmul_rn(a00, w00);
      v = __fadd_rn(v, __fmul_rn(a01, w01));
      v = __fadd_rn(v, __fmul_rn(a10, w10));
      v = __fadd_rn(v, __fmul_rn(a11, w11));
      if (LAYOUT == CP360_LAYOUT_NCHW)
        faces[((b * 6 + f) * C + c) * (int64_t)ww + pix] = v;
      else
        faces[((b * 6 + f) * (int64_t)ww + pix) * C + c] = v;
    }
  }
}

// ---- dispatch -----------------------------------------------------------------------------------
struct E2cArgs {
  const void* frames;
  bool u8;
  const uint32_t* packed;
  float* out;
  int64_t B;
  int Hin, Win, C, w, layout;
  float denom;
  bool norm;
  NormParams nrm;
  cudaStream_t st;
};

template <class GEOM>
static int e2c_dispatch(const E2cArgs& a, const GEOM& geom, int out_h, int out_w) {
  const int tiles_y = (out_h + kE2cTileY - 1) / kE2cTileY;
  dim3 block(kE2cTileX, kE2cTileY);
  dim3 grid((out_w + kE2cTileX - 1) / kE2cTileX, 6 * tiles_y, (unsigned)std::min<int64_t>(a.B, 65535));
  const bool nchw = a.layout == CP360_LAYOUT_NCHW;
  const int64_t B = a.B;
  const int Hin = a.Hin, Win = a.Win, C = a.C, w = a.w;
  cudaStream_t st = a.st;
#define CP360_E2C_LN(KERN, ...)                                                                               \
  do {                                                                                                        \
    if (nchw) {                                                                                               \
      if (a.norm) launch_kernel(KERN<CP360_LAYOUT_NCHW, true, GEOM>, grid, block, 0, st, __VA_ARGS__, geom);  \
      else launch_kernel(KERN<CP360_LAYOUT_NCHW, false, GEOM>, grid, block, 0, st, __VA_ARGS__, geom);        \
    } else {                                                                                                  \
      if (a.norm) launch_kernel(KERN<CP360_LAYOUT_NHWC, true, GEOM>, grid, block, 0, st, __VA_ARGS__, geom);  \
      else launch_kernel(KERN<CP360_LAYOUT_NHWC, false, GEOM>, grid, block, 0, st, __VA_ARGS__, geom);        \
    }                                                                                                         \
  } while (0)
#define CP360_E2C_L(KERN, ...)                                                                        \
  do {                                                                                                \
    if (nchw) launch_kernel(KERN<CP360_LAYOUT_NCHW, GEOM>, grid, block, 0, st, __VA_ARGS__, geom);    \
    else launch_kernel(KERN<CP360_LAYOUT_NHWC, GEOM>, grid, block, 0, st, __VA_ARGS__, geom);         \
  } while (0)
  if (a.u8) {
    const uint8_t* frames = static_cast<const uint8_t*>(a.frames);
    if (C == 3 && ((uintptr_t)frames % 4) == 0) {
      CP360_E2C_LN(e2c_kernel_u8c3, frames, a.packed, a.out, B, Hin, Win, w, a.denom, a.nrm);
    } else {
      CP360_CHECK_ARG(!a.norm, CP360_ERR_ALIGN, "fused normalisation needs 4 B-aligned uint8 frames");
      CP360_E2C_L(e2c_kernel_u8_anyc, frames, a.packed, a.out, B, Hin, Win, C, w, a.denom);
    }
    CP360_LAUNCHED();
    return CP360_OK;
  }
  const float* frames = static_cast<const float*>(a.frames);
  const bool vec_ok = C == 3 && ((uintptr_t)frames % 16) == 0 && ((int64_t)Win * C * 4) % 16 == 0 &&
                      ((int64_t)Hin * Win * C * 4) % 16 == 0;
  if (vec_ok) {
    grid.z = (unsigned)std::min<int64_t>((B + kE2cFramesPerThread - 1) / kE2cFramesPerThread, 65535);
    CP360_E2C_LN(e2c_kernel_c3v, frames, a.packed, a.out, B, Hin, Win, w, a.nrm);
    CP360_LAUNCHED();
    return CP360_OK;
  }
#define CP360_E2C_CASE(CC)                                                                                     \
  case CC:                                                                                                     \
    if (nchw) {                                                                                                \
      if (a.norm) launch_kernel(e2c_kernel<CC, CP360_LAYOUT_NCHW, true, GEOM>, grid, block, 0, st, frames, a.packed, a.out, B, Hin, Win, w, a.nrm, geom);  \
      else launch_kernel(e2c_kernel<CC, CP360_LAYOUT_NCHW, false, GEOM>, grid, block, 0, st, frames, a.packed, a.out, B, Hin, Win, w, a.nrm, geom);        \
    } else {                                                                                                   \
      if (a.norm) launch_kernel(e2c_kernel<CC, CP360_LAYOUT_NHWC, true, GEOM>, grid, block, 0, st, frames, a.packed, a.out, B, Hin, Win, w, a.nrm, geom);  \
      else launch_kernel(e2c_kernel<CC, CP360_LAYOUT_NHWC, false, GEOM>, grid, block, 0, st, frames, a.packed, a.out, B, Hin, Win, w, a.nrm, geom);        \
    }                                                                                                          \
    break;
  switch (C) {
    CP360_E2C_CASE(1)
    CP360_E2C_CASE(3)
    CP360_E2C_CASE(4)
    default:
      CP360_E2C_L(e2c_kernel_anyc, frames, a.packed, a.out, B, Hin, Win, C, w);
  }
#undef CP360_E2C_CASE
#undef CP360_E2C_L
#undef CP360_E2C_LN
  CP360_LAUNCHED();
  return CP360_OK;
}

// validation shared by the three entry points; *run = false when there is nothing to launch
static int e2c_prepare(E2cArgs* a, const float* mean_host, const float* std_host, bool* run) {
  *run = false;
  CP360_CHECK_ARG(a->B >= 0 && a->C >= 0 && a->w > 0 && a->Hin > 0 && a->Win > 0, CP360_ERR_BAD_ARG, "bad size");
  CP360_CHECK_ARG(a->Hin * 2 == a->Win, CP360_ERR_SHAPE,
                  "input must be 2:1 equirectangular (got %dx%d)", a->Win, a->Hin);
  CP360_CHECK_ARG(a->Win <= 65535 && a->Hin <= 32767, CP360_ERR_RANGE, "frames larger than 65535 x 32767 are not supported");
  CP360_CHECK_ARG(a->layout == CP360_LAYOUT_NCHW || a->layout == CP360_LAYOUT_NHWC,
                  CP360_ERR_BAD_ARG, "unknown layout %d", a->layout);
  CP360_CHECK_ARG(!a->u8 || a->denom > 0.0f, CP360_ERR_BAD_ARG, "denom must be positive");
  a->norm = mean_host != nullptr || std_host != nullptr;
  if (a->u8)
    CP360_CHECK_ARG(!a->norm || (mean_host && std_host && a->C == 3), CP360_ERR_BAD_ARG,
                    "fused normalisation of uint8 frames needs mean and std and C == 3");
  else
    CP360_CHECK_ARG(!a->norm || (mean_host && std_host && a->C <= 4 && a->C != 2), CP360_ERR_BAD_ARG,
                    "fused normalisation needs mean and std and C in {1,3,4}");
  if (a->B == 0 || a->C == 0) return CP360_OK;
  CP360_CHECK_ARG(a->frames && a->packed && a->out, CP360_ERR_BAD_ARG, "null pointer");
  const bool wide = a->Win > 2047 || a->Hin > 1023;
  CP360_CHECK_ARG((a->u8 || ((uintptr_t)a->frames % 4) == 0) && ((uintptr_t)a->out % 4) == 0 &&
                      ((uintptr_t)a->packed % (wide ? 8 : 4)) == 0, CP360_ERR_ALIGN,
                  "pointer not aligned (frames / faces 4 B, map %d B)", wide ? 8 : 4);
  int rc = require_device();
  if (rc != CP360_OK) return rc;
  a->nrm = NormParams{};
  if (a->norm)
    for (int c = 0; c < a->C; ++c) { a->nrm.mean[c] = mean_host[c]; a->nrm.std[c] = std_host[c]; }
  *run = true;
  return CP360_OK;
}

}  // namespace cp360

using namespace cp360;

extern "C" int cp360_e2c_fwd(const float* frames, const uint32_t* packed, float* faces, int64_t B,
                             int Hin, int Win, int C, int w, int out_layout,
                             const float* mean_host, const float* std_host, void* stream) {
  E2cArgs a = {frames, false, packed, faces, B, Hin, Win, C, w, out_layout, 1.0f, false, {}, (cudaStream_t)stream};
  bool run;
  int rc = e2c_prepare(&a, mean_host, std_host, &run);
  if (rc != CP360_OK || !run) return rc;
  return e2c_dispatch(a, NoPad{}, w, w);
}

extern "C" int cp360_e2c_fwd_u8(const uint8_t* frames, const uint32_t* packed, float* faces, int64_t B,
                                int Hin, int Win, int C, int w, int out_layout, float denom,
                                const float* mean_host, const float* std_host, void* stream) {
  E2cArgs a = {frames, true, packed, faces, B, Hin, Win, C, w, out_layout, denom, false, {}, (cudaStream_t)stream};
  bool run;
  int rc = e2c_prepare(&a, mean_host, std_host, &run);
  if (rc != CP360_OK || !run) return rc;
  return e2c_dispatch(a, NoPad{}, w, w);
}

extern "C" int cp360_e2c_cubepad_fwd(const void* frames, int frames_u8, const uint32_t* packed, float* padded,
                                     int64_t B, int Hin, int Win, int C, int w, int pl, int pr, int pt,
                                     int pd, float denom, const float* mean_host, const float* std_host,
                                     void* stream) {
  E2cArgs a = {frames, frames_u8 != 0, packed, padded, B, Hin, Win, C, w, CP360_LAYOUT_NCHW,
               frames_u8 ? denom : 1.0f, false, {}, (cudaStream_t)stream};
  WithPad geom;
  CP360_CHECK_ARG(w > 0 && make_geom(w, w, pl, pr, pt, pd, &geom.g), CP360_ERR_SHAPE,
                  "CubePad needs 0 <= pad <= face width (w=%d, pads l%d r%d t%d d%d)", w, pl, pr, pt, pd);
  bool run;
  int rc = e2c_prepare(&a, mean_host, std_host, &run);
  if (rc != CP360_OK || !run) return rc;
  return e2c_dispatch(a, geom, geom.g.Ho, geom.g.Wo);
}

#ifdef CP360_TRACE
extern "C" __attribute__((visibility("default"))) int cp360_trace_bind_e2c(void* rec, unsigned cap, void* n) {
  cp360::TraceBuf tb = {(cp360::TraceRec*)rec, cap, (unsigned*)n};
  return cudaMemcpyToSymbol(cp360::g_tb, &tb, sizeof(tb)) == cudaSuccess ? 0 : CP360_ERR_CUDA;
}
#endif
