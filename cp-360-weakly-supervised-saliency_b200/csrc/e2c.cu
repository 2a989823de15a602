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

namespace cp360 {

struct NormParams { float mean[4]; float std[4]; };

constexpr int kE2cTileX = 32, kE2cTileY = 8;

template <int C>
__device__ __forceinline__ void load_px(const float* __restrict__ p, float (&v)[C]) {
#pragma unroll
  for (int c = 0; c < C; ++c) v[c] = __ldg(p + c);
}

template <int C, int LAYOUT, bool NORM>
__global__ void __launch_bounds__(kE2cTileX * kE2cTileY)
e2c_kernel(const float* __restrict__ frames, const uint32_t* __restrict__ packed,
           float* __restrict__ faces, int64_t B, int Hin, int Win, int w, NormParams nrm) {
  const int ox = blockIdx.x * kE2cTileX + threadIdx.x;
  const int tiles_y = (w + kE2cTileY - 1) / kE2cTileY;
  const int f = blockIdx.y / tiles_y;
  const int oy = (blockIdx.y - f * tiles_y) * kE2cTileY + threadIdx.y;
  if (ox >= w || oy >= w) return;
  const int ww = w * w;
  const uint32_t p = __ldg(packed + (size_t)f * ww + oy * w + ox);
  const int x0 = (int)(p >> 20), y0 = (int)((p >> 10) & 1023u);
  const float fx = (float)((p >> 5) & 31u) * 0.03125f, fy = (float)(p & 31u) * 0.03125f;
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
        __stcs(faces + ((b * 6 + f) * C + c) * (int64_t)ww + oy * w + ox, v);
      else
        faces[((b * 6 + f) * (int64_t)ww + oy * w + ox) * C + c] = v;
    }
  }
}

// any channel count (no normalisation)
template <int LAYOUT>
__global__ void __launch_bounds__(kE2cTileX * kE2cTileY)
e2c_kernel_anyc(const float* __restrict__ frames, const uint32_t* __restrict__ packed,
                float* __restrict__ faces, int64_t B, int Hin, int Win, int C, int w) {
  const int ox = blockIdx.x * kE2cTileX + threadIdx.x;
  const int tiles_y = (w + kE2cTileY - 1) / kE2cTileY;
  const int f = blockIdx.y / tiles_y;
  const int oy = (blockIdx.y - f * tiles_y) * kE2cTileY + threadIdx.y;
  if (ox >= w || oy >= w) return;
  const int ww = w * w;
  const uint32_t p = __ldg(packed + (size_t)f * ww + oy * w + ox);
  const int x0 = (int)(p >> 20), y0 = (int)((p >> 10) & 1023u);
  const float fx = (float)((p >> 5) & 31u) * 0.03125f, fy = (float)(p & 31u) * 0.03125f;
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
        faces[((b * 6 + f) * C + c) * (int64_t)ww + oy * w + ox] = v;
      else
        faces[((b * 6 + f) * (int64_t)ww + oy * w + ox) * C + c] = v;
    }
  }
}

template <int C, int LAYOUT>
static void launch_c(bool norm, dim3 grid, dim3 block, cudaStream_t st, const float* frames,
                     const uint32_t* packed, float* faces, int64_t B, int Hin, int Win, int w,
                     const NormParams& nrm) {
  if (norm)
    e2c_kernel<C, LAYOUT, true><<<grid, block, 0, st>>>(frames, packed, faces, B, Hin, Win, w, nrm);
  else
    e2c_kernel<C, LAYOUT, false><<<grid, block, 0, st>>>(frames, packed, faces, B, Hin, Win, w, nrm);
}

}  // namespace cp360

using namespace cp360;

extern "C" int cp360_e2c_fwd(const float* frames, const uint32_t* packed, float* faces, int64_t B,
                             int Hin, int Win, int C, int w, int out_layout,
                             const float* mean_host, const float* std_host, void* stream) {
  CP360_CHECK_ARG(B >= 0 && C >= 0 && w > 0 && Hin > 0 && Win > 0, CP360_ERR_BAD_ARG, "bad size");
  CP360_CHECK_ARG(Hin * 2 == Win, CP360_ERR_SHAPE,
                  "input must be 2:1 equirectangular (got %dx%d)", Win, Hin);
  CP360_CHECK_ARG(Win <= 2047 && Hin <= 1023, CP360_ERR_RANGE, "packed map supports up to 2047x1023");
  CP360_CHECK_ARG(out_layout == CP360_LAYOUT_NCHW || out_layout == CP360_LAYOUT_NHWC,
                  CP360_ERR_BAD_ARG, "unknown layout %d", out_layout);
  const bool norm = mean_host != nullptr || std_host != nullptr;
  CP360_CHECK_ARG(!norm || (mean_host && std_host && C <= 4 && C != 2), CP360_ERR_BAD_ARG,
                  "fused normalisation needs mean and std and C in {1,3,4}");
  if (B == 0 || C == 0) return CP360_OK;
  CP360_CHECK_ARG(frames && packed && faces, CP360_ERR_BAD_ARG, "null pointer");
  CP360_CHECK_ARG(((uintptr_t)frames % 4) == 0 && ((uintptr_t)faces % 4) == 0 &&
                      ((uintptr_t)packed % 4) == 0, CP360_ERR_ALIGN, "pointer not 4 B aligned");
  int rc = require_device();
  if (rc != CP360_OK) return rc;
  NormParams nrm = {};
  if (norm)
    for (int c = 0; c < C; ++c) { nrm.mean[c] = mean_host[c]; nrm.std[c] = std_host[c]; }
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_y = (w + kE2cTileY - 1) / kE2cTileY;
  dim3 block(kE2cTileX, kE2cTileY);
  dim3 grid((w + kE2cTileX - 1) / kE2cTileX, 6 * tiles_y, (unsigned)std::min<int64_t>(B, 65535));
  const bool nchw = out_layout == CP360_LAYOUT_NCHW;
#define CP360_E2C_CASE(CC)                                                                      \
  case CC:                                                                                      \
    if (nchw) launch_c<CC, CP360_LAYOUT_NCHW>(norm, grid, block, st, frames, packed, faces, B,  \
                                               Hin, Win, w, nrm);                               \
    else launch_c<CC, CP360_LAYOUT_NHWC>(norm, grid, block, st, frames, packed, faces, B, Hin,  \
                                          Win, w, nrm);                                         \
    break;
  switch (C) {
    CP360_E2C_CASE(1)
    CP360_E2C_CASE(3)
    CP360_E2C_CASE(4)
    default:
      if (nchw)
        e2c_kernel_anyc<CP360_LAYOUT_NCHW><<<grid, block, 0, st>>>(frames, packed, faces, B, Hin, Win, C, w);
      else
        e2c_kernel_anyc<CP360_LAYOUT_NHWC><<<grid, block, 0, st>>>(frames, packed, faces, B, Hin, Win, C, w);
  }
#undef CP360_E2C_CASE
  CP360_LAUNCHED();
  return CP360_OK;
}
