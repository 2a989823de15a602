// Host map builders (double precision, built once per resolution).
//
//   cp360_e2c_build_map   replaces Equi2Cube.__init__        utils/equi_to_cube.py:12-110
//   cp360_c2e_build_map   replaces Cube2Equi.__init__        utils/cube_to_equi.py:12-35
//   cp360_c2e_build_plan  the fp32 coordinate arithmetic of  utils/cube_to_equi.py:58-64 plus
//                         torch grid_sample's unnormalise/floor/weights
//
// The integer maps these produce must equal the reference's bit for bit, so the float64
// operation ORDER of the numpy code is followed literally (including its table-lookup inverse
// trigonometry and the 1-based coordinates); compile with -ffp-contract=off.
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "../../include/cp360.h"

namespace cp360 {
void set_error(const char* fmt, ...);
}
using cp360::set_error;

namespace {

struct Mat3 { double m[3][3]; };

Mat3 matmul(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = a.m[i][0] * b.m[0][j];
      s += a.m[i][1] * b.m[1][j];
      s += a.m[i][2] * b.m[2][j];
      r.m[i][j] = s;
    }
  return r;
}

// sph_utils.py:23-38
Mat3 rot_x(double a) { double c = cos(a), s = sin(a); return {{{1, 0, 0}, {0, c, -s}, {0, s, c}}}; }
Mat3 rot_y(double a) { double c = cos(a), s = sin(a); return {{{c, 0, s}, {0, 1, 0}, {-s, 0, c}}}; }
Mat3 rot_z(double a) { double c = cos(a), s = sin(a); return {{{c, -s, 0}, {s, c, 0}, {0, 0, 1}}}; }

// numpy.interp(x, xp, fp) with fp[j] = j (what scipy interp1d(kind='linear') evaluates for the
// reference's lookup tables). Returns false if x is outside [xp[0], xp[n-1]] (interp1d raises).
bool interp_index(const std::vector<double>& xp, double x, double* out) {
  const int n = (int)xp.size();
  if (!(x >= xp[0] && x <= xp[n - 1])) return false;
  // largest j with xp[j] <= x
  int j = (int)(std::upper_bound(xp.begin(), xp.end(), x) - xp.begin()) - 1;
  if (j >= n - 1) { *out = (double)(n - 1); return true; }
  if (xp[j] == x) { *out = (double)j; return true; }
  const double slope = ((double)(j + 1) - (double)j) / (xp[j + 1] - xp[j]);
  *out = slope * (x - xp[j]) + (double)j;
  return true;
}

inline int32_t cv_round_f32_times32(double v) {
  const float f = (float)v;          // equi_to_cube.py:122-125 astype('float32')
  const float s = f * 32.0f;         // cv2: sX[x]*INTER_TAB_SIZE in float
  return (int32_t)lrintf(s);         // cvRound: round half to even (default FP environment)
}

}  // namespace

extern "C" {

int64_t cp360_e2c_map_words(int w, int Hin, int Win) {
  if (w <= 0) return 0;
  return (int64_t)6 * w * w * ((Win > 2047 || Hin > 1023) ? 2 : 1);
}

int cp360_e2c_build_map(int w, int Hin, int Win, double vfov_deg, uint32_t* packed_host,
                        int32_t* sx_host, int32_t* sy_host, double* inx_host, double* iny_host) {
  if (w <= 0 || Hin <= 1 || Win <= 1) { set_error("e2c map: non-positive size"); return CP360_ERR_BAD_ARG; }
  if (Hin * 2 != Win) {   // equi_to_cube.py:15
    set_error("e2c map: input must be 2:1 equirectangular (got %dx%d)", Win, Hin);
    return CP360_ERR_SHAPE;
  }
  if (Win > 65535 || Hin > 32767) {
    set_error("e2c map: frames larger than 65535 x 32767 are not supported (got %dx%d)", Win, Hin);
    return CP360_ERR_RANGE;
  }
  const bool wide = Win > 2047 || Hin > 1023;          // two words per pixel (cp360_e2c_map_words)
  const double pi = M_PI;
  const double vfov = vfov_deg * pi / 180;
  const int views_deg[6][3] = {{180, 0, 0}, {0, -90, 0}, {0, 0, 0}, {-90, 0, 0}, {90, 0, 0}, {0, 90, 0}};
  const double t = tan(vfov / 2);
  const double tl0 = -t * ((double)w / (double)w), tl1 = -t, tl2 = 1;
  const double uv0 = -2 * tl0 / w, uv1 = -2 * tl1 / w, uv2 = 0;

  const int res_acos = 2 * Win, res_atan = 2 * Hin;
  const double step_acos = pi / res_acos, step_atan = pi / res_atan;
  std::vector<double> lut_acos(res_acos + 1), lut_atan(res_atan + 1);
  for (int k = 0; k < res_acos; ++k) lut_acos[k] = -cos((double)k * step_acos);
  lut_acos[res_acos] = 1.0;
  lut_atan[0] = tan(step_atan / 2 - pi / 2);
  for (int k = 1; k < res_atan; ++k) lut_atan[k] = tan((double)k * step_atan - pi / 2);
  lut_atan[res_atan] = tan(-step_atan / 2 + pi / 2);

  const double half_w = Win / 2.0, half_h = Hin / 2.0;
  for (int f = 0; f < 6; ++f) {
    const double yaw = views_deg[f][0] * pi / 180, pitch = views_deg[f][1] * pi / 180,
                 roll = views_deg[f][2] * pi / 180;
    const Mat3 tf = matmul(matmul(rot_y(yaw), rot_x(pitch)), rot_z(roll));
    for (int Y = 0; Y < w; ++Y)
      for (int X = 0; X < w; ++X) {
        const double px = tl0 + uv0 * X, py = tl1 + uv1 * Y, pz = tl2 + uv2 * 1.0;
        double mv[3];
        for (int i = 0; i < 3; ++i) {
          double s = tf.m[i][0] * px;
          s += tf.m[i][1] * py;
          s += tf.m[i][2] * pz;
          mv[i] = s;
        }
        const double xp = mv[0], yp = mv[1], zp = mv[2];
        const double nxz = sqrt(xp * xp + zp * zp);
        double phi = 0, theta = 0;
        if (nxz < 10e-10) {
          phi = yp > 0 ? pi / 2 : -pi / 2;
        } else {
          double ia, ic;
          if (!interp_index(lut_atan, yp / nxz, &ia) || !interp_index(lut_acos, -zp / nxz, &ic)) {
            set_error("e2c map: A value in x_new is outside the interpolation range (vfov=%g)", vfov_deg);
            return CP360_ERR_RANGE;
          }
          phi = ia * step_atan - (pi / 2);
          theta = ic * step_acos;
          if (xp < 0) theta = -theta;
        }
        double in_x = (theta / pi) * half_w + half_w + 1;
        double in_y = (phi / (pi / 2)) * half_h + half_h + 1;
        if (in_x < 1) in_x = 1;
        if (in_x >= Win - 1) in_x = Win - 1;
        if (in_y < 1) in_y = 1;
        if (in_y >= Hin - 1) in_y = Hin - 1;
        const size_t o = ((size_t)f * w + Y) * w + X;
        if (inx_host) inx_host[o] = in_x;
        if (iny_host) iny_host[o] = in_y;
        const int32_t sx = cv_round_f32_times32(in_x), sy = cv_round_f32_times32(in_y);
        if (sx_host) sx_host[o] = sx;
        if (sy_host) sy_host[o] = sy;
        if (packed_host) {
          const uint32_t x0 = (uint32_t)(sx >> 5), y0 = (uint32_t)(sy >> 5);
          if (wide) {
            packed_host[2 * o] = (x0 << 16) | y0;
            packed_host[2 * o + 1] = ((uint32_t)(sx & 31) << 5) | (uint32_t)(sy & 31);
          } else {
            packed_host[o] = (x0 << 20) | (y0 << 10) | ((uint32_t)(sx & 31) << 5) | (uint32_t)(sy & 31);
          }
        }
      }
  }
  return CP360_OK;
}

int cp360_c2e_build_map(int w, int8_t* face_host, double* coord_host) {
  if (w <= 0) { set_error("c2e map: non-positive face width"); return CP360_ERR_BAD_ARG; }
  if (w > 8191) { set_error("c2e map: face width > 8191"); return CP360_ERR_RANGE; }
  const int out_w = 4 * w, out_h = 2 * w;
  const double pi = M_PI, err = 10e-9, eps = 10e-9;
  auto prune = [&](double a) {     // sph_utils.py:70-77
    if (a == 0.0) return err;
    if (a == pi) return pi - err;
    if (a == -pi) return -pi + err;
    if (a == pi / 2) return pi / 2 - err;
    if (a == -pi / 2) return -pi / 2 + err;
    return a;
  };
  for (int Y = 0; Y < out_h; ++Y)
    for (int X = 0; X < out_w; ++X) {
      // xy2angle, sph_utils.py:53-60
      const double xx = 2 * (X + 0.5) / (double)out_w - 1;
      const double yy = 1 - 2 * (Y + 0.5) / (double)out_h;
      const double theta = prune(xx * pi), phi = prune(yy * pi / 2);
      // to_3dsphere, sph_utils.py:63-67
      const double x = 1 * cos(phi) * cos(theta), y = 1 * sin(phi), z = 1 * cos(phi) * sin(theta);
      // get_face, sph_utils.py:88-111 — np.maximum(|x|,|y|,out=|z|): the "max" ignores z
      const double m = std::max(fabs(x), fabs(y));
      const bool xf = m - fabs(x) < eps, yf = m - fabs(y) < eps, zf = m - fabs(z) < eps;
      int face = 0;
      if (x >= 0 && xf) face = 2;   // F
      if (x <= 0 && xf) face = 0;   // B
      if (y >= 0 && yf) face = 5;   // T
      if (y <= 0 && yf) face = 1;   // D
      if (z >= 0 && zf) face = 4;   // R
      if (z <= 0 && zf) face = 3;   // L
      // face_to_cube_coord, sph_utils.py:114-146
      double d0, d1, d2;
      switch (face) {
        case 2: d0 = z; d1 = y; d2 = x; break;
        case 0: d0 = -z; d1 = y; d2 = x; break;
        case 5: d0 = z; d1 = -x; d2 = y; break;
        case 1: d0 = z; d1 = x; d2 = y; break;
        case 4: d0 = -x; d1 = y; d2 = z; break;
        default: d0 = x; d1 = y; d2 = z; break;
      }
      const double x_on = (d0 / fabs(d2) + 1) / 2, y_on = (-d1 / fabs(d2) + 1) / 2;
      // norm_to_cube, sph_utils.py:149-153
      double cx = x_on * (w - 1), cy = y_on * (w - 1);
      if (cx < 0.) cx = 0.;
      if (cx > (w - 1)) cx = (w - 1);
      if (cy < 0.) cy = 0.;
      if (cy > (w - 1)) cy = (w - 1);
      const size_t o = (size_t)Y * out_w + X;
      if (face_host) face_host[o] = (int8_t)face;
      if (coord_host) { coord_host[2 * o] = cx; coord_host[2 * o + 1] = cy; }
    }
  return CP360_OK;
}

int cp360_c2e_build_plan(int w, int align_corners, uint32_t* tap_host, float* wts_host,
                         float* M_out) {
  if (w <= 0) { set_error("c2e plan: non-positive face width"); return CP360_ERR_BAD_ARG; }
  if (w > 8191) { set_error("c2e plan: face width > 8191"); return CP360_ERR_RANGE; }
  const size_t P = (size_t)8 * w * w;
  std::vector<int8_t> face(P);
  std::vector<double> coord(2 * P);
  int rc = cp360_c2e_build_map(w, face.data(), coord.data());
  if (rc != CP360_OK) return rc;
  float M = -INFINITY;                                   // cube_to_equi.py:58 torch.max(gridf)
  for (size_t i = 0; i < 2 * P; ++i) M = std::max(M, (float)coord[i]);
  if (M_out) *M_out = M;
  const float half = M / 2.0f;
  const float fw = (float)w, fwm1 = (float)(w - 1);
  for (size_t i = 0; i < P; ++i) {
    float pix[2];
    for (int k = 0; k < 2; ++k) {
      const float g = (float)coord[2 * i + k];
      const float gn = (g - half) / half;
      // grid_sampler_unnormalize, ATen/native/GridSampler.h:27-36
      pix[k] = align_corners ? ((gn + 1.0f) / 2.0f) * fwm1 : ((gn + 1.0f) * fw - 1.0f) / 2.0f;
    }
    const float ix = pix[0], iy = pix[1];
    const float x_w = floorf(ix), y_n = floorf(iy);
    const float x_e = x_w + 1.0f, y_s = y_n + 1.0f;
    if (wts_host) {
      wts_host[4 * i + 0] = (x_e - ix) * (y_s - iy);   // nw
      wts_host[4 * i + 1] = (ix - x_w) * (y_s - iy);   // ne
      wts_host[4 * i + 2] = (x_e - ix) * (iy - y_n);   // sw
      wts_host[4 * i + 3] = (ix - x_w) * (iy - y_n);   // se
    }
    if (tap_host) {
      int x0 = (int)x_w, y0 = (int)y_n;
      // taps further out than one pixel never contribute; clamp so the packed form holds them
      x0 = std::min(std::max(x0, -1), w);
      y0 = std::min(std::max(y0, -1), w);
      tap_host[i] = ((uint32_t)face[i] << 28) | ((uint32_t)(y0 + 1) << 14) | (uint32_t)(x0 + 1);
    }
  }
  return CP360_OK;
}

int cp360_c2e_build_bwd_plan(int w, int align_corners, int32_t* offsets_host, int32_t* pix_host, float* wts_host) {
  if (w <= 0 || !offsets_host) { set_error("c2e backward plan: bad argument"); return CP360_ERR_BAD_ARG; }
  if (w > 8191) { set_error("c2e backward plan: face width > 8191"); return CP360_ERR_RANGE; }
  const size_t P = (size_t)8 * w * w, NC = (size_t)6 * w * w;
  std::vector<uint32_t> tap(P);
  std::vector<float> wts(4 * P);
  int rc = cp360_c2e_build_plan(w, align_corners, tap.data(), wts.data(), nullptr);
  if (rc != CP360_OK) return rc;
  // the four taps of output pixel i, in the order the forward kernel accumulates them (nw, ne, sw, se)
  auto for_each_tap = [&](size_t i, auto&& fn) {
    const int face = (int)(tap[i] >> 28), y0 = (int)((tap[i] >> 14) & 0x3fffu) - 1, x0 = (int)(tap[i] & 0x3fffu) - 1;
    const int dy[4] = {0, 0, 1, 1}, dx[4] = {0, 1, 0, 1};
    for (int k = 0; k < 4; ++k) {
      const int yy = y0 + dy[k], xx = x0 + dx[k];
      if (yy < 0 || yy >= w || xx < 0 || xx >= w) continue;      // padding_mode='zeros': no gradient
      fn(((size_t)face * w + yy) * w + xx, wts[4 * i + k]);
    }
  };
  std::vector<int32_t> count(NC + 1, 0);
  for (size_t i = 0; i < P; ++i) for_each_tap(i, [&](size_t cell, float) { ++count[cell + 1]; });
  for (size_t c = 0; c < NC; ++c) count[c + 1] += count[c];
  for (size_t c = 0; c <= NC; ++c) offsets_host[c] = count[c];
  if (pix_host && wts_host) {
    std::vector<int32_t> fill(count.begin(), count.end() - 1);
    for (size_t i = 0; i < P; ++i)                               // increasing output pixel: the fixed summation order
      for_each_tap(i, [&](size_t cell, float wt) {
        pix_host[fill[cell]] = (int32_t)i;
        wts_host[fill[cell]] = wt;
        ++fill[cell];
      });
  }
  return CP360_OK;
}

int cp360_c2e_build_cubic_plan(int w, uint32_t* tap_host) {
  if (w <= 0 || !tap_host) { set_error("c2e cubic plan: bad argument"); return CP360_ERR_BAD_ARG; }
  if (w > 512) { set_error("c2e cubic plan: face width > 512"); return CP360_ERR_RANGE; }
  const size_t P = (size_t)8 * w * w;
  std::vector<int8_t> face(P);
  std::vector<double> coord(2 * P);
  int rc = cp360_c2e_build_map(w, face.data(), coord.data());
  if (rc != CP360_OK) return rc;
  for (size_t i = 0; i < P; ++i) {
    // cube_to_equi.py:83 gridf.astype(np.float32), then cv2's float map -> 1/32-pixel fixed point
    const int32_t sx = cv_round_f32_times32(coord[2 * i]), sy = cv_round_f32_times32(coord[2 * i + 1]);
    const uint32_t x0p1 = (uint32_t)(sx >> 5), y0p1 = (uint32_t)(sy >> 5);   // window origin + 1, in [0, w-1]
    tap_host[i] = ((uint32_t)face[i] << 28) | ((uint32_t)(sy & 31) << 23) | ((uint32_t)(sx & 31) << 18) |
                  (y0p1 << 9) | x0p1;
  }
  return CP360_OK;
}

}  // extern "C"
