// Epipolar inverse-depth search + measurement model + fusion for sm_100a: one thread per feature.
//
// Replaces the loop body of Flame::updateFeatureIDepths / Flame::trackFeature
// (src/flame/flame.cc:1280-1752) and everything it calls:
//   stereo::EpipolarGeometry<float>            src/flame/stereo/epipolar_geometry.h:44-428
//   stereo::inverse_depth_filter::{predict, getSearchRegion, search, update}
//                                              src/flame/stereo/inverse_depth_filter.cc:35-305
//   stereo::line_stereo::match                 src/flame/stereo/line_stereo.h:75-385
//   stereo::InverseDepthMeasModel::idepth      src/flame/stereo/inverse_depth_meas_model.cc:49-156
//   utils::bilinearInterp / bilinearWeights    src/flame/utils/image_utils.h:199-272
//   utils::clipLineLiangBarsky                 src/flame/utils/image_utils.cc:269-371
//
// This translation unit is compiled with -fmad=false: every expression below is evaluated with
// separately rounded fp32 operations in the association order of the reference's C++ (its
// x86-64 Release build has no FMA); '/' and sqrt are IEEE-correct (-prec-div/-prec-sqrt
// defaults). The matcher is SSD, not NCC; the gradient tap uses offset 3 on the
// 5-px-padded gradient images.
//
// The walk along the epipolar line is latency-bound (<= ~36 dependent bilinear taps per feature on
// L2-resident u8 images), not bandwidth-bound; see DESIGN.md.
#include <cfloat>
#include <cstring>

#include "internal.h"

namespace fb {

struct FrameDev {
  const unsigned char* img;  // (H + 2 pad) x (W + 2 pad) u8
  const float* gx;           // same size f32
  const float* gy;
  float pose[7];             // camera-to-world: qw qx qy qz tx ty tz
  int valid;
};

#define FB_MAX_FRAMES 64

struct StereoState {
  int width = 0, height = 0, pad = 0;
  float K[9], Kinv[9];
  FrameDev h_frames[FB_MAX_FRAMES];
  FrameDev* d_frames = nullptr;
  unsigned char* d_img[FB_MAX_FRAMES] = {nullptr};
  float* d_gx[FB_MAX_FRAMES] = {nullptr};
  float* d_gy[FB_MAX_FRAMES] = {nullptr};
  size_t cap_pix[FB_MAX_FRAMES] = {0};
  unsigned char* d_raw = nullptr;    // raw u8 image of the frame being prepared
  size_t cap_raw = 0;
  unsigned char* h_stage = nullptr;  // pinned
  size_t h_stage_cap = 0;
  cudaEvent_t stage_free = nullptr;
  flame_b200_feature* d_feats = nullptr;
  flame_b200_track_diag* d_diag = nullptr;
  int* d_count = nullptr;
  int cap_feats = 0;
  flame_b200_feature* h_feats = nullptr;  // pinned mirrors
  flame_b200_track_diag* h_diag = nullptr;
  int* h_count = nullptr;
};

struct StereoArgs {
  flame_b200_track_params P;
  float K[9], Kinv[9];
  int width, height, pad;
  int fnew, curr_pf;
  const FrameDev* frames;
  flame_b200_feature* feats;
  flame_b200_track_diag* diag;
  int n;
  int* count;
};

// ------------------------------------------------------------------------------------------
// small fp helpers (same as the reference's)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float fast_absf(float r) { return (r > 0) ? r : -r; }  // image_utils.h:91
__device__ __forceinline__ float sqrt_rn(float x) { return __fsqrt_rn(x); }  // == (float)sqrt((double)x)

// image_utils.h:199-214 + 230-255, ChannelType = uint8_t
__device__ __forceinline__ float bilinear_u8(const unsigned char* __restrict__ data, int step,
                                             float x, float y) {
  const int xf = (int)x, yf = (int)y;
  const float dx = x - xf, dy = y - yf;
  const float w11 = dx * dy;
  const float w01 = dx - w11;
  const float w10 = dy - w11;
  const float w00 = 1.0f - dx - dy + w11;
  const unsigned char* d = data + (size_t)yf * step + xf;
  return w00 * __ldg(d) + w01 * __ldg(d + 1) + w10 * __ldg(d + step) + w11 * __ldg(d + 1 + step);
}

// ChannelType = float (step in elements)
__device__ __forceinline__ float bilinear_f32(const float* __restrict__ data, int step, float x,
                                              float y) {
  const int xf = (int)x, yf = (int)y;
  const float dx = x - xf, dy = y - yf;
  const float w11 = dx * dy;
  const float w01 = dx - w11;
  const float w10 = dy - w11;
  const float w00 = 1.0f - dx - dy + w11;
  const float* d = data + (size_t)yf * step + xf;
  return w00 * __ldg(d) + w01 * __ldg(d + 1) + w10 * __ldg(d + step) + w11 * __ldg(d + 1 + step);
}

// image_utils.cc:269-371
__device__ bool clip_line(float xmin, float xmax, float ymin, float ymax, float x0, float y0,
                          float x1, float y1, float* x0c, float* y0c, float* x1c, float* y1c) {
  float t0 = 0.0f, t1 = 1.0f;
  const float xdelta = x1 - x0, ydelta = y1 - y0;
#pragma unroll
  for (int edge = 0; edge < 4; edge++) {
    float p, q;
    if (edge == 0) {
      p = -xdelta; q = -(xmin - x0);
    } else if (edge == 1) {
      p = xdelta; q = (xmax - x0);
    } else if (edge == 2) {
      p = -ydelta; q = -(ymin - y0);
    } else {
      p = ydelta; q = (ymax - y0);
    }
    const float r = q / p;
    if (p == 0 && q < 0) return false;
    if (p < 0) {
      if (r > t1) return false;
      else if (r > t0) t0 = r;
    } else if (p > 0) {
      if (r < t0) return false;
      else if (r < t1) t1 = r;
    }
  }
  float a0 = x0 + t0 * xdelta, b0 = y0 + t0 * ydelta;
  float a1 = x0 + t1 * xdelta, b1 = y0 + t1 * ydelta;
  if (a0 < xmin) a0 = xmin;
  if (a0 > xmax) a0 = xmax;
  if (b0 < ymin) b0 = ymin;
  if (b0 > ymax) b0 = ymax;
  if (a1 < xmin) a1 = xmin;
  if (a1 > xmax) a1 = xmax;
  if (b1 < ymin) b1 = ymin;
  if (b1 > ymax) b1 = ymax;
  *x0c = a0; *y0c = b0; *x1c = a1; *y1c = b1;
  return true;
}

// ------------------------------------------------------------------------------------------
// Eigen 3.2.10 / Sophus b474f05 arithmetic the reference relies on (third-party, restated; see
// DESIGN.md "Oracle" for what is and is not pinned)
// ------------------------------------------------------------------------------------------
__device__ void quat_to_rot(const float q[4], float R[9]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0f - (tyy + tzz); R[1] = txy - twz;          R[2] = txz + twy;
  R[3] = txy + twz;          R[4] = 1.0f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;          R[7] = tyz + twx;          R[8] = 1.0f - (txx + tyy);
}

__device__ void quat_rotate(const float q[4], const float v[3], float out[3]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  float uvx = y * v[2] - z * v[1];
  float uvy = z * v[0] - x * v[2];
  float uvz = x * v[1] - y * v[0];
  uvx += uvx; uvy += uvy; uvz += uvz;
  out[0] = (v[0] + w * uvx) + (y * uvz - z * uvy);
  out[1] = (v[1] + w * uvy) + (z * uvx - x * uvz);
  out[2] = (v[2] + w * uvz) + (x * uvy - y * uvx);
}

__device__ void quat_inverse(const float q[4], float out[4]) {
  const float n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0];
  if (n2 > 0) {
    out[0] = q[0] / n2; out[1] = -q[1] / n2; out[2] = -q[2] / n2; out[3] = -q[3] / n2;
  } else {
    out[0] = out[1] = out[2] = out[3] = 0.0f;
  }
}

__device__ void quat_mul(const float a[4], const float b[4], float out[4]) {
  out[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  out[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  out[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  out[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

__device__ void mat3_mul(const float A[9], const float B[9], float C[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float r = A[3 * i] * B[j];
      r += A[3 * i + 1] * B[3 + j];
      r += A[3 * i + 2] * B[6 + j];
      C[3 * i + j] = r;
    }
}

__device__ void mat3_vec(const float A[9], const float v[3], float out[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float r = A[3 * i] * v[0];
    r += A[3 * i + 1] * v[1];
    r += A[3 * i + 2] * v[2];
    out[i] = r;
  }
}

// out = a^-1 * b for poses (q w,x,y,z ; t) — Sophus SE3 inverse/compose (flame.cc:1314)
__device__ void se3_inv_mul(const float a[7], const float b[7], float out[7]) {
  const float qi[4] = {a[0], -a[1], -a[2], -a[3]};
  const float nt[3] = {a[4] * -1.0f, a[5] * -1.0f, a[6] * -1.0f};
  float ti[3], rb[3], q[4];
  quat_rotate(qi, nt, ti);
  quat_rotate(qi, b + 4, rb);
  quat_mul(qi, b, q);
  const float n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0];
  if (n2 != 1.0f) {
    const float s = 2.0f / (1.0f + n2);
    q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
  }
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; out[3] = q[3];
  out[4] = ti[0] + rb[0]; out[5] = ti[1] + rb[1]; out[6] = ti[2] + rb[2];
}

// ------------------------------------------------------------------------------------------
// stereo::EpipolarGeometry<float>
// ------------------------------------------------------------------------------------------
struct EpiGeo {
  float K[9], Kinv[9];
  float q[4], t[3], t_cmp_to_ref[3];
  float KRKinv[9], KRKinv3[3], Kt[3], epipole[2];

  __device__ void init(const float* K_, const float* Kinv_) {
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      K[i] = K_[i];
      Kinv[i] = Kinv_[i];
    }
    epipole[0] = epipole[1] = 0.0f;
  }
  // :86-103
  __device__ void load(const float* q_, const float* t_) {
    for (int i = 0; i < 4; ++i) q[i] = q_[i];
    for (int i = 0; i < 3; ++i) t[i] = t_[i];
    float qi[4], r[3];
    quat_inverse(q, qi);
    quat_rotate(qi, t, r);
    t_cmp_to_ref[0] = -r[0]; t_cmp_to_ref[1] = -r[1]; t_cmp_to_ref[2] = -r[2];
    float R[9], KR[9];
    quat_to_rot(q, R);
    mat3_mul(K, R, KR);
    mat3_mul(KR, Kinv, KRKinv);
    KRKinv3[0] = KRKinv[6]; KRKinv3[1] = KRKinv[7]; KRKinv3[2] = KRKinv[8];
    mat3_vec(K, t, Kt);
    if (t[2] > 0) {
      epipole[0] = (K[0] * t[0] + K[2] * t[2]) / t[2];
      epipole[1] = (K[4] * t[1] + K[5] * t[2]) / t[2];
    }
  }
  // :191-201
  __device__ void max_depth_projection(float ux, float uy, float* ix, float* iy) const {
    const float h[3] = {ux, uy, 1.0f};
    float c[3];
    mat3_vec(KRKinv, h, c);
    const float inv = 1.0f / c[2];
    *ix = c[0] * inv;
    *iy = c[1] * inv;
  }
  // :127-143
  __device__ void project(float ux, float uy, float idepth, float* cx, float* cy) const {
    if (idepth == 0.0f) {
      max_depth_projection(ux, uy, cx, cy);
      return;
    }
    const float depth = 1.0f / idepth;
    const float h[3] = {ux * depth, uy * depth, depth};
    float c[3];
    mat3_vec(KRKinv, h, c);
    c[0] = c[0] + Kt[0]; c[1] = c[1] + Kt[1]; c[2] = c[2] + Kt[2];
    const float inv = 1.0f / c[2];
    *cx = c[0] * inv;
    *cy = c[1] * inv;
  }
  // :153-180
  __device__ void project_idepth(float ux, float uy, float idepth, float* cx, float* cy,
                                 float* new_idepth) const {
    if (idepth == 0.0f) {
      max_depth_projection(ux, uy, cx, cy);
      *new_idepth = 0.0f;
      return;
    }
    const float depth = 1.0f / idepth;
    float p_ref[3] = {Kinv[0] * ux + Kinv[2], Kinv[4] * uy + Kinv[5], 1.0f};
    p_ref[0] *= depth; p_ref[1] *= depth; p_ref[2] *= depth;
    float r[3];
    quat_rotate(q, p_ref, r);
    const float pc[3] = {r[0] + t[0], r[1] + t[1], r[2] + t[2]};
    const float u0 = K[0] * pc[0] + K[2] * pc[2];
    const float u1 = K[4] * pc[1] + K[5] * pc[2];
    *new_idepth = 1.0f / pc[2];
    *cx = u0 * (*new_idepth);
    *cy = u1 * (*new_idepth);
  }
  // :237-263
  __device__ void min_depth_projection(float ux, float uy, float* mx, float* my) const {
    if (t[2] > 0) {
      *mx = epipole[0];
      *my = epipole[1];
    } else if (t[2] == 0) {
      const float ex = K[0] * t[0], ey = K[4] * t[1];
      float ix, iy;
      max_depth_projection(ux, uy, &ix, &iy);
      *mx = ix + (float)(1e6 * (double)ex);
      *my = iy + (float)(1e6 * (double)ey);
    } else {
      const float qp[3] = {Kinv[0] * ux + Kinv[2], Kinv[4] * uy + Kinv[5], 1.0f};
      float r[3];
      quat_rotate(q, qp, r);
      const float min_depth = (1.0f - t[2]) / r[2];
      const float pc[3] = {min_depth * r[0] + t[0], min_depth * r[1] + t[1], min_depth * r[2] + t[2]};
      *mx = (K[0] * pc[0] + K[2] * pc[2]) / pc[2];
      *my = (K[4] * pc[1] + K[5] * pc[2]) / pc[2];
    }
  }
  // :282-300
  __device__ void epiline(float ux, float uy, float* ix, float* iy, float* ex, float* ey) const {
    float zx, zy;
    min_depth_projection(ux, uy, &zx, &zy);
    max_depth_projection(ux, uy, ix, iy);
    float epx = zx - *ix, epy = zy - *iy;
    const float norm2 = epx * epx + epy * epy;
    if ((double)norm2 > 1e-10) {
      // `1.0f / sqrt(norm2)` in the reference: ::sqrt(double) -> the quotient is formed in double
      const float inv_norm = (float)(1.0 / sqrt((double)norm2));
      epx *= inv_norm;
      epy *= inv_norm;
    } else {
      epx = 0.0f;
      epy = 0.0f;
    }
    *ex = epx;
    *ey = epy;
  }
  // :311-331
  __device__ void reference_epiline(float ux, float uy, float* ex, float* ey) const {
    const float rx = -K[0] * t_cmp_to_ref[0] + t_cmp_to_ref[2] * (ux - K[2]);
    const float ry = -K[4] * t_cmp_to_ref[1] + t_cmp_to_ref[2] * (uy - K[5]);
    const float n2 = rx * rx + ry * ry;
    const float inv = (float)(1.0 / sqrt((double)n2));  // :325, double quotient as above
    *ex = rx * inv;
    *ey = ry * inv;
  }
  // :392-407
  __device__ float disparity_to_idepth(float ux, float uy, float ix, float iy, float ex, float ey,
                                       float disp) const {
    const float w = KRKinv3[0] * ux + KRKinv3[1] * uy + KRKinv3[2];
    const float Ax = Kt[0] - Kt[2] * (ix + disp * ex);
    const float Ay = Kt[1] - Kt[2] * (iy + disp * ey);
    const float wd = w * disp;
    const float bx = ex * wd, by = ey * wd;
    const float ATA = Ax * Ax + Ay * Ay;
    const float ATb = Ax * bx + Ay * by;
    return ATb / ATA;
  }
};

// ------------------------------------------------------------------------------------------
// stereo::line_stereo::match (line_stereo.h:75-385)
// ------------------------------------------------------------------------------------------
__device__ int line_stereo_match(float rescale_factor, const float ref_patch[5],
                                 const unsigned char* __restrict__ img, int step, float sx, float sy,
                                 float ex, float ey, float* ucx, float* ucy, float* residual,
                                 const flame_b200_ls_params& params) {
  const float realVal_m2 = ref_patch[0], realVal_m1 = ref_patch[1], realVal = ref_patch[2];
  const float realVal_p1 = ref_patch[3], realVal_p2 = ref_patch[4];

  float incx = ex - sx, incy = ey - sy;
  const float eplLength = sqrt_rn(incx * incx + incy * incy);
  incx *= params.sample_dist / eplLength;
  incy *= params.sample_dist / eplLength;

  float cpx = sx, cpy = sy;
  float val_cp_m2 = bilinear_u8(img, step, cpx - 2.0f * incx, cpy - 2.0f * incy);
  float val_cp_m1 = bilinear_u8(img, step, cpx - incx, cpy - incy);
  float val_cp = bilinear_u8(img, step, cpx, cpy);
  float val_cp_p1 = bilinear_u8(img, step, cpx + incx, cpy + incy);
  float val_cp_p2;

  int loopCounter = 0;
  float best_match_x = -1, best_match_y = -1;
  float best_match_err = FLT_MAX;
  float second_best_match_err = FLT_MAX;
  const float qnan = __uint_as_float(0x7fc00000u);
  float best_match_errPre = qnan, best_match_errPost = qnan;
  float best_match_DiffErrPre = qnan, best_match_DiffErrPost = qnan;
  bool bestWasLastLoop = false;
  float eeLast = -1;
  float e1A = qnan, e1B = qnan, e2A = qnan, e2B = qnan, e3A = qnan, e3B = qnan, e4A = qnan,
        e4B = qnan, e5A = qnan, e5B = qnan;
  int loopCBest = -1, loopCSecond = -1;

  while ((((incx < 0) == (cpx > ex)) && ((incy < 0) == (cpy > ey))) || loopCounter == 0) {
    val_cp_p2 = bilinear_u8(img, step, cpx + 2 * incx, cpy + 2 * incy);
    float ee = 0.0f;
    if (loopCounter % 2 == 0) {
      e1A = val_cp_p2 - realVal_p2; ee += e1A * e1A;
      e2A = val_cp_p1 - realVal_p1; ee += e2A * e2A;
      e3A = val_cp - realVal;       ee += e3A * e3A;
      e4A = val_cp_m1 - realVal_m1; ee += e4A * e4A;
      e5A = val_cp_m2 - realVal_m2; ee += e5A * e5A;
    } else {
      e1B = val_cp_p2 - realVal_p2; ee += e1B * e1B;
      e2B = val_cp_p1 - realVal_p1; ee += e2B * e2B;
      e3B = val_cp - realVal;       ee += e3B * e3B;
      e4B = val_cp_m1 - realVal_m1; ee += e4B * e4B;
      e5B = val_cp_m2 - realVal_m2; ee += e5B * e5B;
    }
    if (ee < best_match_err) {
      second_best_match_err = best_match_err;
      loopCSecond = loopCBest;
      best_match_err = ee;
      loopCBest = loopCounter;
      best_match_errPre = eeLast;
      best_match_DiffErrPre = e1A * e1B + e2A * e2B + e3A * e3B + e4A * e4B + e5A * e5B;
      best_match_errPost = -1;
      best_match_DiffErrPost = -1;
      best_match_x = cpx;
      best_match_y = cpy;
      bestWasLastLoop = true;
    } else {
      if (bestWasLastLoop) {
        best_match_errPost = ee;
        best_match_DiffErrPost = e1A * e1B + e2A * e2B + e3A * e3B + e4A * e4B + e5A * e5B;
        bestWasLastLoop = false;
      }
      if (ee < second_best_match_err) {
        second_best_match_err = ee;
        loopCSecond = loopCounter;
      }
    }
    eeLast = ee;
    val_cp_m2 = val_cp_m1;
    val_cp_m1 = val_cp;
    val_cp = val_cp_p1;
    val_cp_p1 = val_cp_p2;
    cpx += incx;
    cpy += incy;
    loopCounter++;
  }

  *residual = best_match_err;
  if (best_match_err > 4.0f * params.max_cost) return 2;  // FAIL_MAX_COST (:261)
  int dl = loopCBest - loopCSecond;
  dl = (dl > 0) ? dl : -dl;
  if (((float)dl > 1.0f) && (params.second_best_factor * best_match_err > second_best_match_err))
    return 1;  // FAIL_AMBIGUOUS_MATCH (:275)

  if (params.do_subpixel) {
    const float gradPre_pre = -(best_match_errPre - best_match_DiffErrPre);
    const float gradPre_this = +(best_match_err - best_match_DiffErrPre);
    const float gradPost_this = -(best_match_err - best_match_DiffErrPost);
    const float gradPost_post = +(best_match_errPost - best_match_DiffErrPost);
    bool interpPost = false, interpPre = false;
    if (best_match_errPre < 0 || best_match_errPost < 0) {
    } else if ((gradPost_this < 0) ^ (gradPre_this < 0)) {
    } else if ((gradPre_pre < 0) ^ (gradPre_this < 0)) {
      if ((gradPost_post < 0) ^ (gradPost_this < 0)) {
      } else {
        interpPre = true;
      }
    } else if ((gradPost_post < 0) ^ (gradPost_this < 0)) {
      interpPost = true;
    }
    if (interpPre) {
      const float d = gradPre_this / (gradPre_this - gradPre_pre);
      best_match_x -= d * incx;
      best_match_y -= d * incy;
      best_match_err = best_match_err - 2 * d * gradPre_this - (gradPre_pre - gradPre_this) * d * d;
    } else if (interpPost) {
      const float d = gradPost_this / (gradPost_this - gradPost_post);
      best_match_x += d * incx;
      best_match_y += d * incy;
      best_match_err = best_match_err + 2 * d * gradPost_this + (gradPost_post - gradPost_this) * d * d;
    }
  }
  *residual = best_match_err;

  const float sampleDist = params.sample_dist * rescale_factor;
  float gradAlongLine = 0;
  float tmp = realVal_p2 - realVal_p1;
  gradAlongLine += tmp * tmp;
  tmp = realVal_p1 - realVal;
  gradAlongLine += tmp * tmp;
  tmp = realVal - realVal_m1;
  gradAlongLine += tmp * tmp;
  tmp = realVal_m1 - realVal_m2;
  gradAlongLine += tmp * tmp;
  gradAlongLine /= sampleDist * sampleDist;
  if (best_match_err > params.max_cost + sqrt_rn(gradAlongLine) * 20) return 2;  // :370

  *ucx = best_match_x;
  *ucy = best_match_y;
  return 0;
}

// ------------------------------------------------------------------------------------------
// stereo::inverse_depth_filter
// ------------------------------------------------------------------------------------------
// .cc:35-63
__device__ bool filter_predict(const EpiGeo& g, float process_var_factor, float ux, float uy,
                               float mu, float var, float* cx, float* cy, float* mu_pred,
                               float* var_pred) {
  g.project_idepth(ux, uy, mu, cx, cy, mu_pred);
  if (*mu_pred < 0.0f) {
    *mu_pred = 0.0f;
    *var_pred = (float)1e10;
    return false;
  }
  float var_factor4 = *mu_pred / mu;
  var_factor4 *= var_factor4;
  var_factor4 *= var_factor4;
  if ((double)mu < 1e-6) var_factor4 = 1;
  *var_pred = process_var_factor * var_factor4 * var;
  return true;
}

// .cc:65-182
__device__ bool filter_get_search_region(const flame_b200_filter_params& p, const EpiGeo& g,
                                         int width, int height, float ux, float uy, float mu_prior,
                                         float var_prior, float* sx, float* sy, float* ex, float* ey) {
  float id_min = p.idepth_min, id_max = p.idepth_max;
  if (!isnan(mu_prior) && !isnan(var_prior)) {
    const float sigma = sqrt_rn(var_prior);
    id_min = mu_prior - p.search_sigma * sigma;
    id_max = mu_prior + p.search_sigma * sigma;
  }
  id_min = (id_min < p.idepth_min) ? p.idepth_min : id_min;
  id_max = (id_max > p.idepth_max) ? p.idepth_max : id_max;
  if (id_max < id_min) return false;

  g.project(ux, uy, id_min, sx, sy);
  g.project(ux, uy, id_max, ex, ey);
  float dx = *ex - *sx, dy = *ey - *sy;
  float epilength = sqrt_rn(dx * dx + dy * dy);
  if (epilength <= 0) return false;
  const float epx = dx / epilength, epy = dy / epilength;

  const float rx0 = 1.0f, ry0 = 1.0f, rx1 = (float)(width - 1), ry1 = (float)(height - 1);
  float a, b, c, d;
  if (!clip_line(rx0, rx1, ry0, ry1, *sx, *sy, *ex, *ey, &a, &b, &c, &d)) return false;
  *sx = a; *sy = b; *ex = c; *ey = d;
  dx = *ex - *sx; dy = *ey - *sy;
  epilength = sqrt_rn(dx * dx + dy * dy);
  if (epilength <= 0) return false;
  if (epilength < p.epilength_min) {
    const float pad = (p.epilength_min - epilength) / 2.0f;
    *sx -= pad * epx; *sy -= pad * epy;
    *ex += pad * epx; *ey += pad * epy;
  }
  if (epilength > p.epilength_max) {
    epilength = p.epilength_max;
    *ex = *sx + epilength * epx;
    *ey = *sy + epilength * epy;
  }
  if (!clip_line(rx0, rx1, ry0, ry1, *sx, *sy, *ex, *ey, &a, &b, &c, &d)) return false;
  *sx = a; *sy = b; *ex = c; *ey = d;
  return true;
}

// .cc:184-266
__device__ int filter_search(const flame_b200_filter_params& p, const EpiGeo& g,
                             float rescale_factor, const unsigned char* img_ref,
                             const unsigned char* img_cmp, int step, float urx, float ury, float sx,
                             float sy, float ex, float ey, float* ucx, float* ucy) {
  float erx, ery;
  g.reference_epiline(urx, ury, &erx, &ery);
  float ref_patch[5];
  ref_patch[0] = bilinear_u8(img_ref, step, urx - 2 * erx * rescale_factor, ury - 2 * ery * rescale_factor);
  ref_patch[1] = bilinear_u8(img_ref, step, urx - erx * rescale_factor, ury - ery * rescale_factor);
  ref_patch[2] = bilinear_u8(img_ref, step, urx, ury);
  ref_patch[3] = bilinear_u8(img_ref, step, urx + erx * rescale_factor, ury + ery * rescale_factor);
  ref_patch[4] = bilinear_u8(img_ref, step, urx + 2 * erx * rescale_factor, ury + 2 * ery * rescale_factor);
  float ref_patch_grad = 0.0f;
#pragma unroll
  for (int ii = 1; ii < 5; ++ii) {
    const float abs_grad = fast_absf(ref_patch[ii] - ref_patch[ii - 1]);
    if (abs_grad > ref_patch_grad) ref_patch_grad = abs_grad;
  }
  if (ref_patch_grad < p.min_grad_mag) return 1;  // FAIL_REF_PATCH_GRADIENT
  float residual = FLT_MAX;
  const int result = line_stereo_match(rescale_factor, ref_patch, img_cmp, step, sx, sy, ex, ey, ucx,
                                       ucy, &residual, p.sparams);
  if (result == 1) return 2;  // FAIL_AMBIGUOUS_MATCH
  if (result == 2) return 3;  // FAIL_MAX_COST
  return 0;
}

// .cc:268-305
__device__ bool filter_update(float mu_pred, float var_pred, float mu_meas, float var_meas,
                              float* mu_post, float* var_post, float outlier_sigma_thresh) {
  if (!isnan(mu_pred) && (mu_pred > 0.0f)) {
    const float w = var_pred + var_meas;
    *mu_post = (var_meas * mu_pred + var_pred * mu_meas) / w;
    *var_post = (var_pred * var_meas) / w;
  } else {
    *mu_post = mu_meas;
    *var_post = var_meas;
  }
  const float res = (mu_meas - mu_pred);
  const float dist = res * res / var_pred;
  if (dist > outlier_sigma_thresh * outlier_sigma_thresh) return false;
  *mu_post = (*mu_post <= 0) ? 0.0f : *mu_post;
  return true;
}

// inverse_depth_meas_model.cc:49-156
__device__ bool meas_idepth(const flame_b200_meas_params& p, const EpiGeo& g, const float* gradx,
                            const float* grady, int step, float urx, float ury, float ucx, float ucy,
                            float* mu, float* var) {
  float ix, iy, ex, ey;
  g.epiline(urx, ury, &ix, &iy, &ex, &ey);
  const float disp = ex * (ucx - ix) + ey * (ucy - iy);  // epipolar_geometry.h:342-347
  if ((double)disp < 1e-3) {
    *mu = 0.0f; *var = (float)1e10;
    return false;
  }
  *mu = g.disparity_to_idepth(urx, ury, ix, iy, ex, ey, disp);
  if (*mu < 0.0f) {
    *mu = 0.0f; *var = (float)1e10;
    return false;
  }
  const float off = (float)(p.win_size / 2 + 1);  // cc:87
  const float gx = bilinear_f32(gradx, step, ucx + off, ucy + off);
  const float gy = bilinear_f32(grady, step, ucx + off, ucy + off);
  const float gnorm = sqrt_rn(gx * gx + gy * gy);
  if ((double)gnorm < 1e-3) {
    *mu = 0.0f; *var = (float)1e10;
    return false;
  }
  const float ngx = gx / gnorm, ngy = gy / gnorm;
  const float epi_dot_ngrad = ngx * ex + ngy * ey;
  const float geo_var = p.epipolar_line_var / (epi_dot_ngrad * epi_dot_ngrad);
  if ((double)fast_absf(epi_dot_ngrad) < 1e-3) {
    *mu = 0.0f; *var = (float)1e10;
    return false;
  }
  const float epi_dot_grad = gx * ex + gy * ey;
  const float photo_var = 2 * p.pixel_var / (epi_dot_grad * epi_dot_grad);
  const float disp_min = disp - disp / 10;
  const float disp_max = disp + disp / 10;
  const float idepth_min = g.disparity_to_idepth(urx, ury, ix, iy, ex, ey, disp_min);
  const float idepth_max = g.disparity_to_idepth(urx, ury, ix, iy, ex, ey, disp_max);
  const float alpha = (idepth_max - idepth_min) / (disp_max - disp_min);
  *var = alpha * alpha * (geo_var + photo_var);
  return true;
}

// cv::Rect_<int>::contains(Point2f)
__device__ __forceinline__ bool rect_contains(int rx, int ry, int rw, int rh, float px, float py) {
  return (float)rx <= px && px < (float)(rx + rw) && (float)ry <= py && py < (float)(ry + rh);
}

// flame.cc:1344-1376
__device__ __forceinline__ void feature_fail(const flame_b200_track_params& P, flame_b200_feature& f) {
  f.idepth_var *= P.fparams.process_fail_var_factor;
  if (f.idepth_var > P.idepth_var_max) f.valid = 0;
  f.num_dropouts++;
  if (f.num_dropouts > (uint32_t)P.max_dropouts) f.valid = 0;
}

// ------------------------------------------------------------------------------------------
// Flame::updateFeatureIDepths loop body (flame.cc:1308-1496) + Flame::trackFeature (:1536-1752)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_update_feature_idepths(StereoArgs a) {
  const int ii = blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= a.n) return;
  const flame_b200_track_params& P = a.P;
  flame_b200_feature f = a.feats[ii];
  const float qnan = __uint_as_float(0x7fc00000u);
  flame_b200_track_diag dg = {qnan, qnan, qnan, qnan, 0};
  const int step = a.width + 2 * a.pad;
  const FrameDev fnew = a.frames[a.fnew];

  EpiGeo epigeo;
  epigeo.init(a.K, a.Kinv);
  float T[7];
  {
    const FrameDev fref = a.frames[f.frame_id];
    se3_inv_mul(fnew.pose, fref.pose, T);  // :1314
  }
  epigeo.load(T, T + 4);

  bool done = false;
  const float baseline = sqrt_rn(T[4] * T[4] + T[5] * T[5] + T[6] * T[6]);  // :1320
  if (baseline < P.min_baseline) done = true;  // skipped: feature untouched

  float flow_x = qnan, flow_y = qnan;
  if (!done) {
    // ---------------- trackFeature ----------------
    bool track_success = false;
    int stage = 0;
    float ucx, ucy, idepth_cmp, var_cmp;
    const bool pred_success = filter_predict(epigeo, P.fparams.process_var_factor, f.x, f.y,
                                             f.idepth_mu, f.idepth_var, &ucx, &ucy, &idepth_cmp,
                                             &var_cmp);  // :1553
    if (!pred_success) {
      stage = 1;
    } else {
      int row_offset = 0;
      if (P.do_letterbox) row_offset = a.height / 3;
      const int border = (int)(P.rescale_factor_max * P.fparams.win_size / 2 + 1);  // :1576
      const int vr_x = border, vr_y = border + row_offset;
      const int vr_w = a.width - 2 * border, vr_h = a.height - 2 * border - 2 * row_offset;
      float rescale_factor = 1.0f;
      if ((f.idepth_mu > 0.0f) && (idepth_cmp > 0.0f)) rescale_factor = idepth_cmp / f.idepth_mu;

      if ((rescale_factor <= P.rescale_factor_min) || (rescale_factor >= P.rescale_factor_max)) {
        // :1604-1659 re-parent to the current pose-frame
        EpiGeo epipf;
        epipf.init(a.K, a.Kinv);
        float Tp[7];
        const FrameDev fpf = a.frames[a.curr_pf];
        const FrameDev fref = a.frames[f.frame_id];
        se3_inv_mul(fpf.pose, fref.pose, Tp);
        epipf.load(Tp, Tp + 4);
        float upx, upy, idepth_pf, var_pf;
        const bool move_success = filter_predict(epipf, P.fparams.process_var_factor, f.x, f.y,
                                                 f.idepth_mu, f.idepth_var, &upx, &upy, &idepth_pf,
                                                 &var_pf);
        if (!move_success || !rect_contains(vr_x, vr_y, vr_w, vr_h, upx, upy)) {
          f.valid = 0;
          stage = 3;
        } else {
          f.frame_id = (uint32_t)a.curr_pf;
          f.x = upx;
          f.y = upy;
          const float old_idepth = f.idepth_mu;
          f.idepth_mu = idepth_pf;
          float var_factor4 = idepth_pf / old_idepth;
          var_factor4 *= var_factor4;
          var_factor4 *= var_factor4;
          if ((double)idepth_pf < 1e-6) var_factor4 = 1;
          f.idepth_var *= var_factor4;
          stage = 2;
        }
      } else {
        float sx, sy, ex, ey;
        const bool region_success = filter_get_search_region(P.fparams, epigeo, a.width, a.height, f.x,
                                                             f.y, f.idepth_mu, f.idepth_var, &sx, &sy,
                                                             &ex, &ey);  // :1663
        if (!region_success) {
          stage = 4;
        } else if (!rect_contains(vr_x, vr_y, vr_w, vr_h, f.x, f.y)) {
          stage = 5;
        } else {
          const float off = (float)a.pad;  // :1677-1678
          const FrameDev fref = a.frames[f.frame_id];
          const int status = filter_search(P.fparams, epigeo, rescale_factor, fref.img, fnew.img, step,
                                           f.x + off, f.y + off, sx + off, sy + off, ex + off,
                                           ey + off, &ucx, &ucy);  // :1690
          f.search_status = status;
          if (status != 0) {
            stage = 6;
          } else {
            flow_x = ucx - off;
            flow_y = ucy - off;
            track_success = true;
          }
        }
      }
    }

    if (!track_success) {
      feature_fail(P, f);
      dg.stage = stage;
    } else {
      dg.flow_x = flow_x;
      dg.flow_y = flow_y;
      // ---------------- measurement (:1380-1390) ----------------
      float mu_meas, var_meas;
      const bool sense_success = meas_idepth(P.zparams, epigeo, fnew.gx, fnew.gy, step, f.x, f.y,
                                             flow_x, flow_y, &mu_meas, &var_meas);
      dg.mu_meas = mu_meas;
      dg.var_meas = var_meas;
      if (!sense_success) {
        feature_fail(P, f);
        dg.stage = 7;
      } else {
        float mu_post, var_post;
        const bool fuse_success = filter_update(f.idepth_mu, f.idepth_var, mu_meas, var_meas, &mu_post,
                                                &var_post, P.outlier_sigma_thresh);  // :1436
        if (!fuse_success) {
          feature_fail(P, f);
          dg.stage = 8;
        } else {
          if (P.do_meas_fusion) {
            f.idepth_mu = mu_post;
            f.idepth_var = var_post;
          } else {
            f.idepth_mu = mu_meas;
            f.idepth_var = var_meas;
          }
          f.valid = 1;
          f.num_updates++;
          f.num_dropouts = 0;
          dg.stage = 9;
          atomicAdd(a.count, 1);
        }
      }
    }
    a.feats[ii] = f;
  }
  if (a.diag) a.diag[ii] = dg;
}

// ------------------------------------------------------------------------------------------
// Frame preparation: the padded level-0 buffers of utils::Frame::create
// (src/flame/utils/frame.cc:54-63) from the raw u8 image, on the device: img_pad = REFLECT_101
// border, gradients = getCentralGradient<uint8_t,float> (src/flame/utils/image_utils.cc:36-108:
// 0.5*(right-left) inside, forward/backward differences on the first/last column/row) with a
// zero (BORDER_CONSTANT) border. One thread per padded pixel; every output written once.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) p = (p < 0) ? -p : 2 * (len - 1) - p;
  return p;
}

__global__ void __launch_bounds__(256)
k_frame_prep(const unsigned char* __restrict__ img, int width, int height, int border,
             unsigned char* __restrict__ img_pad, float* __restrict__ gx_pad, float* __restrict__ gy_pad) {
  const int pw = width + 2 * border, ph = height + 2 * border;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= pw || y >= ph) return;
  const int sx = x - border, sy = y - border;
  const size_t o = (size_t)y * pw + x;
  img_pad[o] = __ldg(img + (size_t)reflect101(sy, height) * width + reflect101(sx, width));
  float gx = 0.0f, gy = 0.0f;
  if (sx >= 0 && sx < width && sy >= 0 && sy < height) {
    const unsigned char* r = img + (size_t)sy * width;
    if (sx == 0) gx = (float)__ldg(r + 1) - (float)__ldg(r);
    else if (sx == width - 1) gx = (float)__ldg(r + sx) - (float)__ldg(r + sx - 1);
    else gx = 0.5f * ((float)__ldg(r + sx + 1) - (float)__ldg(r + sx - 1));
    if (sy == 0) gy = (float)__ldg(r + width + sx) - (float)__ldg(r + sx);
    else if (sy == height - 1) gy = (float)__ldg(r + sx) - (float)__ldg(r - width + sx);
    else gy = 0.5f * ((float)__ldg(r + width + sx) - (float)__ldg(r - width + sx));
  }
  gx_pad[o] = gx;
  gy_pad[o] = gy;
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static StereoState* get_state(void** st) {
  if (!*st) {
    StereoState* s = new StereoState();
    std::memset(s->h_frames, 0, sizeof(s->h_frames));
    *st = s;
  }
  return static_cast<StereoState*>(*st);
}

int stereo_set_camera(void** state, int width, int height, const float K[9], const float Kinv[9]) {
  StereoState* s = get_state(state);
  s->width = width;
  s->height = height;
  std::memcpy(s->K, K, sizeof(s->K));
  std::memcpy(s->Kinv, Kinv, sizeof(s->Kinv));
  return 0;
}

int stereo_frame_upload(void** state, cudaStream_t st, int fslot, const uint8_t* img_pad,
                        const float* gradx_pad, const float* grady_pad, int pad, const float pose[7]) {
  StereoState* s = get_state(state);
  if (s->width <= 0) {
    set_error("frame_upload: call flame_b200_set_camera first");
    return FLAME_B200_ESTATE;
  }
  if (fslot < 0 || fslot >= FB_MAX_FRAMES || pad < 0) {
    set_error("frame_upload: frame slot out of range");
    return FLAME_B200_EINVAL;
  }
  if (s->pad != 0 && s->pad != pad) {
    bool any = false;
    for (int i = 0; i < FB_MAX_FRAMES; ++i) any |= (s->h_frames[i].valid != 0 && i != fslot);
    if (any) {
      set_error("frame_upload: all frames must share one padding");
      return FLAME_B200_EINVAL;
    }
  }
  s->pad = pad;
  const size_t npix = (size_t)(s->width + 2 * pad) * (s->height + 2 * pad);
  if (npix > s->cap_pix[fslot]) {
    if (s->d_img[fslot]) FB_CUDA(cudaFree(s->d_img[fslot]));
    if (s->d_gx[fslot]) FB_CUDA(cudaFree(s->d_gx[fslot]));
    if (s->d_gy[fslot]) FB_CUDA(cudaFree(s->d_gy[fslot]));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_img[fslot]), npix));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_gx[fslot]), npix * 4));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_gy[fslot]), npix * 4));
    s->cap_pix[fslot] = npix;
  }
  if (!s->d_frames) FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_frames), sizeof(s->h_frames)));
  if (!s->stage_free) FB_CUDA(cudaEventCreateWithFlags(&s->stage_free, cudaEventDisableTiming));
  const size_t need = npix * 9 + sizeof(s->h_frames);
  FB_CUDA(cudaEventSynchronize(s->stage_free));
  if (need > s->h_stage_cap) {
    if (s->h_stage) FB_CUDA(cudaFreeHost(s->h_stage));
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_stage), need));
    s->h_stage_cap = need;
  }
  unsigned char* hs = s->h_stage;
  std::memcpy(hs, gradx_pad, npix * 4);
  std::memcpy(hs + npix * 4, grady_pad, npix * 4);
  std::memcpy(hs + npix * 8, img_pad, npix);
  FrameDev& fr = s->h_frames[fslot];
  fr.img = s->d_img[fslot];
  fr.gx = s->d_gx[fslot];
  fr.gy = s->d_gy[fslot];
  std::memcpy(fr.pose, pose, sizeof(fr.pose));
  fr.valid = 1;
  std::memcpy(hs + npix * 9, s->h_frames, sizeof(s->h_frames));
  FB_CUDA(cudaMemcpyAsync(s->d_gx[fslot], hs, npix * 4, cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemcpyAsync(s->d_gy[fslot], hs + npix * 4, npix * 4, cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemcpyAsync(s->d_img[fslot], hs + npix * 8, npix, cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemcpyAsync(s->d_frames, hs + npix * 9, sizeof(s->h_frames), cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaEventRecord(s->stage_free, st));
  return 0;
}

static int frame_slot_prepare(StereoState* s, int fslot, int pad, size_t npix) {
  if (s->width <= 0) {
    set_error("frame_upload: call flame_b200_set_camera first");
    return FLAME_B200_ESTATE;
  }
  if (fslot < 0 || fslot >= FB_MAX_FRAMES || pad < 0) {
    set_error("frame_upload: frame slot out of range");
    return FLAME_B200_EINVAL;
  }
  if (s->pad != 0 && s->pad != pad) {
    bool any = false;
    for (int i = 0; i < FB_MAX_FRAMES; ++i) any |= (s->h_frames[i].valid != 0 && i != fslot);
    if (any) {
      set_error("frame_upload: all frames must share one padding");
      return FLAME_B200_EINVAL;
    }
  }
  s->pad = pad;
  if (npix > s->cap_pix[fslot]) {
    if (s->d_img[fslot]) FB_CUDA(cudaFree(s->d_img[fslot]));
    if (s->d_gx[fslot]) FB_CUDA(cudaFree(s->d_gx[fslot]));
    if (s->d_gy[fslot]) FB_CUDA(cudaFree(s->d_gy[fslot]));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_img[fslot]), npix));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_gx[fslot]), npix * 4));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_gy[fslot]), npix * 4));
    s->cap_pix[fslot] = npix;
  }
  if (!s->d_frames) FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_frames), sizeof(s->h_frames)));
  if (!s->stage_free) FB_CUDA(cudaEventCreateWithFlags(&s->stage_free, cudaEventDisableTiming));
  return 0;
}

// Raw (unpadded) u8 image in, padded image + gradients computed on the device.
int stereo_frame_upload_raw(void** state, cudaStream_t st, int fslot, const uint8_t* img, int pad,
                            const float pose[7]) {
  StereoState* s = get_state(state);
  const size_t npix = (size_t)(s->width + 2 * pad) * (s->height + 2 * pad);
  if (int rc = frame_slot_prepare(s, fslot, pad, npix)) return rc;
  const size_t nraw = (size_t)s->width * s->height;
  const size_t need = nraw + sizeof(s->h_frames);
  FB_CUDA(cudaEventSynchronize(s->stage_free));
  if (need > s->h_stage_cap) {
    if (s->h_stage) FB_CUDA(cudaFreeHost(s->h_stage));
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_stage), need));
    s->h_stage_cap = need;
  }
  if (nraw > s->cap_raw) {
    if (s->d_raw) FB_CUDA(cudaFree(s->d_raw));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_raw), nraw));
    s->cap_raw = nraw;
  }
  std::memcpy(s->h_stage, img, nraw);
  FrameDev& fr = s->h_frames[fslot];
  fr.img = s->d_img[fslot];
  fr.gx = s->d_gx[fslot];
  fr.gy = s->d_gy[fslot];
  std::memcpy(fr.pose, pose, sizeof(fr.pose));
  fr.valid = 1;
  std::memcpy(s->h_stage + nraw, s->h_frames, sizeof(s->h_frames));
  FB_CUDA(cudaMemcpyAsync(s->d_raw, s->h_stage, nraw, cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemcpyAsync(s->d_frames, s->h_stage + nraw, sizeof(s->h_frames), cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaEventRecord(s->stage_free, st));
  const int pw = s->width + 2 * pad, ph = s->height + 2 * pad;
  k_frame_prep<<<dim3((pw + 255) / 256, ph), 256, 0, st>>>(s->d_raw, s->width, s->height, pad,
                                                         s->d_img[fslot], s->d_gx[fslot], s->d_gy[fslot]);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

int stereo_frame_download(void* state, cudaStream_t st, int fslot, uint8_t* img_pad, float* gradx_pad,
                          float* grady_pad) {
  StereoState* s = static_cast<StereoState*>(state);
  if (!s || fslot < 0 || fslot >= FB_MAX_FRAMES || !s->h_frames[fslot].valid) {
    set_error("frame_download: frame slot not uploaded");
    return FLAME_B200_EINVAL;
  }
  const size_t npix = (size_t)(s->width + 2 * s->pad) * (s->height + 2 * s->pad);
  if (img_pad) FB_CUDA(cudaMemcpyAsync(img_pad, s->d_img[fslot], npix, cudaMemcpyDeviceToHost, st));
  if (gradx_pad) FB_CUDA(cudaMemcpyAsync(gradx_pad, s->d_gx[fslot], npix * 4, cudaMemcpyDeviceToHost, st));
  if (grady_pad) FB_CUDA(cudaMemcpyAsync(grady_pad, s->d_gy[fslot], npix * 4, cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int stereo_update(void* state, cudaStream_t st, const flame_b200_track_params* params, int fnew,
                  int curr_pf, flame_b200_feature* feats, int32_t n_feats,
                  flame_b200_track_diag* diag, int32_t* n_updates) {
  StereoState* s = static_cast<StereoState*>(state);
  if (!s || s->width <= 0) {
    set_error("update_feature_idepths: camera / frames not set");
    return FLAME_B200_ESTATE;
  }
  if (fnew < 0 || fnew >= FB_MAX_FRAMES || curr_pf < 0 || curr_pf >= FB_MAX_FRAMES ||
      !s->h_frames[fnew].valid || !s->h_frames[curr_pf].valid) {
    set_error("update_feature_idepths: frame slot not uploaded");
    return FLAME_B200_EINVAL;
  }
  for (int i = 0; i < n_feats; ++i)
    if (feats[i].frame_id >= FB_MAX_FRAMES || !s->h_frames[feats[i].frame_id].valid) {
      set_error("update_feature_idepths: feature parented to a frame slot that was not uploaded");
      return FLAME_B200_EINVAL;
    }
  if (n_updates) *n_updates = 0;
  if (n_feats == 0) return 0;
  if (n_feats > s->cap_feats) {
    const int n = n_feats + n_feats / 2;
    if (s->d_feats) FB_CUDA(cudaFree(s->d_feats));
    if (s->d_diag) FB_CUDA(cudaFree(s->d_diag));
    if (s->h_feats) FB_CUDA(cudaFreeHost(s->h_feats));
    if (s->h_diag) FB_CUDA(cudaFreeHost(s->h_diag));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_feats), (size_t)n * sizeof(flame_b200_feature)));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_diag), (size_t)n * sizeof(flame_b200_track_diag)));
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_feats), (size_t)n * sizeof(flame_b200_feature)));
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_diag), (size_t)n * sizeof(flame_b200_track_diag)));
    s->cap_feats = n;
  }
  if (!s->d_count) {
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_count), sizeof(int)));
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_count), sizeof(int)));
  }
  std::memcpy(s->h_feats, feats, (size_t)n_feats * sizeof(flame_b200_feature));
  FB_CUDA(cudaMemcpyAsync(s->d_feats, s->h_feats, (size_t)n_feats * sizeof(flame_b200_feature),
                          cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemsetAsync(s->d_count, 0, sizeof(int), st));
  StereoArgs a;
  a.P = *params;
  std::memcpy(a.K, s->K, sizeof(a.K));
  std::memcpy(a.Kinv, s->Kinv, sizeof(a.Kinv));
  a.width = s->width;
  a.height = s->height;
  a.pad = s->pad;
  a.fnew = fnew;
  a.curr_pf = curr_pf;
  a.frames = s->d_frames;
  a.feats = s->d_feats;
  a.diag = s->d_diag;
  a.n = n_feats;
  a.count = s->d_count;
  k_update_feature_idepths<<<(n_feats + 127) / 128, 128, 0, st>>>(a);
  count_launch();
  FB_CUDA(cudaGetLastError());
  FB_CUDA(cudaMemcpyAsync(s->h_feats, s->d_feats, (size_t)n_feats * sizeof(flame_b200_feature),
                          cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaMemcpyAsync(s->h_diag, s->d_diag, (size_t)n_feats * sizeof(flame_b200_track_diag),
                          cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaMemcpyAsync(s->h_count, s->d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  std::memcpy(feats, s->h_feats, (size_t)n_feats * sizeof(flame_b200_feature));
  if (diag) std::memcpy(diag, s->h_diag, (size_t)n_feats * sizeof(flame_b200_track_diag));
  if (n_updates) *n_updates = *s->h_count;
  return 0;
}

// ------------------------------------------------------------------------------------------
// Flame::projectGraph, per-vertex body (flame.cc:1892-1899): pos, x <- EpipolarGeometry::project
// ------------------------------------------------------------------------------------------
struct ProjectArgs {
  float K[9], Kinv[9], q[4], t[3];
  float scale;
};

__global__ void __launch_bounds__(128) k_graph_project(GraphDev g, float* __restrict__ pos,
                                                       ProjectArgs a, float* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  EpiGeo e;
  e.init(a.K, a.Kinv);
  e.load(a.q, a.t);
  const int o = g.vperm[v];
  const float x = g.vprim[v].x;
  float cx, cy, idn;
  e.project_idepth(pos[2 * o], pos[2 * o + 1], x * a.scale, &cx, &cy, &idn);
  pos[2 * o] = cx;
  pos[2 * o + 1] = cy;
  g.vprim[v].x = idn / a.scale;
  out[2 * o] = cx;
  out[2 * o + 1] = cy;
  out[2 * (size_t)g.V + o] = idn;
}

int stereo_graph_project(void* state, cudaStream_t st, const GraphDev& g, float* d_pos,
                         const float q[4], const float t[3], float scale, float* d_out) {
  StereoState* s = static_cast<StereoState*>(state);
  if (!s || s->width <= 0) {
    set_error("graph_project: camera not set (flame_b200_set_camera)");
    return FLAME_B200_ESTATE;
  }
  if (g.V == 0) return 0;
  ProjectArgs a;
  std::memcpy(a.K, s->K, sizeof(a.K));
  std::memcpy(a.Kinv, s->Kinv, sizeof(a.Kinv));
  std::memcpy(a.q, q, sizeof(a.q));
  std::memcpy(a.t, t, sizeof(a.t));
  a.scale = scale;
  k_graph_project<<<(g.V + 127) / 128, 128, 0, st>>>(g, d_pos, a, d_out);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

void stereo_destroy(void* state) {
  StereoState* s = static_cast<StereoState*>(state);
  if (!s) return;
  for (int i = 0; i < FB_MAX_FRAMES; ++i) {
    if (s->d_img[i]) cudaFree(s->d_img[i]);
    if (s->d_gx[i]) cudaFree(s->d_gx[i]);
    if (s->d_gy[i]) cudaFree(s->d_gy[i]);
  }
  if (s->d_frames) cudaFree(s->d_frames);
  if (s->d_raw) cudaFree(s->d_raw);
  if (s->h_stage) cudaFreeHost(s->h_stage);
  if (s->stage_free) cudaEventDestroy(s->stage_free);
  if (s->d_feats) cudaFree(s->d_feats);
  if (s->d_diag) cudaFree(s->d_diag);
  if (s->d_count) cudaFree(s->d_count);
  if (s->h_feats) cudaFreeHost(s->h_feats);
  if (s->h_diag) cudaFreeHost(s->h_diag);
  if (s->h_count) cudaFreeHost(s->h_count);
  delete s;
}

}  // namespace fb
