// Epipolar inverse-depth search + measurement model + fusion for sm_100a: one WARP per feature.
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
// Design (the reference walks the epipolar line with a scalar sliding window, one feature at a time):
//   * k_epigeo_table: the geometry of a (parent frame -> new frame) pair depends on the frame, not on the
//     feature; one thread per frame slot builds it once per call instead of once per feature.
//   * k_track_features: one warp per feature. The scalar stages (predict, search region, clipping,
//     measurement model, fusion) are evaluated uniformly by the warp. The SSD walk is done in parallel:
//     the bounding patch of the clipped epipolar segment in the padded u8 comparison image is staged into
//     shared memory with ONE TMA tile load (cp.async.bulk.tensor.2d + mbarrier); lane k computes sample
//     position k (the reference's running sum cp += inc, so the same rounded positions), takes its
//     bilinear tap from the staged patch, forms the 5-tap SSD of step k from the shared samples in the
//     reference's summation order, and a warp arg-min with first-index tie-break (plus a second pass
//     excluding the winner) yields best / second-best / neighbours exactly as the sequential scan does.
//
// This translation unit is compiled with -fmad=false: every expression below is evaluated with
// separately rounded fp32 operations in the association order of the reference's C++ (its
// x86-64 Release build has no FMA); '/' and sqrt are IEEE-correct (-prec-div/-prec-sqrt
// defaults). The matcher is SSD, not NCC; the gradient tap uses offset 3 on the
// 5-px-padded gradient images.
#include <cuda.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>

#include "internal.h"

namespace fb {

struct FrameDev {
  const unsigned char* img;  // (H + 2 pad) rows of `pitch` bytes, (W + 2 pad) valid
  const float* gx;           // same geometry, f32 (pitch in elements)
  const float* gy;
  float pose[7];             // camera-to-world: qw qx qy qz tx ty tz
  int valid;
};

#define FB_MAX_FRAMES 64
#define FB_SEARCH_MAX_STEPS 128  // SSD positions per feature the warp kernel holds (default params: <= 36)
// The staged comparison-image patch: FB_TILE_W x FB_TILE_H bytes. A TMA tile load needs the box to START on
// a 16-byte boundary of the row (measured: an unaligned inner coordinate raises an illegal-instruction
// fault), so the patch origin is rounded down to a multiple of 16 px and the box is 64 wide: >= 49 usable
// columns, 48 rows - the default parameters need at most 43 x 43.
#define FB_TILE_W 64
#define FB_TILE_H 48

// Geometry of one (parent frame -> target frame) pair: stereo::EpipolarGeometry<float> after loadGeometry.
struct GeoDev {
  float q[4], t[3], t_cmp_to_ref[3];
  float KRKinv[9], Kt[3], epipole[2];
  float baseline;  // |t| (flame.cc:1320)
  int pad_;
};

struct StereoState {
  int width = 0, height = 0, pad = 0, pitch = 0;
  float K[9], Kinv[9];
  FrameDev h_frames[FB_MAX_FRAMES];
  FrameDev* d_frames = nullptr;
  unsigned char* d_img[FB_MAX_FRAMES] = {nullptr};
  float* d_gx[FB_MAX_FRAMES] = {nullptr};
  float* d_gy[FB_MAX_FRAMES] = {nullptr};
  size_t cap_pix[FB_MAX_FRAMES] = {0};
  int slot_w[FB_MAX_FRAMES] = {0}, slot_h[FB_MAX_FRAMES] = {0}, slot_pad[FB_MAX_FRAMES] = {0};
  CUtensorMap tmap[FB_MAX_FRAMES];   // TMA descriptor of every frame slot's padded u8 image
  unsigned char* d_raw = nullptr;    // raw u8 image of the frame being prepared
  size_t cap_raw = 0;
  unsigned char* h_stage = nullptr;  // pinned
  size_t h_stage_cap = 0;
  cudaEvent_t stage_free = nullptr;
  GeoDev* d_geo = nullptr;           // [2][FB_MAX_FRAMES]: parent -> new frame, parent -> current pose-frame
  // the resident feature set: in (read by the search), out (written), swapped by a committing download
  flame_b200_feature* d_feats = nullptr;
  flame_b200_feature* d_feats_out = nullptr;
  flame_b200_track_diag* d_diag = nullptr;
  int* d_count = nullptr;
  int cap_feats = 0, n_feats = 0;
  bool tracked = false;
  flame_b200_feature* h_feats = nullptr;  // pinned mirrors
  flame_b200_track_diag* h_diag = nullptr;
  int* h_count = nullptr;
  cudaEvent_t feats_free = nullptr;
  unsigned char* d_detect = nullptr;  // feature detection scratch + pinned mirror
  unsigned char* h_detect = nullptr;
  size_t detect_cap = 0;
};

struct StereoArgs {
  flame_b200_track_params P;
  float K[9], Kinv[9];
  int width, height, pad, pitch;
  int fnew, curr_pf;
  const FrameDev* frames;
  const GeoDev* geo;
  const flame_b200_feature* feats;
  flame_b200_feature* feats_out;
  flame_b200_track_diag* diag;
  int n;
  int* count;
  int stage_patches;  // 0: taps straight from the image (FLAME_B200_SEARCH_TMA=0, for A/B measurements)
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
  // geometry precomputed by k_epigeo_table (bit-identical to load(): same expressions, evaluated once)
  __device__ void load_table(const GeoDev& g) {
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = g.q[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      t[i] = g.t[i];
      t_cmp_to_ref[i] = g.t_cmp_to_ref[i];
      Kt[i] = g.Kt[i];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) KRKinv[i] = g.KRKinv[i];
    KRKinv3[0] = KRKinv[6]; KRKinv3[1] = KRKinv[7]; KRKinv3[2] = KRKinv[8];
    epipole[0] = g.epipole[0];
    epipole[1] = g.epipole[1];
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
// stereo::line_stereo::match (line_stereo.h:75-385), warp-parallel.
//
// The reference slides a 5-tap window along the line: at step k (k = 0, 1, ...) the window holds the
// samples S[k-2] .. S[k+2], where S[-2], S[-1], S[0], S[1] are taken at cp0 - 2 inc, cp0 - inc, cp0,
// cp0 + inc and S[k+2] at cp_k + 2 inc, cp_k being the running sum cp_{k-1} + inc. The SSD of step k is
// ee_k = sum_{i=+2..-2} (S[k+i] - ref[i])^2 in that order. The window is an optimisation, not a dependency:
// here lane k forms cp_k with the same k rounded additions, takes S[k+2] from the staged patch, and ee_k
// from the shared sample array. best = first minimum of ee, second = first minimum of the others - what
// the sequential "if (ee < best) ... else if (ee < second)" scan produces.
// ------------------------------------------------------------------------------------------
struct WarpScratch {               // per warp, shared memory
  float samp[FB_SEARCH_MAX_STEPS + 4];  // samp[j + 2] = S[j]
  float ee[FB_SEARCH_MAX_STEPS];
  float cx[FB_SEARCH_MAX_STEPS], cy[FB_SEARCH_MAX_STEPS];
};

struct PatchView {  // where the bilinear taps of the comparison image come from
  const unsigned char* base;
  int pitch;        // bytes per row
  int x0, y0;       // image coordinates of base[0]
};

__device__ __forceinline__ float tap_u8(const PatchView& v, float x, float y) {
  const int xf = (int)x, yf = (int)y;
  const float dx = x - xf, dy = y - yf;
  const float w11 = dx * dy;
  const float w01 = dx - w11;
  const float w10 = dy - w11;
  const float w00 = 1.0f - dx - dy + w11;
  const unsigned char* d = v.base + (yf - v.y0) * v.pitch + (xf - v.x0);
  return w00 * d[0] + w01 * d[1] + w10 * d[v.pitch] + w11 * d[1 + v.pitch];
}

// SSD of the 5-tap window at step k and its residuals (order p2, p1, c, m1, m2 as in :160-176)
__device__ __forceinline__ float window_ssd(const float* samp, int k, const float ref[5], float e[5]) {
  float ee = 0.0f;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    e[i] = samp[k + 4 - i] - ref[4 - i];  // S[k + 2 - i] - ref[+2 - i]
    ee += e[i] * e[i];
  }
  return ee;
}

__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
    v = w < v ? w : v;
  }
  return v;
}

// Returns 0 SUCCESS, 1 FAIL_AMBIGUOUS_MATCH, 2 FAIL_MAX_COST (line_stereo.h:38-43), 3 = more than
// FB_SEARCH_MAX_STEPS positions (rejected on the host before the launch). All arguments warp-uniform.
// `view` must cover every tap of the walk: the staged patch when the walk's bounding box fits it,
// otherwise the image itself.
__device__ int warp_line_match(const flame_b200_ls_params& params, float rescale_factor, const float ref[5],
                               const PatchView& view, float sx, float sy, float ex, float ey, WarpScratch& ws,
                               float* ucx, float* ucy) {
  const int lane = threadIdx.x & 31;
  float incx = ex - sx, incy = ey - sy;
  const float epl = sqrt_rn(incx * incx + incy * incy);
  incx *= params.sample_dist / epl;
  incy *= params.sample_dist / epl;

  // the four samples in front of the first window
  if (lane < 4) {
    float px = sx, py = sy;
    if (lane == 0) { px = sx - 2.0f * incx; py = sy - 2.0f * incy; }
    if (lane == 1) { px = sx - incx; py = sy - incy; }
    if (lane == 3) { px = sx + incx; py = sy + incy; }
    ws.samp[lane] = tap_u8(view, px, py);
  }
  // positions cp_k, the loop condition of every step, and the leading sample of every executed step
  int n = -1;  // number of executed steps
  float bx = sx, by = sy;  // cp at the first step of the round
  for (int k0 = 0; k0 < FB_SEARCH_MAX_STEPS && n < 0; k0 += 32) {
    float cx = bx, cy = by;
    for (int i = 0; i < 31; ++i)
      if (i < lane) {
        cx += incx;
        cy += incy;
      }
    const int k = k0 + lane;
    const bool go = (((incx < 0) == (cx > ex)) && ((incy < 0) == (cy > ey))) || k == 0;  // :150
    const unsigned stop = __ballot_sync(0xffffffffu, !go);
    const int first_stop = stop ? (__ffs(stop) - 1) : 32;
    if (lane < first_stop) {
      ws.cx[k] = cx;
      ws.cy[k] = cy;
      ws.samp[k + 4] = tap_u8(view, cx + 2.0f * incx, cy + 2.0f * incy);
    }
    if (stop) n = k0 + first_stop;
    bx = __shfl_sync(0xffffffffu, cx, 31) + incx;
    by = __shfl_sync(0xffffffffu, cy, 31) + incy;
  }
  if (n < 0) return 3;
  __syncwarp();
  // SSD of every step; first minimum
  unsigned long long best_key = ~0ull;
  for (int k = lane; k < n; k += 32) {
    float e[5];
    const float ee = window_ssd(ws.samp, k, ref, e);
    ws.ee[k] = ee;
    const unsigned long long key = ((unsigned long long)__float_as_uint(ee) << 32) | (unsigned)k;  // ee >= 0
    best_key = key < best_key ? key : best_key;
  }
  best_key = warp_min_u64(best_key);
  const int kb = (int)(best_key & 0xffffffffu);
  float best = __uint_as_float((unsigned)(best_key >> 32));
  __syncwarp();
  // first minimum among the other steps
  unsigned long long second_key = ((unsigned long long)__float_as_uint(FLT_MAX) << 32) | 0xffffffffu;
  for (int k = lane; k < n; k += 32)
    if (k != kb) {
      const unsigned long long key = ((unsigned long long)__float_as_uint(ws.ee[k]) << 32) | (unsigned)k;
      second_key = key < second_key ? key : second_key;
    }
  second_key = warp_min_u64(second_key);
  const float second = __uint_as_float((unsigned)(second_key >> 32));
  const int ks = (int)(second_key & 0xffffffffu);  // 0xffffffff == -1: no second candidate

  if (best > 4.0f * params.max_cost) return 2;  // :261
  int dl = kb - ks;
  dl = dl > 0 ? dl : -dl;
  if (((float)dl > 1.0f) && (params.second_best_factor * best > second)) return 1;  // :275

  float mx = ws.cx[kb], my = ws.cy[kb];
  if (params.do_subpixel && kb > 0 && kb + 1 < n) {  // otherwise one of the neighbour errors is the -1 sentinel (:287)
    // errors of the neighbours and the residual cross terms with them (:187-206)
    float e0[5], ep[5], en[5];
    window_ssd(ws.samp, kb, ref, e0);
    const float err_pre = window_ssd(ws.samp, kb - 1, ref, ep);
    const float err_post = window_ssd(ws.samp, kb + 1, ref, en);
    float cross_pre = e0[0] * ep[0], cross_post = e0[0] * en[0];
#pragma unroll
    for (int i = 1; i < 5; ++i) {
      cross_pre += e0[i] * ep[i];
      cross_post += e0[i] * en[i];
    }
    // derivatives of the SSD parabola pieces left and right of the minimum (:281-284)
    const float g_pre_pre = -(err_pre - cross_pre);
    const float g_pre_this = +(best - cross_pre);
    const float g_post_this = -(best - cross_post);
    const float g_post_post = +(err_post - cross_post);
    const bool flip_mid = (g_post_this < 0) != (g_pre_this < 0);
    const bool flip_pre = (g_pre_pre < 0) != (g_pre_this < 0);
    const bool flip_post = (g_post_post < 0) != (g_post_this < 0);
    if (!flip_mid && flip_pre && !flip_post) {  // zero crossing on the left piece only (:299-306)
      const float d = g_pre_this / (g_pre_this - g_pre_pre);
      mx -= d * incx;
      my -= d * incy;
      best = best - 2 * d * g_pre_this - (g_pre_pre - g_pre_this) * d * d;
    } else if (!flip_mid && !flip_pre && flip_post) {  // on the right piece only (:308-313)
      const float d = g_post_this / (g_post_this - g_post_post);
      mx += d * incx;
      my += d * incy;
      best = best + 2 * d * g_post_this + (g_post_post - g_post_this) * d * d;
    }
  }

  // acceptance against the texture of the reference patch (:352-372)
  const float sd = params.sample_dist * rescale_factor;
  float grad2 = 0.0f;
#pragma unroll
  for (int i = 4; i >= 1; --i) {
    const float d = ref[i] - ref[i - 1];
    grad2 += d * d;
  }
  grad2 /= sd * sd;
  if (best > params.max_cost + sqrt_rn(grad2) * 20) return 2;
  *ucx = mx;
  *ucy = my;
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

// ------------------------------------------------------------------------------------------
// Per-frame geometry table: EpipolarGeometry::loadGeometry (epipolar_geometry.h:86-103) of
// T = target^-1 * parent (flame.cc:1314, :1610) for every uploaded parent frame slot and the two targets
// (row 0: the new frame, row 1: the current pose-frame used for re-parenting).
// ------------------------------------------------------------------------------------------
__global__ void k_epigeo_table(StereoArgs a, GeoDev* __restrict__ out) {
  const int f = threadIdx.x, which = blockIdx.x;
  if (f >= FB_MAX_FRAMES || !a.frames[f].valid) return;
  const FrameDev target = a.frames[which == 0 ? a.fnew : a.curr_pf];
  const FrameDev parent = a.frames[f];
  float T[7];
  se3_inv_mul(target.pose, parent.pose, T);
  EpiGeo e;
  e.init(a.K, a.Kinv);
  e.load(T, T + 4);
  GeoDev g;
  for (int i = 0; i < 4; ++i) g.q[i] = e.q[i];
  for (int i = 0; i < 3; ++i) {
    g.t[i] = e.t[i];
    g.t_cmp_to_ref[i] = e.t_cmp_to_ref[i];
    g.Kt[i] = e.Kt[i];
  }
  for (int i = 0; i < 9; ++i) g.KRKinv[i] = e.KRKinv[i];
  g.epipole[0] = e.epipole[0];
  g.epipole[1] = e.epipole[1];
  g.baseline = sqrt_rn(T[4] * T[4] + T[5] * T[5] + T[6] * T[6]);  // flame.cc:1320
  g.pad_ = 0;
  out[which * FB_MAX_FRAMES + f] = g;
}

// ------------------------------------------------------------------------------------------
// TMA + mbarrier helpers (one patch per warp)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void patch_load_issue(const CUtensorMap* tmap, void* dst, unsigned long long* bar,
                                                 int x0, int y0) {
  const uint32_t b = smem_addr(bar), d = smem_addr(dst);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(FB_TILE_W * FB_TILE_H) : "memory");
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(d),
      "l"(tmap), "r"(x0), "r"(y0), "r"(b)
      : "memory");
}
__device__ __forceinline__ void patch_load_wait(unsigned long long* bar) {
  const uint32_t b = smem_addr(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(b)
        : "memory");
  }
}

// flame.cc:1344-1376
__device__ __forceinline__ void feature_fail(const flame_b200_track_params& P, flame_b200_feature& f) {
  f.idepth_var *= P.fparams.process_fail_var_factor;
  if (f.idepth_var > P.idepth_var_max) f.valid = 0;
  f.num_dropouts++;
  if (f.num_dropouts > (uint32_t)P.max_dropouts) f.valid = 0;
}

// cv::Rect_<int>::contains takes a Point_<int>: the Point2f is converted with saturate_cast<int>
// (cvRound, round half to even) first and the rounded point is tested (OpenCV 3.2 core/types.hpp)
__device__ __forceinline__ bool rect_contains(int rx, int ry, int rw, int rh, float px, float py) {
  const int ix = __float2int_rn(px), iy = __float2int_rn(py);
  return rx <= ix && ix < rx + rw && ry <= iy && iy < ry + rh;
}

// ------------------------------------------------------------------------------------------
// Flame::updateFeatureIDepths loop body (flame.cc:1308-1496) + Flame::trackFeature (:1536-1752):
// one warp per feature, all scalar stages evaluated uniformly by the warp, lane 0 writes.
// ------------------------------------------------------------------------------------------
#define TRACK_WARPS 4
__global__ void __launch_bounds__(32 * TRACK_WARPS)
k_track_features(StereoArgs a, const __grid_constant__ CUtensorMap tmap_cmp) {
  __shared__ __align__(128) unsigned char s_tile[TRACK_WARPS][FB_TILE_W * FB_TILE_H];
  __shared__ WarpScratch s_ws[TRACK_WARPS];
  __shared__ __align__(8) unsigned long long s_bar[TRACK_WARPS];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ii = blockIdx.x * TRACK_WARPS + w;
  if (ii >= a.n) return;  // warps are independent: no CTA-wide synchronisation below
  const flame_b200_track_params& P = a.P;
  flame_b200_feature f = a.feats[ii];
  const float qnan = __uint_as_float(0x7fc00000u);
  flame_b200_track_diag dg = {qnan, qnan, qnan, qnan, 0};
  const FrameDev fnew = a.frames[a.fnew];
  const GeoDev& geo = a.geo[f.frame_id];

  if (!(geo.baseline < P.min_baseline)) {  // :1320-1324: too little parallax -> feature untouched
    EpiGeo epigeo;
    epigeo.init(a.K, a.Kinv);
    epigeo.load_table(geo);
    // ---------------- trackFeature ----------------
    int stage = 0;  // 0 = tracked; otherwise where it stopped
    float flow_x = qnan, flow_y = qnan;
    float ucx, ucy, idepth_cmp, var_cmp;
    if (!filter_predict(epigeo, P.fparams.process_var_factor, f.x, f.y, f.idepth_mu, f.idepth_var, &ucx, &ucy,
                        &idepth_cmp, &var_cmp)) {  // :1553
      stage = 1;
    } else {
      const int row_offset = P.do_letterbox ? a.height / 3 : 0;
      const int border = (int)(P.rescale_factor_max * P.fparams.win_size / 2 + 1);  // :1576
      const int vr_x = border, vr_y = border + row_offset;
      const int vr_w = a.width - 2 * border, vr_h = a.height - 2 * border - 2 * row_offset;
      float rescale_factor = 1.0f;
      if ((f.idepth_mu > 0.0f) && (idepth_cmp > 0.0f)) rescale_factor = idepth_cmp / f.idepth_mu;

      if ((rescale_factor <= P.rescale_factor_min) || (rescale_factor >= P.rescale_factor_max)) {
        // :1604-1659 the patch scale changed too much: move the feature to the current pose-frame
        EpiGeo epipf;
        epipf.init(a.K, a.Kinv);
        epipf.load_table(a.geo[FB_MAX_FRAMES + f.frame_id]);
        float upx, upy, idepth_pf, var_pf;
        const bool moved = filter_predict(epipf, P.fparams.process_var_factor, f.x, f.y, f.idepth_mu,
                                          f.idepth_var, &upx, &upy, &idepth_pf, &var_pf);
        if (!moved || !rect_contains(vr_x, vr_y, vr_w, vr_h, upx, upy)) {
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
        if (!filter_get_search_region(P.fparams, epigeo, a.width, a.height, f.x, f.y, f.idepth_mu, f.idepth_var,
                                      &sx, &sy, &ex, &ey)) {  // :1663
          stage = 4;
        } else if (!rect_contains(vr_x, vr_y, vr_w, vr_h, f.x, f.y)) {
          stage = 5;
        } else {
          // ---- inverse_depth_filter::search (.cc:184-266) in padded image coordinates (:1677-1678) ----
          const float off = (float)a.pad;
          const float urx = f.x + off, ury = f.y + off;
          sx += off; sy += off; ex += off; ey += off;
          // stage the bounding patch of the walk while the reference patch is gathered: the taps reach from
          // start - 2 inc to end + 2 inc (+1 px for the bilinear footprint, +1 px of rounding slack)
          float incx = ex - sx, incy = ey - sy;
          const float epl = sqrt_rn(incx * incx + incy * incy);
          incx *= P.fparams.sparams.sample_dist / epl;
          incy *= P.fparams.sparams.sample_dist / epl;
          const float ax = sx - 2.0f * incx, ay = sy - 2.0f * incy, bx = ex + 2.0f * incx, by = ey + 2.0f * incy;
          const int tx0 = ((int)fminf(ax, bx) - 1) & ~15, ty0 = (int)fminf(ay, by) - 1;
          PatchView view;
          view.base = fnew.img;  // the image itself unless the walk fits the staged patch
          view.pitch = a.pitch;
          view.x0 = 0;
          view.y0 = 0;
          const bool staged = a.stage_patches && (int)fmaxf(ax, bx) + 2 - tx0 < FB_TILE_W &&
                              (int)fmaxf(ay, by) + 2 - ty0 < FB_TILE_H && tx0 >= 0 && ty0 >= 0;
          if (staged && lane == 0) patch_load_issue(&tmap_cmp, s_tile[w], &s_bar[w], tx0, ty0);
          // reference patch: 5 samples along the epipolar line of the parent image (lanes 0..4)
          float erx, ery;
          epigeo.reference_epiline(urx, ury, &erx, &ery);
          const FrameDev fref = a.frames[f.frame_id];
          float rv = 0.0f;
          if (lane < 5) {
            float px = urx, py = ury;
            if (lane == 0) { px = urx - 2 * erx * rescale_factor; py = ury - 2 * ery * rescale_factor; }
            if (lane == 1) { px = urx - erx * rescale_factor; py = ury - ery * rescale_factor; }
            if (lane == 3) { px = urx + erx * rescale_factor; py = ury + ery * rescale_factor; }
            if (lane == 4) { px = urx + 2 * erx * rescale_factor; py = ury + 2 * ery * rescale_factor; }
            rv = bilinear_u8(fref.img, a.pitch, px, py);
          }
          float ref[5];
#pragma unroll
          for (int i = 0; i < 5; ++i) ref[i] = __shfl_sync(0xffffffffu, rv, i);
          float ref_grad = 0.0f;
#pragma unroll
          for (int i = 1; i < 5; ++i) {
            const float g = fast_absf(ref[i] - ref[i - 1]);
            if (g > ref_grad) ref_grad = g;
          }
          __syncwarp();
          if (staged) {  // the patch is consumed (or at least awaited) on every path: the CTA may not retire under it
            patch_load_wait(&s_bar[w]);
            view.base = s_tile[w];
            view.pitch = FB_TILE_W;
            view.x0 = tx0;
            view.y0 = ty0;
          }
          int status;
          if (ref_grad < P.fparams.min_grad_mag) {
            status = 1;  // FAIL_REF_PATCH_GRADIENT
          } else {
            const int r = warp_line_match(P.fparams.sparams, rescale_factor, ref, view, sx, sy, ex, ey, s_ws[w],
                                          &ucx, &ucy);
            status = (r == 0) ? 0 : (r == 1 ? 2 : 3);  // SUCCESS / FAIL_AMBIGUOUS_MATCH / FAIL_MAX_COST
          }
          f.search_status = status;
          if (status != 0) {
            stage = 6;
          } else {
            flow_x = ucx - off;
            flow_y = ucy - off;
          }
        }
      }
    }

    if (stage != 0) {
      feature_fail(P, f);
      dg.stage = stage;
    } else {
      dg.flow_x = flow_x;
      dg.flow_y = flow_y;
      // ---------------- measurement (:1380-1390) and fusion (:1436) ----------------
      float mu_meas, var_meas, mu_post, var_post;
      const bool sensed = meas_idepth(P.zparams, epigeo, fnew.gx, fnew.gy, a.pitch, f.x, f.y, flow_x, flow_y,
                                      &mu_meas, &var_meas);
      dg.mu_meas = mu_meas;
      dg.var_meas = var_meas;
      if (!sensed) {
        feature_fail(P, f);
        dg.stage = 7;
      } else if (!filter_update(f.idepth_mu, f.idepth_var, mu_meas, var_meas, &mu_post, &var_post,
                                P.outlier_sigma_thresh)) {
        feature_fail(P, f);
        dg.stage = 8;
      } else {
        f.idepth_mu = P.do_meas_fusion ? mu_post : mu_meas;
        f.idepth_var = P.do_meas_fusion ? var_post : var_meas;
        f.valid = 1;
        f.num_updates++;
        f.num_dropouts = 0;
        dg.stage = 9;
        if (lane == 0) atomicAdd(a.count, 1);
      }
    }
  }
  if (lane == 0) {
    a.feats_out[ii] = f;
    if (a.diag) a.diag[ii] = dg;
  }
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
k_frame_prep(const unsigned char* __restrict__ img, int width, int height, int border, int pitch,
             unsigned char* __restrict__ img_pad, float* __restrict__ gx_pad, float* __restrict__ gy_pad) {
  const int pw = width + 2 * border, ph = height + 2 * border;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= pw || y >= ph) return;
  const int sx = x - border, sy = y - border;
  const size_t o = (size_t)y * pitch + x;
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
    std::memset(s->tmap, 0, sizeof(s->tmap));
    *st = s;
  }
  return static_cast<StereoState*>(*st);
}

int stereo_set_camera(void** state, int width, int height, const float K[9], const float Kinv[9]) {
  StereoState* s = get_state(state);
  if (width != s->width || height != s->height) {
    // frame slots uploaded under another image size no longer match: they have to be uploaded again
    for (int i = 0; i < FB_MAX_FRAMES; ++i) s->h_frames[i].valid = 0;
    s->pad = 0;
  }
  s->width = width;
  s->height = height;
  std::memcpy(s->K, K, sizeof(s->K));
  std::memcpy(s->Kinv, Kinv, sizeof(s->Kinv));
  return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point table (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int encode_image_map(CUtensorMap* out, unsigned char* img, int valid_w, int rows, int pitch) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    FB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    if (!p || q != cudaDriverEntryPointSuccess) {
      set_error("cuTensorMapEncodeTiled is not available in this driver");
      return FLAME_B200_ECUDA;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)valid_w, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch};
  const cuuint32_t box[2] = {FB_TILE_W, FB_TILE_H};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, img, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return FLAME_B200_ECUDA;
  }
  return 0;
}

// Validates the slot / padding, sizes the slot's pitched buffers (pitch = padded width rounded up to 64
// bytes: TMA needs 16-byte row strides) and (re)builds its TMA descriptor.
static int frame_slot_prepare(StereoState* s, int fslot, int pad) {
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
  s->pitch = (s->width + 2 * pad + 63) / 64 * 64;
  const int rows = s->height + 2 * pad;
  const size_t npix = (size_t)s->pitch * rows;
  if (npix > s->cap_pix[fslot]) {
    s->h_frames[fslot].valid = 0;
    s->cap_pix[fslot] = 0;
    for (void** p : {(void**)&s->d_img[fslot], (void**)&s->d_gx[fslot], (void**)&s->d_gy[fslot]}) {
      if (*p) FB_CUDA(cudaFree(*p));
      *p = nullptr;
    }
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_img[fslot]), npix));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_gx[fslot]), npix * 4));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_gy[fslot]), npix * 4));
    s->cap_pix[fslot] = npix;
    s->slot_w[fslot] = 0;
  }
  if (s->slot_w[fslot] != s->width || s->slot_h[fslot] != s->height || s->slot_pad[fslot] != pad) {
    if (int rc = encode_image_map(&s->tmap[fslot], s->d_img[fslot], s->width + 2 * pad, rows, s->pitch)) return rc;
    s->slot_w[fslot] = s->width;
    s->slot_h[fslot] = s->height;
    s->slot_pad[fslot] = pad;
  }
  if (!s->d_frames) FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_frames), sizeof(s->h_frames)));
  if (!s->stage_free) FB_CUDA(cudaEventCreateWithFlags(&s->stage_free, cudaEventDisableTiming));
  return 0;
}

static int stage_reserve(StereoState* s, size_t need) {
  FB_CUDA(cudaEventSynchronize(s->stage_free));
  if (need > s->h_stage_cap) {
    s->h_stage_cap = 0;
    if (s->h_stage) FB_CUDA(cudaFreeHost(s->h_stage));
    s->h_stage = nullptr;
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_stage), need));
    s->h_stage_cap = need;
  }
  return 0;
}

static void frame_publish(StereoState* s, int fslot, const float pose[7], unsigned char* hs_frames) {
  FrameDev& fr = s->h_frames[fslot];
  fr.img = s->d_img[fslot];
  fr.gx = s->d_gx[fslot];
  fr.gy = s->d_gy[fslot];
  std::memcpy(fr.pose, pose, sizeof(fr.pose));
  fr.valid = 1;
  std::memcpy(hs_frames, s->h_frames, sizeof(s->h_frames));
}

int stereo_frame_upload(void** state, cudaStream_t st, int fslot, const uint8_t* img_pad,
                        const float* gradx_pad, const float* grady_pad, int pad, const float pose[7]) {
  StereoState* s = get_state(state);
  if (int rc = frame_slot_prepare(s, fslot, pad)) return rc;
  const int pw = s->width + 2 * pad, rows = s->height + 2 * pad;
  const size_t n = (size_t)pw * rows;
  if (int rc = stage_reserve(s, n * 9 + sizeof(s->h_frames))) return rc;
  unsigned char* hs = s->h_stage;
  std::memcpy(hs, gradx_pad, n * 4);
  std::memcpy(hs + n * 4, grady_pad, n * 4);
  std::memcpy(hs + n * 8, img_pad, n);
  frame_publish(s, fslot, pose, hs + n * 9);
  const size_t pitch = (size_t)s->pitch;
  FB_CUDA(cudaMemcpy2DAsync(s->d_gx[fslot], pitch * 4, hs, (size_t)pw * 4, (size_t)pw * 4, rows,
                            cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemcpy2DAsync(s->d_gy[fslot], pitch * 4, hs + n * 4, (size_t)pw * 4, (size_t)pw * 4, rows,
                            cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemcpy2DAsync(s->d_img[fslot], pitch, hs + n * 8, (size_t)pw, (size_t)pw, rows,
                            cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemcpyAsync(s->d_frames, hs + n * 9, sizeof(s->h_frames), cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaEventRecord(s->stage_free, st));
  return 0;
}

// Raw (unpadded) u8 image in, padded image + gradients computed on the device.
int stereo_frame_upload_raw(void** state, cudaStream_t st, int fslot, const uint8_t* img, int pad,
                            const float pose[7]) {
  StereoState* s = get_state(state);
  if (int rc = frame_slot_prepare(s, fslot, pad)) return rc;
  const size_t nraw = (size_t)s->width * s->height;
  if (int rc = stage_reserve(s, nraw + sizeof(s->h_frames))) return rc;
  if (nraw > s->cap_raw) {
    s->cap_raw = 0;
    if (s->d_raw) FB_CUDA(cudaFree(s->d_raw));
    s->d_raw = nullptr;
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_raw), nraw));
    s->cap_raw = nraw;
  }
  std::memcpy(s->h_stage, img, nraw);
  frame_publish(s, fslot, pose, s->h_stage + nraw);
  FB_CUDA(cudaMemcpyAsync(s->d_raw, s->h_stage, nraw, cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemcpyAsync(s->d_frames, s->h_stage + nraw, sizeof(s->h_frames), cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaEventRecord(s->stage_free, st));
  const int pw = s->width + 2 * pad, ph = s->height + 2 * pad;
  k_frame_prep<<<dim3((pw + 255) / 256, ph), 256, 0, st>>>(s->d_raw, s->width, s->height, pad, s->pitch,
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
  const size_t pw = (size_t)(s->width + 2 * s->pad), rows = (size_t)(s->height + 2 * s->pad), pitch = (size_t)s->pitch;
  if (img_pad) FB_CUDA(cudaMemcpy2DAsync(img_pad, pw, s->d_img[fslot], pitch, pw, rows, cudaMemcpyDeviceToHost, st));
  if (gradx_pad)
    FB_CUDA(cudaMemcpy2DAsync(gradx_pad, pw * 4, s->d_gx[fslot], pitch * 4, pw * 4, rows, cudaMemcpyDeviceToHost, st));
  if (grady_pad)
    FB_CUDA(cudaMemcpy2DAsync(grady_pad, pw * 4, s->d_gy[fslot], pitch * 4, pw * 4, rows, cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// ---- the resident feature set -------------------------------------------------------------------
static int feats_reserve(StereoState* s, int n_feats) {
  if (n_feats > s->cap_feats) {
    const int n = n_feats + n_feats / 2;
    s->cap_feats = 0;
    s->n_feats = 0;
    for (void** p : {(void**)&s->d_feats, (void**)&s->d_feats_out, (void**)&s->d_diag}) {
      if (*p) FB_CUDA(cudaFree(*p));
      *p = nullptr;
    }
    if (s->h_feats) FB_CUDA(cudaFreeHost(s->h_feats));
    if (s->h_diag) FB_CUDA(cudaFreeHost(s->h_diag));
    s->h_feats = nullptr;
    s->h_diag = nullptr;
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_feats), (size_t)n * sizeof(flame_b200_feature)));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_feats_out), (size_t)n * sizeof(flame_b200_feature)));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_diag), (size_t)n * sizeof(flame_b200_track_diag)));
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_feats), (size_t)n * sizeof(flame_b200_feature)));
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_diag), (size_t)n * sizeof(flame_b200_track_diag)));
    s->cap_feats = n;
  }
  if (!s->d_count) {
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_count), sizeof(int)));
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_count), sizeof(int)));
  }
  if (!s->d_geo) FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_geo), 2 * FB_MAX_FRAMES * sizeof(GeoDev)));
  if (!s->feats_free) FB_CUDA(cudaEventCreateWithFlags(&s->feats_free, cudaEventDisableTiming));
  return 0;
}

int stereo_features_upload(void* state, cudaStream_t st, const flame_b200_feature* feats, int32_t n_feats) {
  StereoState* s = static_cast<StereoState*>(state);
  if (!s || s->width <= 0) {
    set_error("features_upload: camera / frames not set");
    return FLAME_B200_ESTATE;
  }
  for (int i = 0; i < n_feats; ++i) {
    if (feats[i].frame_id >= FB_MAX_FRAMES || !s->h_frames[feats[i].frame_id].valid) {
      set_error("update_feature_idepths: feature parented to a frame slot that was not uploaded");
      return FLAME_B200_EINVAL;
    }
    // the kernel addresses the images with these: a non-finite coordinate must not reach it
    if (!std::isfinite(feats[i].x) || !std::isfinite(feats[i].y) || !std::isfinite(feats[i].idepth_mu) ||
        !std::isfinite(feats[i].idepth_var)) {
      set_error("update_feature_idepths: feature with a non-finite position / idepth");
      return FLAME_B200_EINVAL;
    }
  }
  if (int rc = feats_reserve(s, n_feats)) return rc;
  FB_CUDA(cudaEventSynchronize(s->feats_free));
  s->n_feats = n_feats;
  s->tracked = false;
  if (n_feats == 0) return 0;
  std::memcpy(s->h_feats, feats, (size_t)n_feats * sizeof(flame_b200_feature));
  FB_CUDA(cudaMemcpyAsync(s->d_feats, s->h_feats, (size_t)n_feats * sizeof(flame_b200_feature),
                          cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaEventRecord(s->feats_free, st));
  return 0;
}

// Asynchronous: resident set (in) -> results (out) + diagnostics + update count, all on the device.
int stereo_features_track(void* state, cudaStream_t st, const flame_b200_track_params* params, int fnew,
                          int curr_pf) {
  StereoState* s = static_cast<StereoState*>(state);
  if (!s || s->width <= 0 || !s->d_feats) {
    set_error("features_track: camera / frames / features not set");
    return FLAME_B200_ESTATE;
  }
  if (fnew < 0 || fnew >= FB_MAX_FRAMES || curr_pf < 0 || curr_pf >= FB_MAX_FRAMES ||
      !s->h_frames[fnew].valid || !s->h_frames[curr_pf].valid) {
    set_error("update_feature_idepths: frame slot not uploaded");
    return FLAME_B200_EINVAL;
  }
  const flame_b200_filter_params& fp = params->fparams;
  if (!(fp.sparams.sample_dist > 0.0f) ||
      !(std::fmax(fp.epilength_max, fp.epilength_min) / fp.sparams.sample_dist + 4.0f <= (float)FB_SEARCH_MAX_STEPS)) {
    set_error("update_feature_idepths: epilength / sample_dist allows more than " +
              std::to_string(FB_SEARCH_MAX_STEPS - 4) + " SSD positions per feature");
    return FLAME_B200_EINVAL;
  }
  s->tracked = true;
  FB_CUDA(cudaMemsetAsync(s->d_count, 0, sizeof(int), st));
  if (s->n_feats == 0) return 0;
  StereoArgs a;
  a.P = *params;
  std::memcpy(a.K, s->K, sizeof(a.K));
  std::memcpy(a.Kinv, s->Kinv, sizeof(a.Kinv));
  a.width = s->width;
  a.height = s->height;
  a.pad = s->pad;
  a.pitch = s->pitch;
  a.fnew = fnew;
  a.curr_pf = curr_pf;
  a.frames = s->d_frames;
  a.geo = s->d_geo;
  a.feats = s->d_feats;
  a.feats_out = s->d_feats_out;
  a.diag = s->d_diag;
  a.n = s->n_feats;
  a.count = s->d_count;
  a.stage_patches = 1;
  if (const char* sv = getenv("FLAME_B200_SEARCH_TMA")) a.stage_patches = atoi(sv);
  k_epigeo_table<<<2, FB_MAX_FRAMES, 0, st>>>(a, s->d_geo);
  k_track_features<<<(s->n_feats + TRACK_WARPS - 1) / TRACK_WARPS, 32 * TRACK_WARPS, 0, st>>>(a, s->tmap[fnew]);
  count_launch(2);
  FB_CUDA(cudaGetLastError());
  return 0;
}

// Synchronises. commit != 0: the results become the resident set (the next track continues from them).
int stereo_features_download(void* state, cudaStream_t st, flame_b200_feature* feats,
                             flame_b200_track_diag* diag, int32_t* n_updates, int commit) {
  StereoState* s = static_cast<StereoState*>(state);
  if (!s || !s->d_feats) {
    set_error("features_download: no resident feature set");
    return FLAME_B200_ESTATE;
  }
  const int n = s->n_feats;
  const flame_b200_feature* src = s->tracked ? s->d_feats_out : s->d_feats;
  if (n > 0 && feats)
    FB_CUDA(cudaMemcpyAsync(s->h_feats, src, (size_t)n * sizeof(flame_b200_feature), cudaMemcpyDeviceToHost, st));
  if (n > 0 && diag && s->tracked)
    FB_CUDA(cudaMemcpyAsync(s->h_diag, s->d_diag, (size_t)n * sizeof(flame_b200_track_diag),
                            cudaMemcpyDeviceToHost, st));
  if (n_updates && s->tracked)
    FB_CUDA(cudaMemcpyAsync(s->h_count, s->d_count, sizeof(int), cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaEventRecord(s->feats_free, st));
  FB_CUDA(cudaStreamSynchronize(st));
  if (n > 0 && feats) std::memcpy(feats, s->h_feats, (size_t)n * sizeof(flame_b200_feature));
  if (n > 0 && diag && s->tracked) std::memcpy(diag, s->h_diag, (size_t)n * sizeof(flame_b200_track_diag));
  if (n_updates) *n_updates = s->tracked ? *s->h_count : 0;
  if (commit && s->tracked) {
    std::swap(s->d_feats, s->d_feats_out);
    s->tracked = false;
  }
  return 0;
}

int stereo_features_count(void* state) {
  StereoState* s = static_cast<StereoState*>(state);
  return s ? s->n_feats : 0;
}

int stereo_update(void* state, cudaStream_t st, const flame_b200_track_params* params, int fnew,
                  int curr_pf, flame_b200_feature* feats, int32_t n_feats,
                  flame_b200_track_diag* diag, int32_t* n_updates) {
  if (n_updates) *n_updates = 0;
  if (int rc = stereo_features_upload(state, st, feats, n_feats)) return rc;
  if (int rc = stereo_features_track(state, st, params, fnew, curr_pf)) return rc;
  if (n_feats == 0) return 0;
  return stereo_features_download(state, st, feats, diag, n_updates, 1);
}

// ------------------------------------------------------------------------------------------
// Flame::projectGraph, per-vertex body (flame.cc:1892-1899): pos, x <- EpipolarGeometry::project
// ------------------------------------------------------------------------------------------
struct ProjectArgs {
  float K[9], Kinv[9], q[4], t[3];
  float scale;
  // removal rules of Flame::projectGraph (flame.cc:1872-1917); marks == nullptr: projection only
  unsigned char* marks;       // [V] caller order: 1 = the reference would put the vertex into vtx_to_remove
  float rx, ry, rw, rh;       // cv::Rect_<float> valid_region
  int grad_check;             // params.do_grad_check_after_projection
  float grad_thresh2;         // params.min_grad_mag^2
  const float* gx;            // padded gradients of the new frame (nullable unless grad_check)
  const float* gy;
  int pitch, pad;
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
  if (a.marks) {
    // cv::Rect_<float>::contains(Point2f): x <= pt.x < x + width, y <= pt.y < y + height (:1901)
    const bool inside = a.rx <= cx && cx < a.rx + a.rw && a.ry <= cy && cy < a.ry + a.rh;
    bool remove = !inside || (idn < 0.0f);
    if (!remove && a.grad_check) {
      // bilinearInterp<float, float> on the UNPADDED gradient images (:1908-1911): same weights, the taps
      // shifted by the padding of the device copies
      const int xf = (int)cx, yf = (int)cy;
      const float dx = cx - xf, dy = cy - yf;
      const float w11 = dx * dy, w01 = dx - w11, w10 = dy - w11, w00 = 1.0f - dx - dy + w11;
      const size_t at = (size_t)(yf + a.pad) * a.pitch + (xf + a.pad);
      const float gx = w00 * a.gx[at] + w01 * a.gx[at + 1] + w10 * a.gx[at + a.pitch] + w11 * a.gx[at + 1 + a.pitch];
      const float gy = w00 * a.gy[at] + w01 * a.gy[at + 1] + w10 * a.gy[at + a.pitch] + w11 * a.gy[at + 1 + a.pitch];
      if (gx * gx + gy * gy < a.grad_thresh2) remove = true;
    }
    a.marks[o] = remove ? 1 : 0;
  }
}

// rules == nullptr: projection only. d_marks: V bytes of device scratch.
int stereo_graph_project(void* state, cudaStream_t st, const GraphDev& g, float* d_pos,
                         const float q[4], const float t[3], float scale, float* d_out,
                         const flame_b200_project_rules* rules, int fnew, unsigned char* d_marks) {
  StereoState* s = static_cast<StereoState*>(state);
  if (!s || s->width <= 0) {
    set_error("graph_project: camera not set (flame_b200_set_camera)");
    return FLAME_B200_ESTATE;
  }
  ProjectArgs a;
  std::memset(&a, 0, sizeof(a));
  std::memcpy(a.K, s->K, sizeof(a.K));
  std::memcpy(a.Kinv, s->Kinv, sizeof(a.Kinv));
  std::memcpy(a.q, q, sizeof(a.q));
  std::memcpy(a.t, t, sizeof(a.t));
  a.scale = scale;
  if (rules) {
    if (rules->do_grad_check_after_projection &&
        (fnew < 0 || fnew >= FB_MAX_FRAMES || !s->h_frames[fnew].valid)) {
      set_error("graph_project: the gradient check needs the new frame's slot (flame_b200_frame_upload)");
      return FLAME_B200_EINVAL;
    }
    const int row_offset = rules->do_letterbox ? s->height / 3 : 0;
    const int border = (int)(rules->rescale_factor_max * rules->win_size / 2 + 1);  // :1881
    a.marks = d_marks;
    a.rx = (float)border;
    a.ry = (float)(border + row_offset);
    a.rw = (float)(s->width - 2 * border);
    a.rh = (float)(s->height - 2 * border - 2 * row_offset);
    a.grad_check = rules->do_grad_check_after_projection;
    a.grad_thresh2 = rules->min_grad_mag * rules->min_grad_mag;
    if (a.grad_check) {
      a.gx = s->d_gx[fnew];
      a.gy = s->d_gy[fnew];
    }
    a.pitch = s->pitch;
    a.pad = s->pad;
  }
  if (g.V == 0) return 0;
  k_graph_project<<<(g.V + 127) / 128, 128, 0, st>>>(g, d_pos, a, d_out);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Flame::detectFeatures (flame.cc:1192-1277): on a grid of win x win cells, the pixel with the largest
// squared image gradient along the epipolar direction (prev <- ref geometry), among the pixels whose
// gradient magnitude and epipolar gradient both exceed min_grad_mag; cells that already hold a feature
// are masked. One CTA per cell; ties go to the LAST pixel in row-major scan order (the reference's `>=`).
// ------------------------------------------------------------------------------------------
struct DetectArgs {
  float K[9], Kinv[9];
  float pose_ref[7], pose_prev[7];
  const float* gx;  // padded gradients of the reference frame
  const float* gy;
  int width, height, pitch, pad;
  int border, row_offset, win, wclvl, hclvl;
  float thresh2;
  unsigned long long* best;   // [hclvl * wclvl] (epigrad^2 bits << 32) | scan index, 0 = none
  unsigned char* mask;        // [hclvl * wclvl] 255 = free, 0 = already holds a feature
  const float* curr;          // [2 n_curr] current features
  int n_curr;
  float* out;                 // [2 max_out] detections (x, y)
  int max_out;
  int* n_out;
};

__global__ void k_detect_mask(DetectArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_curr) return;
  // uint32_t x_clvl = curr_feats[ii].x / win_size (:1201-1202): float division, truncated
  const unsigned xc = (unsigned)(a.curr[2 * i] / (float)a.win), yc = (unsigned)(a.curr[2 * i + 1] / (float)a.win);
  if (xc < (unsigned)a.wclvl && yc < (unsigned)a.hclvl) a.mask[yc * a.wclvl + xc] = 0;
}

__global__ void __launch_bounds__(256) k_detect_cells(DetectArgs a) {
  __shared__ float s_t[3];  // t_cmp_to_ref of the prev <- ref geometry: all referenceEpiline needs besides K
  __shared__ unsigned long long s_best[8];
  const int cj = blockIdx.x, ci = blockIdx.y;
  if (threadIdx.x == 0) {
    float T[7];
    se3_inv_mul(a.pose_prev, a.pose_ref, T);  // T_ref_to_prev = fprev.pose^-1 * fref.pose (:1207)
    EpiGeo e;
    e.init(a.K, a.Kinv);
    e.load(T, T + 4);
    s_t[0] = e.t_cmp_to_ref[0];
    s_t[1] = e.t_cmp_to_ref[1];
    s_t[2] = e.t_cmp_to_ref[2];
  }
  __syncthreads();
  EpiGeo e;
  e.init(a.K, a.Kinv);
  e.t_cmp_to_ref[0] = s_t[0];
  e.t_cmp_to_ref[1] = s_t[1];
  e.t_cmp_to_ref[2] = s_t[2];
  unsigned long long best = 0ull;
  const int i_lo = max(ci * a.win, a.border + a.row_offset), i_hi = min((ci + 1) * a.win, a.height - a.border - a.row_offset);
  const int j_lo = max(cj * a.win, a.border), j_hi = min((cj + 1) * a.win, a.width - a.border);
  const int w = j_hi - j_lo, n = (i_hi > i_lo && w > 0) ? (i_hi - i_lo) * w : 0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const int ii = i_lo + k / w, jj = j_lo + k % w;
    // the cell of a pixel is static_cast<float>(ii) / win_size truncated (:1243-1244) = ii / win for these sizes
    const size_t at = (size_t)(ii + a.pad) * a.pitch + (jj + a.pad);
    const float gx = a.gx[at], gy = a.gy[at];
    if (gx * gx + gy * gy < a.thresh2) continue;
    float ex, ey;
    e.reference_epiline((float)ii, (float)jj, &ex, &ey);  // sic: cv::Point2f(ii, jj), row first (:1231)
    const float epigrad = gx * ex + gy * ey;
    const float epigrad2 = epigrad * epigrad;
    if (epigrad2 < a.thresh2 || epigrad2 != epigrad2) continue;  // NaN (pixel at the epipole) never wins `>=`
    const unsigned long long key = ((unsigned long long)__float_as_uint(epigrad2) << 32) | (unsigned)(ii * a.width + jj);
    best = key > best ? key : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long v = __shfl_xor_sync(0xffffffffu, best, o);
    best = v > best ? v : best;
  }
  if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) best = s_best[k] > best ? s_best[k] : best;
    a.best[ci * a.wclvl + cj] = best;
  }
}

// one CTA: detections in row-major cell order (:1262-1270)
__global__ void __launch_bounds__(1024) k_detect_compact(DetectArgs a) {
  __shared__ int s_warp[32];
  __shared__ int s_run, s_tot;
  const int n = a.hclvl * a.wclvl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (int b = 0; b < n; b += 1024) {
    const int i = b + threadIdx.x;
    unsigned long long key = 0ull;
    bool sel = false;
    if (i < n) {
      key = a.best[i];
      // cmask > 0 && best_gradsc > 0 (:1266)
      sel = a.mask[i] > 0 && __uint_as_float((unsigned)(key >> 32)) > 0.0f;
    }
    const unsigned m = __ballot_sync(0xffffffffu, sel);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      const int v = s_warp[lane];
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
      }
      s_warp[lane] = inc - v;  // exclusive offset of every warp
      if (lane == 31) s_tot = inc;
    }
    __syncthreads();
    const int pos = s_run + s_warp[warp] + __popc(m & ((1u << lane) - 1u));
    if (sel && pos < a.max_out) {
      const unsigned idx = (unsigned)(key & 0xffffffffu);
      a.out[2 * pos] = (float)(idx % (unsigned)a.width);      // cv::Point2f(jj, ii)
      a.out[2 * pos + 1] = (float)(idx / (unsigned)a.width);
    }
    __syncthreads();
    if (threadIdx.x == 0) s_run += s_tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *a.n_out = s_run;
}

int stereo_detect(void* state, cudaStream_t st, const flame_b200_detect_params* p, int fref, int fprev,
                  const float* curr_xy, int n_curr, float* out_xy, int max_out, int32_t* n_out) {
  StereoState* s = static_cast<StereoState*>(state);
  if (!s || s->width <= 0) {
    set_error("detect_features: camera / frames not set");
    return FLAME_B200_ESTATE;
  }
  if (fref < 0 || fref >= FB_MAX_FRAMES || fprev < 0 || fprev >= FB_MAX_FRAMES || !s->h_frames[fref].valid ||
      !s->h_frames[fprev].valid || p->detection_win_size < 1 || max_out < 0 || n_curr < 0) {
    set_error("detect_features: bad arguments / frame slot not uploaded");
    return FLAME_B200_EINVAL;
  }
  DetectArgs a;
  std::memset(&a, 0, sizeof(a));
  std::memcpy(a.K, s->K, sizeof(a.K));
  std::memcpy(a.Kinv, s->Kinv, sizeof(a.Kinv));
  std::memcpy(a.pose_ref, s->h_frames[fref].pose, sizeof(a.pose_ref));
  std::memcpy(a.pose_prev, s->h_frames[fprev].pose, sizeof(a.pose_prev));
  a.gx = s->d_gx[fref];
  a.gy = s->d_gy[fref];
  a.width = s->width;
  a.height = s->height;
  a.pitch = s->pitch;
  a.pad = s->pad;
  a.row_offset = p->do_letterbox ? s->height / 3 : 0;
  a.border = (int)(p->rescale_factor_max * p->win_size / 2 + 1);  // :848
  a.win = p->detection_win_size;
  a.hclvl = (s->height + a.win - 1) / a.win;  // fast_ceil(height / win)
  a.wclvl = (s->width + a.win - 1) / a.win;
  a.thresh2 = p->min_grad_mag * p->min_grad_mag;
  const size_t cells = (size_t)a.hclvl * a.wclvl;
  // device scratch: [best cells u64 | n_out int (+pad) | out 2*max_out f32 | curr 2*n_curr f32 | mask cells u8]
  const size_t need = cells * 8 + 16 + (size_t)max_out * 8 + (size_t)n_curr * 8 + cells;
  if (need > s->detect_cap) {
    s->detect_cap = 0;
    if (s->d_detect) FB_CUDA(cudaFree(s->d_detect));
    s->d_detect = nullptr;
    if (s->h_detect) FB_CUDA(cudaFreeHost(s->h_detect));
    s->h_detect = nullptr;
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->d_detect), need));
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s->h_detect), need));
    s->detect_cap = need;
  }
  unsigned char* d = s->d_detect;
  a.best = reinterpret_cast<unsigned long long*>(d);
  a.n_out = reinterpret_cast<int*>(d + cells * 8);
  a.out = reinterpret_cast<float*>(d + cells * 8 + 16);
  float* d_curr = a.out + 2 * (size_t)max_out;
  a.curr = d_curr;
  a.n_curr = n_curr;
  a.mask = reinterpret_cast<unsigned char*>(d_curr + 2 * (size_t)n_curr);
  a.max_out = max_out;
  FB_CUDA(cudaStreamSynchronize(st));  // the pinned mirror is reused call to call
  if (n_curr > 0) {
    std::memcpy(s->h_detect, curr_xy, (size_t)n_curr * 8);
    FB_CUDA(cudaMemcpyAsync(d_curr, s->h_detect, (size_t)n_curr * 8, cudaMemcpyHostToDevice, st));
  }
  FB_CUDA(cudaMemsetAsync(a.mask, 255, cells, st));
  if (n_curr > 0) {
    k_detect_mask<<<(n_curr + 255) / 256, 256, 0, st>>>(a);
    count_launch();
  }
  k_detect_cells<<<dim3(a.wclvl, a.hclvl), 256, 0, st>>>(a);
  k_detect_compact<<<1, 1024, 0, st>>>(a);
  count_launch(2);
  FB_CUDA(cudaGetLastError());
  unsigned char* h = s->h_detect;
  FB_CUDA(cudaMemcpyAsync(h, a.n_out, 16 + (size_t)max_out * 8, cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  const int n = *reinterpret_cast<int*>(h);
  if (n_out) *n_out = n;
  if (out_xy && n > 0) std::memcpy(out_xy, h + 16, (size_t)std::min(n, max_out) * 8);
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
  if (s->d_detect) cudaFree(s->d_detect);
  if (s->h_detect) cudaFreeHost(s->h_detect);
  if (s->d_raw) cudaFree(s->d_raw);
  if (s->h_stage) cudaFreeHost(s->h_stage);
  if (s->stage_free) cudaEventDestroy(s->stage_free);
  if (s->d_feats) cudaFree(s->d_feats);
  if (s->d_feats_out) cudaFree(s->d_feats_out);
  if (s->d_geo) cudaFree(s->d_geo);
  if (s->feats_free) cudaEventDestroy(s->feats_free);
  if (s->d_diag) cudaFree(s->d_diag);
  if (s->d_count) cudaFree(s->d_count);
  if (s->h_feats) cudaFreeHost(s->h_feats);
  if (s->h_diag) cudaFreeHost(s->h_diag);
  if (s->h_count) cudaFreeHost(s->h_count);
  delete s;
}

}  // namespace fb
