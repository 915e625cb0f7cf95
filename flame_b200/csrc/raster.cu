// Dense per-pixel projection of the Delaunay mesh (inverse-depth map) for sm_100a.
//
// Replaces flame::utils::interpolateMesh (src/flame/utils/image_utils.cc:373-396) over
// flame::utils::DrawShadedTriangleBarycentric (src/flame/utils/rasterization.cc:164-245).
//
// Reference semantics reproduced exactly:
//  - vertices are rounded to integer pixels with round-half-to-even (cv::Point2f -> cv::Point);
//  - the triangle is drawn with its vertices REVERSED (Triangle emits CW, the rasteriser wants CCW);
//  - a pixel is covered when all three edge functions are >= 0 (inclusive: shared edges and
//    vertices are written by every covering triangle) — the LAST triangle in the list wins;
//  - value = (v1*w1 + (v2*w2 + v3*w3)) / (w1 + (w2 + w3)), one IEEE divide, this association;
//  - the reference sweeps the bounding box 4 pixels at a time and addresses the image LINEARLY
//    (p.y*cols + p.x), so a degenerate (zero-area) triangle, whose edge functions vanish on a
//    whole line, also writes NaN up to 3 pixels past xmax (possibly wrapping to the next row).
//
// Edge functions are integer-valued and exact in fp32 for images up to ~2896 px on a side
// (|A*x + B*y + C| < 2^24), so direct evaluation equals the reference's incremental stepping.
//
// Formulation: the reference scatters triangles in order with read-modify-write, so the LAST
// (highest-index) covering triangle wins. Here coverage and shading are split:
//   k_tri_setup    one thread per triangle: rounding, edge coefficients, sweep footprint
//   k_owner_small  one warp per triangle (16 x 2 pixel steps over its footprint): exact integer
//                  coverage test, atomicMax((epoch << 20) | triangle) into a per-pixel owner word
//   k_owner_big    one CTA per large-footprint triangle (hull slivers, texture-poor regions): rows
//                  over warps, each row clipped to a conservative span from the three edge functions
//   k_resolve      four pixels per thread: read the owner, shade that one triangle, write the pixel once
// The owner word carries the projection's epoch, so the buffer never needs clearing between
// projections (stale words lose every atomicMax and are ignored by k_resolve); when the 12-bit epoch
// is about to wrap k_resolve zeroes the words it reads. Triangle ids use 20 bits (T < 2^20).
#include <algorithm>

#include "internal.h"

namespace fb {

#define OWNER_T_BITS 20
#define EPOCH_MAX 4095u
#define BIG_AREA 1024   // sweep-footprint pixels above which a triangle goes to k_owner_big

__device__ __forceinline__ int min3(int a, int b, int c) { return min(a, min(b, c)); }
__device__ __forceinline__ int max3(int a, int b, int c) { return max(a, max(b, c)); }

// Edge::init (rasterization.h:129-153): A*x + B*y + C with separately rounded mul/add.
__device__ __forceinline__ float edge_at(float A, float B, float C, int x, int y) {
  return __fadd_rn(__fadd_rn(__fmul_rn(A, (float)x), __fmul_rn(B, (float)y)), C);
}

// Coverage + value of triangle `s` at sweep position (x, y); x may exceed cols-1 (4-px overshoot).
__device__ __forceinline__ bool tri_eval(const TriSetup& s, int x, int y, float* val) {
  if (y < s.ymin || y > s.ymax || x < s.xmin || x >= s.xend) return false;
  const float w1 = edge_at(s.A1, s.B1, s.C1, x, y);
  const float w2 = edge_at(s.A2, s.B2, s.C2, x, y);
  const float w3 = edge_at(s.A3, s.B3, s.C3, x, y);
  if (!(w1 >= 0.0f && w2 >= 0.0f && w3 >= 0.0f)) return false;
  const float norm = __fadd_rn(w1, __fadd_rn(w2, w3));
  const float num = __fadd_rn(__fmul_rn(s.v1, w1), __fadd_rn(__fmul_rn(s.v2, w2), __fmul_rn(s.v3, w3)));
  *val = __fdiv_rn(num, norm);
  return true;
}

__global__ void __launch_bounds__(256) k_tri_setup(const MeshDev* __restrict__ ms) {
  const MeshDev m = ms[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t == 0) {  // open a new epoch for this projection (read by the later kernels only)
    const unsigned e = (unsigned)m.counters[1];
    m.counters[1] = (int)(e >= EPOCH_MAX ? 1u : e + 1u);
  }
  if (t >= m.T) return;
  TriSetup s;
  s.flags = 0;
  s.fx1 = -1;
  s.pad0 = s.pad1 = 0;
  if ((m.tri_ok && !m.tri_ok[t]) || (m.tri_filt && !m.tri_filt[t])) {
    s.ymin = 1; s.ymax = 0; s.xmin = 0; s.xend = 0;  // empty footprint
    s.A1 = s.B1 = s.C1 = s.A2 = s.B2 = s.C2 = s.A3 = s.B3 = s.C3 = s.v1 = s.v2 = s.v3 = 0.0f;
    m.setup[t] = s;
    return;
  }
  const int ia = m.tris[3 * t], ib = m.tris[3 * t + 1], ic = m.tris[3 * t + 2];
  // image_utils.cc:385-391: p1 = vertices[tri[2]], p2 = vertices[tri[1]], p3 = vertices[tri[0]]
  const int p1x = __float2int_rn(m.pos[2 * ic]), p1y = __float2int_rn(m.pos[2 * ic + 1]);
  const int p2x = __float2int_rn(m.pos[2 * ib]), p2y = __float2int_rn(m.pos[2 * ib + 1]);
  const int p3x = __float2int_rn(m.pos[2 * ia]), p3y = __float2int_rn(m.pos[2 * ia + 1]);
  s.v1 = m.values[ic]; s.v2 = m.values[ib]; s.v3 = m.values[ia];
  s.xmin = min3(p1x, p2x, p3x); s.ymin = min3(p1y, p2y, p3y);
  const int xmax = max3(p1x, p2x, p3x);
  s.ymax = max3(p1y, p2y, p3y);
  s.xend = s.xmin + ((xmax - s.xmin) / 4 + 1) * 4;  // columns visited by the 4-wide sweep
  // w1 = e23, w2 = e31, w3 = e12 (rasterization.cc:177-179)
  s.A1 = (float)(p3y - p2y); s.B1 = (float)(p2x - p3x); s.C1 = (float)(p3x * p2y - p2x * p3y);
  s.A2 = (float)(p1y - p3y); s.B2 = (float)(p3x - p1x); s.C2 = (float)(p1x * p3y - p3x * p1y);
  s.A3 = (float)(p2y - p1y); s.B3 = (float)(p1x - p2x); s.C3 = (float)(p2x * p1y - p1x * p2y);
  const long long area2 = (long long)(p2x - p1x) * (p3y - p1y) - (long long)(p2y - p1y) * (p3x - p1x);
  const bool degenerate = (area2 == 0);
  // footprint in pixels: a non-degenerate triangle covers nothing outside [xmin, xmax]; a degenerate
  // one (edge functions vanish on a line) also covers the sweep's overshoot columns up to xend - 1
  const int fx0 = max(s.xmin, 0), fy0 = max(s.ymin, 0);
  const int fx1 = min(degenerate ? s.xend - 1 : xmax, m.cols - 1), fy1 = min(s.ymax, m.rows - 1);
  const bool wraps = degenerate && (s.xend > m.cols);
  s.fx1 = fx1;
  if (degenerate) s.flags |= 4;
  if (wraps) s.flags |= 2;
  const long long w = (wraps ? s.xend - 1 : fx1) - fx0 + 1, h = fy1 - fy0 + 1;
  if (w > 0 && h > 0 && w * h > BIG_AREA) {
    s.flags |= 8;
    m.big[atomicAdd(&m.counters[0], 1)] = t;
  }
  m.setup[t] = s;
}

// coverage at sweep position (x, y) -> owner word of the pixel the reference's linear addressing
// (p.y * cols + p.x) lands on
__device__ __forceinline__ void claim(const MeshDev& m, int x, int y, unsigned key) {
  if (x >= m.cols) {  // degenerate triangle whose 4-px sweep ran past the right border
    x -= m.cols;
    y += 1;
    if (y >= m.rows) return;
  }
  atomicMax(&m.owner[(size_t)y * m.cols + x], key);
}

__global__ void __launch_bounds__(256) k_owner_small(const MeshDev* __restrict__ ms) {
  const MeshDev m = ms[blockIdx.y];
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (t >= m.T) return;
  const int lane = threadIdx.x & 31;
  const TriSetup s = m.setup[t];
  if ((s.flags & 8) || s.ymin > s.ymax) return;
  const unsigned key = ((unsigned)m.counters[1] << OWNER_T_BITS) | (unsigned)t;
  const int fx0 = max(s.xmin, 0), fy0 = max(s.ymin, 0), fy1 = min(s.ymax, m.rows - 1);
  if (!(s.flags & 2)) {
    // 8 x 4 lanes; a lane walks down its column in steps of 4 rows, stepping the (integer-valued, hence
    // exactly representable) edge functions by 4*B
    const int fx1 = s.fx1;
    const int y0 = fy0 + (lane >> 3);
    const float fy = (float)y0;
    const float b1 = __fadd_rn(__fmul_rn(s.B1, fy), s.C1), b2 = __fadd_rn(__fmul_rn(s.B2, fy), s.C2),
                b3 = __fadd_rn(__fmul_rn(s.B3, fy), s.C3);
    const float s1 = __fmul_rn(s.B1, 4.0f), s2 = __fmul_rn(s.B2, 4.0f), s3 = __fmul_rn(s.B3, 4.0f);
    const size_t pstep = (size_t)4 * m.cols;
    for (int x = fx0 + (lane & 7); x <= fx1; x += 8) {
      const float fx = (float)x;
      float w1 = __fadd_rn(__fmul_rn(s.A1, fx), b1);
      float w2 = __fadd_rn(__fmul_rn(s.A2, fx), b2);
      float w3 = __fadd_rn(__fmul_rn(s.A3, fx), b3);
      unsigned int* ptr = m.owner + (size_t)y0 * m.cols + x;
      for (int y = y0; y <= fy1; y += 4) {
        if (w1 >= 0.0f && w2 >= 0.0f && w3 >= 0.0f) atomicMax(ptr, key);
        w1 = __fadd_rn(w1, s1);
        w2 = __fadd_rn(w2, s2);
        w3 = __fadd_rn(w3, s3);
        ptr += pstep;
      }
    }
    return;
  }
  // degenerate triangle whose sweep overshoot crosses the right border: generic positions + wrap
  const int fx1 = s.xend - 1;
  for (int y = fy0 + (lane >> 4); y <= fy1; y += 2) {
    for (int x = fx0 + (lane & 15); x <= fx1; x += 16) {
      const float w1 = edge_at(s.A1, s.B1, s.C1, x, y);
      const float w2 = edge_at(s.A2, s.B2, s.C2, x, y);
      const float w3 = edge_at(s.A3, s.B3, s.C3, x, y);
      if (w1 >= 0.0f && w2 >= 0.0f && w3 >= 0.0f) claim(m, x, y, key);
    }
  }
}

// conservative integer bounds of { x : A*x + d >= 0 } (exact test follows); false = empty row
__device__ __forceinline__ bool clip_span(float A, float d, int* lo, int* hi) {
  if (A == 0.0f) return d >= 0.0f;
  const float q = fminf(fmaxf(__fdiv_rn(-d, A), -1.0e6f), 1.0e6f);
  const int f = (int)floorf(q);
  if (A > 0.0f) *lo = max(*lo, f - 1);
  else *hi = min(*hi, f + 1);
  return true;
}

__global__ void __launch_bounds__(256) k_owner_big(const MeshDev* __restrict__ ms) {
  const MeshDev m = ms[blockIdx.y];
  const int n_big = m.counters[0];
  const unsigned epoch = (unsigned)m.counters[1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = blockIdx.x; b < n_big; b += gridDim.x) {
    const int t = m.big[b];
    const TriSetup s = m.setup[t];
    const unsigned key = (epoch << OWNER_T_BITS) | (unsigned)t;
    const int fx0 = max(s.xmin, 0), fy0 = max(s.ymin, 0), fy1 = min(s.ymax, m.rows - 1);
    const int fx1 = (s.flags & 2) ? s.xend - 1 : s.fx1;
    for (int y = fy0 + warp; y <= fy1; y += 8) {
      const float fy = (float)y;
      int lo = fx0, hi = fx1;
      if (!clip_span(s.A1, __fadd_rn(__fmul_rn(s.B1, fy), s.C1), &lo, &hi)) continue;
      if (!clip_span(s.A2, __fadd_rn(__fmul_rn(s.B2, fy), s.C2), &lo, &hi)) continue;
      if (!clip_span(s.A3, __fadd_rn(__fmul_rn(s.B3, fy), s.C3), &lo, &hi)) continue;
      for (int x = lo + lane; x <= hi; x += 32) {
        const float w1 = edge_at(s.A1, s.B1, s.C1, x, y);
        const float w2 = edge_at(s.A2, s.B2, s.C2, x, y);
        const float w3 = edge_at(s.A3, s.B3, s.C3, x, y);
        if (w1 >= 0.0f && w2 >= 0.0f && w3 >= 0.0f) claim(m, x, y, key);
      }
    }
  }
}

// Four consecutive (linear) pixels per thread: shade each with its owner triangle.
__global__ void __launch_bounds__(256) k_resolve(const MeshDev* __restrict__ ms, int fill_nan) {
  const MeshDev m = ms[blockIdx.y];
  const unsigned epoch = (unsigned)m.counters[1];
  const bool wipe = (epoch >= EPOCH_MAX);  // the next projection restarts at epoch 1
  const int n = m.rows * m.cols;
  if (blockIdx.x == 0 && threadIdx.x == 0) m.counters[0] = 0;  // re-arm the big list
  const float qnan = __uint_as_float(0x7fc00000u);  // quiet_NaN, flame.cc:412
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += gridDim.x * blockDim.x * 4) {
    unsigned o[4];
    const bool full = (i + 4 <= n);
    if (full) {
      const uint4 v = *reinterpret_cast<const uint4*>(m.owner + i);
      o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
      if (wipe) *reinterpret_cast<uint4*>(m.owner + i) = make_uint4(0u, 0u, 0u, 0u);
    } else {
      for (int k = 0; k < 4; ++k) {
        o[k] = (i + k < n) ? m.owner[i + k] : 0u;
        if (wipe && i + k < n) m.owner[i + k] = 0u;
      }
    }
    int py = i / m.cols, px = i - py * m.cols;
    float out[4];
    bool have[4];
    int last_t = -1;
    TriSetup s;
    float w1 = 0.0f, w2 = 0.0f, w3 = 0.0f;
    bool stepping = false;  // w1..w3 hold the edge functions of triangle last_t at (px, py)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      have[k] = false;
      out[k] = qnan;
      if ((o[k] >> OWNER_T_BITS) == epoch) {
        const int t = (int)(o[k] & ((1u << OWNER_T_BITS) - 1u));
        if (t != last_t) {
          s = m.setup[t];
          last_t = t;
          stepping = false;
        }
        if (!(s.flags & 2)) {
          // the owner word says this triangle covers (px, py): shade without re-testing
          if (!stepping) {
            w1 = edge_at(s.A1, s.B1, s.C1, px, py);
            w2 = edge_at(s.A2, s.B2, s.C2, px, py);
            w3 = edge_at(s.A3, s.B3, s.C3, px, py);
            stepping = true;
          }
          const float norm = __fadd_rn(w1, __fadd_rn(w2, w3));
          const float num = __fadd_rn(__fmul_rn(s.v1, w1), __fadd_rn(__fmul_rn(s.v2, w2), __fmul_rn(s.v3, w3)));
          out[k] = __fdiv_rn(num, norm);
          have[k] = true;
        } else {
          float v;
          // the sweep row itself wins over the wrapped tail of the row above (written later)
          if (tri_eval(s, px, py, &v) || tri_eval(s, px + m.cols, py - 1, &v)) {
            have[k] = true;
            out[k] = v;
          }
        }
      } else {
        stepping = false;
      }
      if (stepping) {  // one pixel to the right: exact for integer-valued edge functions
        w1 = __fadd_rn(w1, s.A1);
        w2 = __fadd_rn(w2, s.A2);
        w3 = __fadd_rn(w3, s.A3);
      }
      if (++px == m.cols) {
        px = 0;
        ++py;
        stepping = false;
      }
    }
    if (full && (fill_nan || (have[0] && have[1] && have[2] && have[3]))) {
      *reinterpret_cast<float4*>(m.img + i) = make_float4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (i + k < n && (have[k] || fill_nan)) m.img[i + k] = out[k];
    }
  }
}

__global__ void k_values_from_graph(const GraphDev* __restrict__ gs, const MeshDev* __restrict__ ms,
                                    float scale) {
  const GraphDev g = gs[blockIdx.y];
  const MeshDev m = ms[blockIdx.y];
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < g.V; v += gridDim.x * blockDim.x)
    m.values[g.vperm[v]] = __fmul_rn(g.vprim[v].x, scale);  // vtx_idepths_ = x * graph_scale_, flame.cc:376
}

int raster_values_from_graph(cudaStream_t st, const GraphDev* gs_dev, const MeshDev* ms_dev, int n,
                             int maxV, float scale) {
  if (maxV == 0 || n == 0) return 0;
  k_values_from_graph<<<dim3(std::min((maxV + 255) / 256, 64), n), 256, 0, st>>>(gs_dev, ms_dev, scale);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Display filters (Flame::obliqueTriangleFilter / edgeLengthFilter / idepthTriangleFilter,
// flame.cc:2207-2356, combined as at flame.cc:389-407) and the coverage statistic (flame.cc:434-443)
// ------------------------------------------------------------------------------------------
struct TriFilterArgs {
  flame_b200_tri_filter_params p;
  float Kinv[9];
};

// p = Kinv * (x, y, 1) / idepth, Eigen's evaluation order (flame.cc:2226-2239)
__device__ __forceinline__ void backproject(const float* Kinv, float x, float y, float id, float out[3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    float r = __fmul_rn(Kinv[3 * i], x);
    r = __fadd_rn(r, __fmul_rn(Kinv[3 * i + 1], y));
    r = __fadd_rn(r, __fmul_rn(Kinv[3 * i + 2], 1.0f));
    out[i] = __fdiv_rn(r, id);
  }
}
__device__ __forceinline__ void normalize3(float v[3]) {  // Eigen 3.2: *this /= norm()
  float n2 = __fmul_rn(v[0], v[0]);
  n2 = __fadd_rn(n2, __fmul_rn(v[1], v[1]));
  n2 = __fadd_rn(n2, __fmul_rn(v[2], v[2]));
  const float n = __fsqrt_rn(n2);
  v[0] = __fdiv_rn(v[0], n); v[1] = __fdiv_rn(v[1], n); v[2] = __fdiv_rn(v[2], n);
}

__global__ void __launch_bounds__(256) k_tri_filter(MeshDev m, TriFilterArgs a, unsigned char* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m.T) return;
  const int ia = m.tris[3 * t], ib = m.tris[3 * t + 1], ic = m.tris[3 * t + 2];
  const float id0 = m.values[ia], id1 = m.values[ib], id2 = m.values[ic];
  const float ax = m.pos[2 * ia], ay = m.pos[2 * ia + 1];
  const float bx = m.pos[2 * ib], by = m.pos[2 * ib + 1];
  const float cx = m.pos[2 * ic], cy = m.pos[2 * ic + 1];
  bool ok = true;
  if (a.p.do_oblique_triangle_filter) {
    float p0[3], p1[3], p2[3];
    backproject(a.Kinv, ax, ay, id0, p0);
    backproject(a.Kinv, bx, by, id1, p1);
    backproject(a.Kinv, cx, cy, id2, p2);
    const float d1[3] = {__fsub_rn(p1[0], p0[0]), __fsub_rn(p1[1], p0[1]), __fsub_rn(p1[2], p0[2])};
    const float d2[3] = {__fsub_rn(p2[0], p0[0]), __fsub_rn(p2[1], p0[1]), __fsub_rn(p2[2], p0[2])};
    float n[3] = {__fsub_rn(__fmul_rn(d1[1], d2[2]), __fmul_rn(d1[2], d2[1])),
                  __fsub_rn(__fmul_rn(d1[2], d2[0]), __fmul_rn(d1[0], d2[2])),
                  __fsub_rn(__fmul_rn(d1[0], d2[1]), __fmul_rn(d1[1], d2[0]))};
    normalize3(n);
    float ray[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) ray[i] = __fdiv_rn(__fadd_rn(__fadd_rn(p0[i], p1[i]), p2[i]), 3.0f);
    normalize3(ray);
    float dot = __fmul_rn(ray[0], n[0]);
    dot = __fadd_rn(dot, __fmul_rn(ray[1], n[1]));
    dot = __fadd_rn(dot, __fmul_rn(ray[2], n[2]));
    const float angle = fabsf(acosf(dot));  // libdevice acosf: within 2 ulp of libm's (see the test)
    if (angle > a.p.oblique_normal_thresh) ok = false;
    float min_id = (id0 < id1) ? id0 : id1;
    min_id = (min_id < id2) ? min_id : id2;
    float max_id = (id0 > id1) ? id0 : id1;
    max_id = (max_id > id2) ? max_id : id2;
    const float diff = __fsub_rn(max_id, min_id);
    if (__fdiv_rn(diff, max_id) > a.p.oblique_idepth_diff_factor) ok = false;
    if (diff > a.p.oblique_idepth_diff_abs) ok = false;
  }
  if (a.p.do_edge_length_filter) {
    float thr = __fmul_rn(a.p.edge_length_thresh, (float)m.cols);
    thr = __fmul_rn(thr, thr);
    const float d01x = __fsub_rn(ax, bx), d01y = __fsub_rn(ay, by);
    const float d02x = __fsub_rn(ax, cx), d02y = __fsub_rn(ay, cy);
    const float d12x = __fsub_rn(bx, cx), d12y = __fsub_rn(by, cy);
    const float dist01 = __fadd_rn(__fmul_rn(d01x, d01x), __fmul_rn(d01y, d01y));
    const float dist02 = __fadd_rn(__fmul_rn(d02x, d02x), __fmul_rn(d02y, d02y));
    const float dist12 = __fadd_rn(__fmul_rn(d12x, d12x), __fmul_rn(d12y, d12y));
    if (dist01 > thr || dist02 > thr || dist12 > thr) ok = false;
  }
  if (a.p.do_idepth_triangle_filter) {
    const float mean = __fdiv_rn(__fadd_rn(__fadd_rn(id0, id1), id2), 3.0f);
    if (mean < a.p.min_triangle_idepth) ok = false;
  }
  out[t] = ok ? 1 : 0;
}

int raster_tri_filter(cudaStream_t st, const MeshDev& m, const flame_b200_tri_filter_params& p,
                      const float Kinv[9], unsigned char* out) {
  if (m.T == 0) return 0;
  TriFilterArgs a;
  a.p = p;
  for (int i = 0; i < 9; ++i) a.Kinv[i] = Kinv[i];
  k_tri_filter<<<(m.T + 255) / 256, 256, 0, st>>>(m, a, out);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Flame::getVertexNormals (flame.cc:2554-2640): per vertex, the running average of the outward normals of
// its triangles IN TRIANGLE ORDER, re-normalised after every update. Vertex -> triangle lists are built on
// the device (count / scan / fill); a vertex thread walks its list in ascending triangle id.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_vt_count(MeshDev m, int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 3 * m.T) atomicAdd(&cnt[m.tris[i]], 1);
}

// one CTA of 1024 threads: exclusive scan of cnt[0..n) into off[0..n], zero cnt (reused as fill cursor)
__global__ void __launch_bounds__(1024) k_vt_scan(int* __restrict__ cnt, int* __restrict__ off, int n) {
  __shared__ int s_warp[32];
  __shared__ int s_run;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int c = (i < n) ? cnt[i] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += v;
      }
      s_warp[lane] = w;
    }
    __syncthreads();
    const int run = s_run;
    if (i < n) {
      off[i] = run + (warp > 0 ? s_warp[warp - 1] : 0) + incl - c;
      cnt[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_run = run + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) off[n] = s_run;
}

__global__ void __launch_bounds__(256) k_vt_fill(MeshDev m, int* __restrict__ cursor, const int* __restrict__ off,
                                                 int* __restrict__ list) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * m.T) return;
  const int v = m.tris[i];
  list[off[v] + atomicAdd(&cursor[v], 1)] = i / 3;
}

__global__ void __launch_bounds__(128) k_vertex_normals(MeshDev m, TriFilterArgs a, int V,
                                                        const int* __restrict__ off, const int* __restrict__ list,
                                                        float* __restrict__ normals) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  float nx = 0.0f, ny = 0.0f, nz = 0.0f;
  int count = 0, last = -1;
  const int b = off[v], e = off[v + 1];
  for (int step = b; step < e; ++step) {
    int t = 0x7fffffff;  // next triangle of this vertex in ascending id (lists are short: selection)
    for (int j = b; j < e; ++j) {
      const int c = list[j];
      if (c > last && c < t) t = c;
    }
    if (t == 0x7fffffff) break;  // a triangle that names the vertex twice appears twice in the list
    last = t;
    const int ia = m.tris[3 * t], ib = m.tris[3 * t + 1], ic = m.tris[3 * t + 2];
    const float id0 = m.values[ia], id1 = m.values[ib], id2 = m.values[ic];
    if ((id0 <= 0) || (id1 <= 0) || (id2 <= 0)) continue;
    float p0[3], p1[3], p2[3];
    backproject(a.Kinv, m.pos[2 * ia], m.pos[2 * ia + 1], id0, p0);
    backproject(a.Kinv, m.pos[2 * ib], m.pos[2 * ib + 1], id1, p1);
    backproject(a.Kinv, m.pos[2 * ic], m.pos[2 * ic + 1], id2, p2);
    const float d1[3] = {__fsub_rn(p1[0], p0[0]), __fsub_rn(p1[1], p0[1]), __fsub_rn(p1[2], p0[2])};
    const float d2[3] = {__fsub_rn(p2[0], p0[0]), __fsub_rn(p2[1], p0[1]), __fsub_rn(p2[2], p0[2])};
    float n[3] = {__fsub_rn(__fmul_rn(d2[1], d1[2]), __fmul_rn(d2[2], d1[1])),
                  __fsub_rn(__fmul_rn(d2[2], d1[0]), __fmul_rn(d2[0], d1[2])),
                  __fsub_rn(__fmul_rn(d2[0], d1[1]), __fmul_rn(d2[1], d1[0]))};  // delta2 x delta1: outward
    normalize3(n);
    // a triangle that lists this vertex k times updates it k times (flame.cc:2607-2625)
    const int reps = (ia == v) + (ib == v) + (ic == v);
    for (int r = 0; r < reps; ++r) {
      const float c = (float)count, c1 = (float)(count + 1);
      float nv[3] = {__fdiv_rn(__fadd_rn(__fmul_rn(c, nx), n[0]), c1), __fdiv_rn(__fadd_rn(__fmul_rn(c, ny), n[1]), c1),
                     __fdiv_rn(__fadd_rn(__fmul_rn(c, nz), n[2]), c1)};
      normalize3(nv);
      nx = nv[0]; ny = nv[1]; nz = nv[2];
      ++count;
    }
  }
  normals[3 * (size_t)v] = nx;
  normals[3 * (size_t)v + 1] = ny;
  normals[3 * (size_t)v + 2] = nz;
}

// scratch: 2 (V + 1) + 3 T ints [cnt | off | list]
int raster_vertex_normals(cudaStream_t st, const MeshDev& m, int V, const float Kinv[9], int* scratch,
                          float* d_normals) {
  if (V == 0) return 0;
  int* cnt = scratch;
  int* off = scratch + (V + 1);
  int* list = off + (V + 1);
  FB_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(V + 1) * sizeof(int), st));
  if (m.T > 0) {
    k_vt_count<<<(3 * m.T + 255) / 256, 256, 0, st>>>(m, cnt);
    count_launch();
  }
  k_vt_scan<<<1, 1024, 0, st>>>(cnt, off, V);
  count_launch();
  if (m.T > 0) {
    k_vt_fill<<<(3 * m.T + 255) / 256, 256, 0, st>>>(m, cnt, off, list);
    count_launch();
  }
  TriFilterArgs a;
  flame_b200_tri_filter_default_params(&a.p);
  for (int i = 0; i < 9; ++i) a.Kinv[i] = Kinv[i];
  k_vertex_normals<<<(V + 127) / 128, 128, 0, st>>>(m, a, V, off, list, d_normals);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

__global__ void __launch_bounds__(256) k_coverage(MeshDev m, unsigned long long* __restrict__ count) {
  const int n = m.rows * m.cols;
  int mine = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float v = m.img[i];
    mine += !(v != v);
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(count, (unsigned long long)mine);
}

int raster_coverage(cudaStream_t st, const MeshDev& m, unsigned long long* d_count) {
  FB_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
  const int n = m.rows * m.cols;
  if (n > 0) {
    k_coverage<<<std::min((n + 255) / 256, 148 * 8), 256, 0, st>>>(m, d_count);
    count_launch();
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ms_dev: device array of n MeshDev. Epoch and big-list counter re-arm themselves on the device, so a
// projection needs no memset and no host round trip.
int raster_project(cudaStream_t st, const MeshDev* ms_dev, int n, int maxT, int max_pix, int fill_nan) {
  if (n == 0 || max_pix == 0) return 0;
  k_tri_setup<<<dim3(std::max(1, (maxT + 255) / 256), n), 256, 0, st>>>(ms_dev);
  count_launch();
  if (maxT > 0) {
    k_owner_small<<<dim3((maxT + 7) / 8, n), 256, 0, st>>>(ms_dev);
    k_owner_big<<<dim3(32, n), 256, 0, st>>>(ms_dev);
    count_launch(2);
  }
  k_resolve<<<dim3((max_pix / 4 + 255) / 256 + 1, n), 256, 0, st>>>(ms_dev, fill_nan);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb
