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
// Edge functions are integer-valued: the reference's fp32 evaluation is exact while
// |A*x + B*y + C| < 2^24, i.e. for images up to 2896 px on a side (enforced at mesh_bind), so they are
// evaluated here in int32 and converted (exactly) to fp32 for the shading.
//
// Formulation. The reference scatters triangles in list order with read-modify-write; the result at a
// pixel is the shading of the HIGHEST-index triangle that covers it. Here every pixel is written once:
//   k_tri_static  one thread per triangle, only when the mesh or its vertex positions changed: integer edge
//                 coefficients, sweep footprint, and insertion of the triangle into the fixed-capacity bins
//                 of the 16x8-pixel screen tiles its footprint overlaps. A triangle that overlaps more than
//                 256 tiles, wraps past the right border, or finds a bin full goes to the mesh's "wide"
//                 list instead (hull slivers; every tile tests that short list against its rectangle).
//   k_tiles       one WARP per tile, 4 pixels per lane: the tile's triangles are staged through shared
//                 memory (ballot-compacted, no block barrier), every pixel keeps the highest covering
//                 triangle id in a register, shades once from that triangle's record and is stored with
//                 128-bit writes. No atomics, no per-pixel owner buffer in HBM: the DRAM traffic of a
//                 projection is the image plus the (cached, L2-resident) bins.
// The bins depend on topology and positions only, so a slot that is projected repeatedly (new x, same
// graph) pays for k_tiles alone.
#include <algorithm>
#include <cmath>

#include "internal.h"

namespace fb {

#define TILE_W 16
#define TILE_H 8
#define TILE_WARPS 4       // tiles per CTA
#define MAX_BIN_TILES 256  // a triangle overlapping more tiles goes to the wide list
#define COORD_LIM 8191     // |rounded vertex coordinate| clamp: keeps the int32 edge functions overflow-free

__device__ __forceinline__ int min3(int a, int b, int c) { return min(a, min(b, c)); }
__device__ __forceinline__ int max3(int a, int b, int c) { return max(a, max(b, c)); }

__device__ __forceinline__ int tiles_x(const MeshDev& m) { return (m.cols + TILE_W - 1) / TILE_W; }
__device__ __forceinline__ int tiles_y(const MeshDev& m) { return (m.rows + TILE_H - 1) / TILE_H; }

__device__ __forceinline__ int round_coord(float v) {
  return max(-COORD_LIM, min(COORD_LIM, __float2int_rn(v)));  // cv::Point2f -> cv::Point (half to even)
}

// Footprint of the pixels triangle s may WRITE, clipped to the image: columns [fx0, fx1], rows [fy0, fy1];
// empty when fx0 > fx1 or fy0 > fy1. (A wrapping triangle additionally writes the head of the next rows.)
__device__ __forceinline__ void footprint(const TriStatic& s, const MeshDev& m, int* fx0, int* fx1, int* fy0,
                                          int* fy1) {
  *fx0 = max(s.xmin, 0);
  *fy0 = max(s.ymin, 0);
  *fx1 = min(s.fx1, m.cols - 1);
  *fy1 = min(s.ymax, m.rows - 1);
}

// Eight lanes per triangle: the record is computed redundantly, the footprint's tiles are spread over the lanes.
__global__ void __launch_bounds__(256) k_tri_static(const MeshDev* __restrict__ ms) {
  const MeshDev m = ms[blockIdx.y];
  const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 3, sub = threadIdx.x & 7;
  const unsigned group = 0xffu << (threadIdx.x & 24);  // the eight lanes of this triangle
  if (t >= m.T) return;  // whole groups leave together (T is a per-group bound)
  TriStatic s;
  const int ia = m.tris[3 * t], ib = m.tris[3 * t + 1], ic = m.tris[3 * t + 2];
  // image_utils.cc:385-391: p1 = vertices[tri[2]], p2 = vertices[tri[1]], p3 = vertices[tri[0]]
  const int p1x = round_coord(m.pos[2 * ic]), p1y = round_coord(m.pos[2 * ic + 1]);
  const int p2x = round_coord(m.pos[2 * ib]), p2y = round_coord(m.pos[2 * ib + 1]);
  const int p3x = round_coord(m.pos[2 * ia]), p3y = round_coord(m.pos[2 * ia + 1]);
  s.xmin = min3(p1x, p2x, p3x); s.ymin = min3(p1y, p2y, p3y);
  const int xmax = max3(p1x, p2x, p3x);
  s.ymax = max3(p1y, p2y, p3y);
  s.xend = s.xmin + ((xmax - s.xmin) / 4 + 1) * 4;  // columns visited by the 4-wide sweep
  // w1 = e23, w2 = e31, w3 = e12 (rasterization.cc:177-179; Edge::init rasterization.h:129-153)
  s.A1 = p3y - p2y; s.B1 = p2x - p3x; s.C1 = p3x * p2y - p2x * p3y;
  s.A2 = p1y - p3y; s.B2 = p3x - p1x; s.C2 = p1x * p3y - p3x * p1y;
  s.A3 = p2y - p1y; s.B3 = p1x - p2x; s.C3 = p2x * p1y - p1x * p2y;
  const long long area2 = (long long)(p2x - p1x) * (p3y - p1y) - (long long)(p2y - p1y) * (p3x - p1x);
  const bool degenerate = (area2 == 0);
  // a non-degenerate triangle covers nothing outside [xmin, xmax]; a degenerate one (edge functions vanish
  // on a line) also covers the sweep's overshoot columns up to xend - 1
  s.fx1 = degenerate ? s.xend - 1 : xmax;
  const bool wraps = degenerate && (s.xend > m.cols);  // overshoot columns land on the head of the next row
  s.flags = (degenerate ? 4 : 0) | (wraps ? 2 : 0);
  s.i1 = ic; s.i2 = ib; s.i3 = ia;
  s.pad0 = s.pad1 = 0;
  int fx0, fx1, fy0, fy1;
  footprint(s, m, &fx0, &fx1, &fy0, &fy1);
  // triangles the caller marked invalid (tri_valid / vtx_valid, fixed at mesh_bind like the bins) are not binned
  if ((m.tri_ok && !m.tri_ok[t]) || fy0 > fy1 || (!wraps && fx0 > fx1)) {
    s.flags |= 1;  // writes nothing
  } else {
    const int tx0 = fx0 / TILE_W, tx1 = fx1 / TILE_W, ty0 = fy0 / TILE_H, ty1 = fy1 / TILE_H;
    const int nx = tx1 - tx0 + 1, n = nx * (ty1 - ty0 + 1);
    bool wide = wraps || n > MAX_BIN_TILES;
    if (!wide) {
      const int ntx = tiles_x(m);
      for (int i = sub; i < n; i += 8) {
        const int tyi = ty0 + i / nx, txi = tx0 + i % nx;
        if (!degenerate) {
          // exact reject: an edge function that is negative at its most favourable corner of the tile is
          // negative on the whole tile (integers, no rounding) - about a third of the bounding-box tiles
          const int xl = txi * TILE_W, xh = xl + TILE_W - 1, yl = tyi * TILE_H, yh = yl + TILE_H - 1;
          const int m1 = s.A1 * (s.A1 >= 0 ? xh : xl) + s.B1 * (s.B1 >= 0 ? yh : yl) + s.C1;
          const int m2 = s.A2 * (s.A2 >= 0 ? xh : xl) + s.B2 * (s.B2 >= 0 ? yh : yl) + s.C2;
          const int m3 = s.A3 * (s.A3 >= 0 ? xh : xl) + s.B3 * (s.B3 >= 0 ? yh : yl) + s.C3;
          if ((m1 | m2 | m3) < 0) continue;
        }
        const int tile = tyi * ntx + txi;
        const int at = atomicAdd(&m.tile_cnt[tile], 1);
        if (at < m.bin_cap) m.bins[(size_t)tile * m.bin_cap + at] = t;
        else wide = true;  // bin full: the wide list keeps the triangle reachable from every tile
      }
      wide = __any_sync(group, wide);
    }
    if (wide) {
      s.flags |= 8;
      if (sub == 0) m.wide[atomicAdd(&m.counters[0], 1)] = t;
    }
  }
  if (sub == 0) m.stat[t] = s;
}

struct __align__(16) StagedTri {  // shared-memory copy of one triangle of the current chunk (five 16-byte words)
  int A1, B1, C1, key;  // key = (triangle id << 5) | position in the chunk: the per-pixel maximum names the
  int A2, B2, C2, flags;  //     winner AND where its copy is staged
  int A3, B3, C3, xmin;
  int xend, ymin, ymax, pad;
  float v1, v2, v3, pad2;  // values at p1, p2, p3
};

// coverage of sweep position (x, y) by a DEGENERATE triangle: the edge test alone is true on the whole
// line, the sweep only visits [xmin, xend) x [ymin, ymax] (x may exceed cols - 1: 4-px overshoot)
template <typename Tri>
__device__ __forceinline__ bool degenerate_covers(const Tri& s, int x, int y, int* w1, int* w2, int* w3) {
  if (y < s.ymin || y > s.ymax || x < s.xmin || x >= s.xend) return false;
  *w1 = s.A1 * x + s.B1 * y + s.C1;
  *w2 = s.A2 * x + s.B2 * y + s.C2;
  *w3 = s.A3 * x + s.B3 * y + s.C3;
  return (*w1 | *w2 | *w3) >= 0;
}

__global__ void __launch_bounds__(32 * TILE_WARPS, 10) k_tiles(const MeshDev* __restrict__ ms, int fill_nan) {
  const MeshDev m = ms[blockIdx.z];
  const int ntx = tiles_x(m), nty = tiles_y(m);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = blockIdx.x * TILE_WARPS + warp, ty = blockIdx.y;  // a CTA = TILE_WARPS neighbouring tiles of a row
  if (tx >= ntx || ty >= nty) return;  // warps are independent: no block-wide barrier below
  const int tile = ty * ntx + tx;
  const int px = tx * TILE_W + (lane & 3) * 4, py = ty * TILE_H + (lane >> 2);
  __shared__ StagedTri s_all[TILE_WARPS][32];
  StagedTri* s_tri = s_all[warp];
  int best[4] = {-1, -1, -1, -1};  // per pixel: key of the highest covering triangle so far
  const int nbin = m.T > 0 ? min(m.tile_cnt[tile], m.bin_cap) : 0;
  const int nwide = m.T > 0 ? m.counters[0] : 0;
  const int* bin = m.bins + (size_t)tile * m.bin_cap;
  const int x_lo = tx * TILE_W, x_hi = x_lo + TILE_W - 1, y_lo = ty * TILE_H, y_hi = y_lo + TILE_H - 1;
  for (int e0 = 0; e0 < nbin + nwide; e0 += 32) {
    const int e = e0 + lane;
    bool use = e < nbin + nwide;
    int t = 0;
    TriStatic s;
    if (use) {
      t = e < nbin ? bin[e] : m.wide[e - nbin];
      s = m.stat[t];
      use = !(m.tri_filt && !m.tri_filt[t]);  // (invalid triangles were never binned)
    }
    if (use) {
      if (e >= nbin) {  // wide list: does its footprint touch this tile at all?
        int fx0, fx1, fy0, fy1;
        footprint(s, m, &fx0, &fx1, &fy0, &fy1);
        use = fx0 <= x_hi && fx1 >= x_lo && fy0 <= y_hi && fy1 >= y_lo;
        if ((s.flags & 2) && !use)  // wrapped tail: columns [0, xend - 1 - cols] of rows fy0 + 1 .. fy1 + 1
          use = x_lo <= s.xend - 1 - m.cols && fy0 + 1 <= y_hi && fy1 + 1 >= y_lo;
      }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, use);
    __syncwarp();  // the previous chunk has been consumed
    if (use) {
      StagedTri& d = s_tri[__popc(mask & ((1u << lane) - 1u))];
      d.A1 = s.A1; d.B1 = s.B1; d.C1 = s.C1; d.A2 = s.A2; d.B2 = s.B2; d.C2 = s.C2;
      d.A3 = s.A3; d.B3 = s.B3; d.C3 = s.C3;
      d.xmin = s.xmin; d.xend = s.xend; d.ymin = s.ymin; d.ymax = s.ymax; d.flags = s.flags;
      d.key = (t << 5) | __popc(mask & ((1u << lane) - 1u));
      // values at p1, p2, p3 = reversed triangle order (image_utils.cc:385-391)
      d.v1 = m.values[s.i1];
      d.v2 = m.values[s.i2];
      d.v3 = m.values[s.i3];
    }
    __syncwarp();
    const int n = __popc(mask);
    for (int j = 0; j < n; ++j) {
      const StagedTri& q = s_tri[j];
      if (!(q.flags & 4)) {
        // edge functions at the lane's first pixel; one step to the right adds A (integers: exact)
        int w1 = q.A1 * px + q.B1 * py + q.C1;
        int w2 = q.A2 * px + q.B2 * py + q.C2;
        int w3 = q.A3 * px + q.B3 * py + q.C3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          best[k] = max(best[k], (w1 | w2 | w3) >= 0 ? q.key : -1);
          w1 += q.A1; w2 += q.A2; w3 += q.A3;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int w1, w2, w3;
          bool hit = degenerate_covers(q, px + k, py, &w1, &w2, &w3);
          if (!hit && (q.flags & 2)) hit = degenerate_covers(q, px + k + m.cols, py - 1, &w1, &w2, &w3);
          if (hit) best[k] = max(best[k], q.key);
        }
      }
    }
  }
  if (py >= m.rows || px >= m.cols) return;
  // shade every pixel once with its winner. One chunk (the rule): the winner's coefficients and values are
  // still staged in shared memory, the low key bits say where; several chunks: fetched again by id.
  const bool staged = nbin + nwide <= 32;
  const float qnan = __uint_as_float(0x7fc00000u);  // quiet_NaN, flame.cc:412
  float out[4];
  bool all = true;
  int last = -1;
  StagedTri s;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    out[k] = qnan;
    if (best[k] < 0) {
      all = false;
      continue;
    }
    if (best[k] != last) {
      last = best[k];
      if (staged) {
        s = s_tri[last & 31];
      } else {
        const int t = last >> 5;
        const TriStatic g = m.stat[t];
        s.A1 = g.A1; s.B1 = g.B1; s.C1 = g.C1; s.A2 = g.A2; s.B2 = g.B2; s.C2 = g.C2;
        s.A3 = g.A3; s.B3 = g.B3; s.C3 = g.C3;
        s.xmin = g.xmin; s.xend = g.xend; s.ymin = g.ymin; s.ymax = g.ymax; s.flags = g.flags;
        s.v1 = m.values[g.i1];
        s.v2 = m.values[g.i2];
        s.v3 = m.values[g.i3];
      }
    }
    int w1, w2, w3;
    if (!(s.flags & 4)) {
      w1 = s.A1 * (px + k) + s.B1 * py + s.C1;
      w2 = s.A2 * (px + k) + s.B2 * py + s.C2;
      w3 = s.A3 * (px + k) + s.B3 * py + s.C3;
    } else if (!degenerate_covers(s, px + k, py, &w1, &w2, &w3)) {
      // the sweep row itself wins over the wrapped tail of the row above (written earlier by the same sweep)
      degenerate_covers(s, px + k + m.cols, py - 1, &w1, &w2, &w3);
    }
    const float f1 = (float)w1, f2 = (float)w2, f3 = (float)w3;  // exact: |w| < 2^24
    const float norm = __fadd_rn(f1, __fadd_rn(f2, f3));
    const float num = __fadd_rn(__fmul_rn(s.v1, f1), __fadd_rn(__fmul_rn(s.v2, f2), __fmul_rn(s.v3, f3)));
    out[k] = __fdiv_rn(num, norm);
  }
  float* dst = m.img + (size_t)py * m.cols + px;
  if (px + 4 <= m.cols && (m.cols & 3) == 0 && (fill_nan || all)) {
    *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[2], out[3]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (px + k < m.cols && (best[k] >= 0 || fill_nan)) dst[k] = out[k];
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

__global__ void __launch_bounds__(256) k_tri_filter(MeshDev m, TriFilterArgs a, unsigned char* __restrict__ out,
                                                    float* __restrict__ angle_out) {
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
    if (angle_out) angle_out[t] = angle;
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
                      const float Kinv[9], unsigned char* out, float* angle_out) {
  if (m.T == 0) return 0;
  TriFilterArgs a;
  a.p = p;
  for (int i = 0; i < 9; ++i) a.Kinv[i] = Kinv[i];
  k_tri_filter<<<(m.T + 255) / 256, 256, 0, st>>>(m, a, out, angle_out);
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

// ms_dev: device array of n MeshDev. rebuild != 0: topology or positions of (some of) the meshes changed
// since their bins were built; the caller has zeroed [counters | tile_cnt] of every mesh of the batch.
int raster_project(cudaStream_t st, const MeshDev* ms_dev, int n, int maxT, int max_rows, int max_cols, int fill_nan,
                   int rebuild) {
  if (n == 0 || max_rows <= 0 || max_cols <= 0) return 0;
  const int gx = ((max_cols + TILE_W - 1) / TILE_W + TILE_WARPS - 1) / TILE_WARPS, gy = (max_rows + TILE_H - 1) / TILE_H;
  if (rebuild && maxT > 0) {
    k_tri_static<<<dim3((maxT + 31) / 32, n), 256, 0, st>>>(ms_dev);
    count_launch();
  }
  k_tiles<<<dim3(gx, gy, n), 32 * TILE_WARPS, 0, st>>>(ms_dev, fill_nan);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

int raster_tiles(int rows, int cols) { return ((cols + TILE_W - 1) / TILE_W) * ((rows + TILE_H - 1) / TILE_H); }

// Bin capacity for a mesh of T triangles on a rows x cols image: twice the expected number of triangles
// whose bounding box touches a 16x8 tile (triangles of equal area tiling the image), at least 16.
int raster_bin_cap(int T, int rows, int cols) {
  if (T <= 0) return 16;
  const double a = (double)rows * cols / T, side = std::sqrt(2.0 * a);
  const double per_tile = (side + TILE_W) * (side + TILE_H) / a;
  const int cap = (int)(2.0 * per_tile) + 8;
  return std::min(std::max((cap + 7) / 8 * 8, 16), 1024);
}

}  // namespace fb
