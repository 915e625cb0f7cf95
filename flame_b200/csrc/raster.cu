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
// Formulation: the reference scatters triangles in order with read-modify-write; here every PIXEL
// gathers "the highest-index triangle that covers me" from a per-tile bin and is written exactly
// once (4*W*H bytes of image traffic instead of fill + per-triangle RMW):
//   k_tri_setup   one thread per triangle: rounding, edge coefficients, footprint -> tile counts
//   k_tile_scan   one CTA: exclusive scan of the tile counts
//   k_tri_fill    one thread per triangle: (tile, triangle) pairs
//   k_tile_raster one CTA per 32x8 pixel block = four 8x8 tiles, two warps per tile: pixels walk
//                 their tile's bin (+ the short "large" list); bins hold only triangles whose
//                 edge functions do not exclude the whole tile (exact integer box test)
// Triangles with a big bounding box (hull slivers: few tiles actually overlapped) are binned by a
// whole warp (k_big_count / k_big_fill) against a global pair budget; only what exceeds the budget,
// or wraps past the right image border (degenerate), goes to the "large" list that every tile
// scans. Pair buffer: SMALL_TILES * T + BIG_BUDGET_PER_TILE * n_tiles entries.
#include <algorithm>

#include "internal.h"

namespace fb {

#define TILE_W 8
#define TILE_H 8
#define CTA_TILES 4   // a CTA rasterises CTA_TILES horizontally adjacent tiles (32 x 8 pixels)
#define SMALL_TILES FB_SMALL_TILES_PER_TRI
#define BIG_BUDGET_PER_TILE FB_BIG_BUDGET_PER_TILE

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

// Conservative-exact test "does the closed triangle (all three w >= 0) miss the pixel box
// [x0,x1] x [y0,y1] entirely?": an edge function that is negative at its most favourable corner is
// negative on the whole box. Integer-valued floats: exact.
__device__ __forceinline__ bool tile_outside(const TriSetup& s, int x0, int y0, int x1, int y1) {
  const float m1 = edge_at(s.A1, s.B1, s.C1, s.A1 >= 0.0f ? x1 : x0, s.B1 >= 0.0f ? y1 : y0);
  const float m2 = edge_at(s.A2, s.B2, s.C2, s.A2 >= 0.0f ? x1 : x0, s.B2 >= 0.0f ? y1 : y0);
  const float m3 = edge_at(s.A3, s.B3, s.C3, s.A3 >= 0.0f ? x1 : x0, s.B3 >= 0.0f ? y1 : y0);
  return (m1 < 0.0f) || (m2 < 0.0f) || (m3 < 0.0f);
}

__global__ void __launch_bounds__(256) k_tri_setup(const MeshDev* __restrict__ ms) {
  const MeshDev m = ms[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m.T) return;
  TriSetup s;
  s.flags = 0;
  s.fx1 = -1;
  s.pad0 = s.pad1 = 0;
  if (m.tri_ok && !m.tri_ok[t]) {
    s.ymin = 1; s.ymax = 0; s.xmin = 0; s.xend = 0;  // empty footprint
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
  // footprint in pixels: a non-degenerate triangle covers nothing outside [xmin, xmax]
  const int fx0 = max(s.xmin, 0), fy0 = max(s.ymin, 0);
  const int fx1 = min(degenerate ? s.xend - 1 : xmax, m.cols - 1), fy1 = min(s.ymax, m.rows - 1);
  const bool wraps = degenerate && (s.xend > m.cols);
  s.fx1 = fx1;
  const int tx0 = fx0 / TILE_W, tx1 = fx1 / TILE_W, ty0 = fy0 / TILE_H, ty1 = fy1 / TILE_H;
  const long long ntiles = (fx1 >= fx0 && fy1 >= fy0) ? (long long)(tx1 - tx0 + 1) * (ty1 - ty0 + 1) : 0;
  if (degenerate) s.flags |= 4;
  if (wraps) {
    s.flags |= 1 | 2;
    m.large[atomicAdd(&m.counters[0], 1)] = t;
  } else if (ntiles > SMALL_TILES) {
    s.flags |= 8;  // big bounding box: binned by a warp in k_big_count / k_big_fill
    m.big[atomicAdd(&m.counters[2], 1)] = t;
  } else {
    const int tiles_x = (m.cols + TILE_W - 1) / TILE_W;
    for (int ty = ty0; ty <= ty1 && ntiles > 0; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) {
        if (!degenerate && tile_outside(s, tx * TILE_W, ty * TILE_H, tx * TILE_W + TILE_W - 1,
                                        ty * TILE_H + TILE_H - 1))
          continue;
        atomicAdd(&m.tile_count[ty * tiles_x + tx], 1);
      }
  }
  m.setup[t] = s;
}

// One warp per big-bounding-box triangle: exact count of overlapped tiles, then either bin it
// (flags |= 16) within the global pair budget or hand it to the large list (flags |= 1).
__global__ void __launch_bounds__(256) k_big_count(const MeshDev* __restrict__ ms) {
  const MeshDev m = ms[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n_big = m.counters[2];
  const int tiles_x = (m.cols + TILE_W - 1) / TILE_W;
  const int budget = BIG_BUDGET_PER_TILE * m.n_tiles;
  for (int b = warp; b < n_big; b += nwarps) {
    const int t = m.big[b];
    TriSetup s = m.setup[t];
    const int fx0 = max(s.xmin, 0), fy0 = max(s.ymin, 0), fx1 = s.fx1, fy1 = min(s.ymax, m.rows - 1);
    const int tx0 = fx0 / TILE_W, ty0 = fy0 / TILE_H;
    const int nx = fx1 / TILE_W - tx0 + 1, ny = fy1 / TILE_H - ty0 + 1;
    const bool exact = !(s.flags & 4);
    int cnt = 0;
    for (int i = lane; i < nx * ny; i += 32) {
      const int tx = tx0 + i % nx, ty = ty0 + i / nx;
      cnt += !(exact && tile_outside(s, tx * TILE_W, ty * TILE_H, tx * TILE_W + TILE_W - 1,
                                     ty * TILE_H + TILE_H - 1));
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    int ok = 0;
    if (lane == 0) {
      const int old = atomicAdd(&m.counters[4], cnt);
      ok = (old + cnt <= budget);
      if (!ok) {
        atomicSub(&m.counters[4], cnt);
        m.large[atomicAdd(&m.counters[0], 1)] = t;
      }
      m.setup[t].flags = s.flags | (ok ? 16 : 1);
    }
    ok = __shfl_sync(0xffffffffu, ok, 0);
    if (ok)
      for (int i = lane; i < nx * ny; i += 32) {
        const int tx = tx0 + i % nx, ty = ty0 + i / nx;
        if (!(exact && tile_outside(s, tx * TILE_W, ty * TILE_H, tx * TILE_W + TILE_W - 1,
                                    ty * TILE_H + TILE_H - 1)))
          atomicAdd(&m.tile_count[ty * tiles_x + tx], 1);
      }
  }
}

__global__ void __launch_bounds__(256) k_big_fill(const MeshDev* __restrict__ ms) {
  const MeshDev m = ms[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n_big = m.counters[3];
  const int tiles_x = (m.cols + TILE_W - 1) / TILE_W;
  for (int b = warp; b < n_big; b += nwarps) {
    const int t = m.big[b];
    const TriSetup s = m.setup[t];
    if (!(s.flags & 16)) continue;
    const int fx0 = max(s.xmin, 0), fy0 = max(s.ymin, 0), fx1 = s.fx1, fy1 = min(s.ymax, m.rows - 1);
    const int tx0 = fx0 / TILE_W, ty0 = fy0 / TILE_H;
    const int nx = fx1 / TILE_W - tx0 + 1, ny = fy1 / TILE_H - ty0 + 1;
    const bool exact = !(s.flags & 4);
    for (int i = lane; i < nx * ny; i += 32) {
      const int tx = tx0 + i % nx, ty = ty0 + i / nx;
      if (exact && tile_outside(s, tx * TILE_W, ty * TILE_H, tx * TILE_W + TILE_W - 1, ty * TILE_H + TILE_H - 1))
        continue;
      const int tile = ty * tiles_x + tx;
      m.pairs[m.tile_off[tile] + atomicAdd(&m.tile_count[tile], 1)] = t;
    }
  }
}

// one CTA of 1024 threads per slot: exclusive scan of tile_count -> tile_off (n_tiles + 1) in
// coalesced chunks of 1024 tiles (warp-shuffle scan + one cross-warp step), zero the cursors
__global__ void __launch_bounds__(1024) k_tile_scan(const MeshDev* __restrict__ ms) {
  const MeshDev m = ms[blockIdx.y];
  __shared__ int s_warp[32];
  __shared__ int s_run;
  const int n = m.n_tiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_run = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int cnt = (i < n) ? m.tile_count[i] : 0;
    int incl = cnt;
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
      s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    const int run = s_run;
    const int excl = run + (warp > 0 ? s_warp[warp - 1] : 0) + incl - cnt;
    if (i < n) {
      m.tile_off[i] = excl;
      m.tile_count[i] = 0;  // reused as the fill cursor
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_run = run + s_warp[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    m.tile_off[n] = s_run;
    // latch this projection's list lengths, re-arm the accumulators
    m.counters[1] = m.counters[0];
    m.counters[0] = 0;
    m.counters[3] = m.counters[2];
    m.counters[2] = 0;
    m.counters[4] = 0;
  }
}

__global__ void __launch_bounds__(256) k_tri_fill(const MeshDev* __restrict__ ms) {
  const MeshDev m = ms[blockIdx.y];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m.T) return;
  const TriSetup s = m.setup[t];
  if (s.flags & (1 | 8)) return;  // large list / binned by k_big_fill
  if (s.ymin > s.ymax) return;
  // exactly the footprint counted in k_tri_setup
  const int fx0 = max(s.xmin, 0), fy0 = max(s.ymin, 0);
  const int fx1 = s.fx1, fy1 = min(s.ymax, m.rows - 1);
  if (fx1 < fx0 || fy1 < fy0) return;
  const int tiles_x = (m.cols + TILE_W - 1) / TILE_W;
  for (int ty = fy0 / TILE_H; ty <= fy1 / TILE_H; ++ty)
    for (int tx = fx0 / TILE_W; tx <= fx1 / TILE_W; ++tx) {
      if (!(s.flags & 4) && tile_outside(s, tx * TILE_W, ty * TILE_H, tx * TILE_W + TILE_W - 1,
                                         ty * TILE_H + TILE_H - 1))
        continue;
      const int tile = ty * tiles_x + tx;
      m.pairs[m.tile_off[tile] + atomicAdd(&m.tile_count[tile], 1)] = t;
    }
}

#define RCHUNK 16
// CTA = 4 warps = 4 horizontally adjacent 8x8 tiles; a warp owns one tile, a lane two pixels
// (x, y) and (x, y + 4) that share the A*x terms of the edge functions.
__global__ void __launch_bounds__(128) k_tile_raster(const MeshDev* __restrict__ ms, int fill_nan) {
  const MeshDev m = ms[blockIdx.y];
  const int tiles_x = (m.cols + TILE_W - 1) / TILE_W;
  const int ctas_x = (tiles_x + CTA_TILES - 1) / CTA_TILES;
  const int cy = blockIdx.x / ctas_x, cx = blockIdx.x % ctas_x;
  if (cy * TILE_H >= m.rows) return;
  const int sub = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tx = cx * CTA_TILES + sub, ty = cy;
  if (tx >= tiles_x) return;  // whole warp: no CTA-wide barriers below
  const int tile = ty * tiles_x + tx;
  const int px = tx * TILE_W + (lane & 7);
  const int py0 = ty * TILE_H + (lane >> 3), py1 = py0 + 4;
  __shared__ TriSetup s_tri[CTA_TILES][RCHUNK];
  __shared__ int s_id[CTA_TILES][RCHUNK];
  const int beg = m.tile_off[tile], n_bin = m.tile_off[tile + 1] - beg;
  const int n_large = m.counters[1];
  const int n = n_bin + n_large;
  int best0 = -1, best1 = -1;
  float val0 = 0.0f, val1 = 0.0f;
  const float fx = (float)px, fy0 = (float)py0, fy1 = (float)py1;
  for (int c0 = 0; c0 < n; c0 += RCHUNK) {
    const int nc = min(RCHUNK, n - c0);
    __syncwarp();
    if (lane < nc) {
      const int e = c0 + lane;
      const int t = (e < n_bin) ? m.pairs[beg + e] : m.large[e - n_bin];
      s_id[sub][lane] = t;
      s_tri[sub][lane] = m.setup[t];
    }
    __syncwarp();
    for (int j = 0; j < nc; ++j) {
      const int t = s_id[sub][j];
      const TriSetup& s = s_tri[sub][j];
      const bool in_x = (px >= s.xmin) && (px < s.xend);
      const bool c0ok = in_x && (t > best0) && (py0 >= s.ymin) && (py0 <= s.ymax);
      const bool c1ok = in_x && (t > best1) && (py1 >= s.ymin) && (py1 <= s.ymax);
      if (c0ok || c1ok) {
        // (A*x + B*y) + C, each operation rounded (exact: integer-valued)
        const float ax1 = __fmul_rn(s.A1, fx), ax2 = __fmul_rn(s.A2, fx), ax3 = __fmul_rn(s.A3, fx);
        if (c0ok) {
          const float w1 = __fadd_rn(__fadd_rn(ax1, __fmul_rn(s.B1, fy0)), s.C1);
          const float w2 = __fadd_rn(__fadd_rn(ax2, __fmul_rn(s.B2, fy0)), s.C2);
          const float w3 = __fadd_rn(__fadd_rn(ax3, __fmul_rn(s.B3, fy0)), s.C3);
          if (w1 >= 0.0f && w2 >= 0.0f && w3 >= 0.0f) {
            const float num = __fadd_rn(__fmul_rn(s.v1, w1), __fadd_rn(__fmul_rn(s.v2, w2), __fmul_rn(s.v3, w3)));
            val0 = __fdiv_rn(num, __fadd_rn(w1, __fadd_rn(w2, w3)));
            best0 = t;
          }
        }
        if (c1ok) {
          const float w1 = __fadd_rn(__fadd_rn(ax1, __fmul_rn(s.B1, fy1)), s.C1);
          const float w2 = __fadd_rn(__fadd_rn(ax2, __fmul_rn(s.B2, fy1)), s.C2);
          const float w3 = __fadd_rn(__fadd_rn(ax3, __fmul_rn(s.B3, fy1)), s.C3);
          if (w1 >= 0.0f && w2 >= 0.0f && w3 >= 0.0f) {
            const float num = __fadd_rn(__fmul_rn(s.v1, w1), __fadd_rn(__fmul_rn(s.v2, w2), __fmul_rn(s.v3, w3)));
            val1 = __fdiv_rn(num, __fadd_rn(w1, __fadd_rn(w2, w3)));
            best1 = t;
          }
        }
      }
      if (s.flags & 2) {
        // degenerate triangle whose 4-px sweep ran past the right border: the reference's linear
        // addressing lands on (x - cols, y + 1)
        float v;
        if (t > best0 && tri_eval(s, px + m.cols, py0 - 1, &v)) { best0 = t; val0 = v; }
        if (t > best1 && tri_eval(s, px + m.cols, py1 - 1, &v)) { best1 = t; val1 = v; }
      }
    }
  }
  if (px < m.cols) {
    const unsigned int qnan = 0x7fc00000u;  // quiet_NaN, flame.cc:412
    if (py0 < m.rows) {
      const size_t i = (size_t)py0 * m.cols + px;
      if (best0 >= 0) m.img[i] = val0;
      else if (fill_nan) m.img[i] = __uint_as_float(qnan);
    }
    if (py1 < m.rows) {
      const size_t i = (size_t)py1 * m.cols + px;
      if (best1 >= 0) m.img[i] = val1;
      else if (fill_nan) m.img[i] = __uint_as_float(qnan);
    }
  }
  if (lane == 0) m.tile_count[tile] = 0;  // fill cursor -> zero for the next projection
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

// ms_dev: device array of n MeshDev. The tile counters / list lengths re-arm themselves (k_tile_scan,
// k_tile_raster), so a projection needs no memset.
int raster_project(cudaStream_t st, const MeshDev* ms_dev, int n, int maxT, int max_ctas, int fill_nan) {
  if (n == 0 || max_ctas == 0) return 0;
  if (maxT > 0) {
    k_tri_setup<<<dim3((maxT + 255) / 256, n), 256, 0, st>>>(ms_dev);
    count_launch();
  }
  if (maxT > 0) {
    k_big_count<<<dim3(16, n), 256, 0, st>>>(ms_dev);
    count_launch();
  }
  k_tile_scan<<<dim3(1, n), 1024, 0, st>>>(ms_dev);
  count_launch();
  if (maxT > 0) {
    k_tri_fill<<<dim3((maxT + 255) / 256, n), 256, 0, st>>>(ms_dev);
    k_big_fill<<<dim3(16, n), 256, 0, st>>>(ms_dev);
    count_launch(2);
  }
  k_tile_raster<<<dim3(max_ctas, n), 128, 0, st>>>(ms_dev, fill_nan);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

int raster_tile_count(int rows, int cols) {
  return ((cols + TILE_W - 1) / TILE_W) * ((rows + TILE_H - 1) / TILE_H);
}
int raster_cta_count(int rows, int cols) {
  const int tiles_x = (cols + TILE_W - 1) / TILE_W;
  return ((tiles_x + CTA_TILES - 1) / CTA_TILES) * ((rows + TILE_H - 1) / TILE_H);
}

}  // namespace fb
