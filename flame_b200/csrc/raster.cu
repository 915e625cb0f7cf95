// Dense per-pixel projection of the Delaunay mesh (inverse-depth map) for sm_100a.
//
// Replaces flame::utils::interpolateMesh (src/flame/utils/image_utils.cc:373-396) over
// flame::utils::DrawShadedTriangleBarycentric (src/flame/utils/rasterization.cc:164-245).
//
// Reference semantics reproduced exactly:
//  - vertices are rounded to integer pixels with round-half-to-even (cv::Point2f -> cv::Point);
//  - the triangle is drawn with its vertices REVERSED (Triangle emits CW, the rasteriser wants CCW);
//  - a pixel is covered when all three edge functions are >= 0 (inclusive: shared edges and
//    vertices are written by every covering triangle) — LAST triangle in the list wins;
//  - value = (v1*w1 + (v2*w2 + v3*w3)) / (w1 + (w2 + w3)), one IEEE divide, this association;
//  - the reference sweeps the bounding box 4 pixels at a time and addresses the image LINEARLY
//    (p.y*cols + p.x), so a degenerate (zero-area) triangle, whose edge functions vanish on a
//    whole line, also writes NaN up to 3 pixels past xmax (possibly wrapping to the next row).
//
// Edge functions are integer-valued and exact in fp32 for images up to ~2896 px on a side
// (|A*x + B*y + C| < 2^24), so direct evaluation equals the reference's incremental stepping.
//
// Round-1 implementation: "last triangle wins" is resolved with one 64-bit atomicMax per covered
// pixel on key = (triangle_index + 1) << 32 | value_bits, then a resolve pass writes the image.
#include <algorithm>

#include "internal.h"

namespace fb {

__device__ __forceinline__ int min3(int a, int b, int c) { return min(a, min(b, c)); }
__device__ __forceinline__ int max3(int a, int b, int c) { return max(a, max(b, c)); }

// Edge::init (rasterization.h:129-153): A*x + B*y + C with separately rounded mul/add.
struct EdgeFn {
  float A, B, C;
  __device__ __forceinline__ void init(int v0x, int v0y, int v1x, int v1y) {
    A = (float)(v1y - v0y);
    B = (float)(v0x - v1x);
    C = (float)(v1x * v0y - v0x * v1y);
  }
  __device__ __forceinline__ float at(int x, int y) const {
    return __fadd_rn(__fadd_rn(__fmul_rn(A, (float)x), __fmul_rn(B, (float)y)), C);
  }
};

// One warp per triangle; lanes stride over the (4-px padded) bounding box.
__global__ void __launch_bounds__(256)
k_raster_scatter(MeshDev m, const float* __restrict__ pos, const float* __restrict__ values) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= m.T) return;
  const int t = warp;
  if (m.tri_ok && !m.tri_ok[t]) return;
  const int ia = m.tris[3 * t], ib = m.tris[3 * t + 1], ic = m.tris[3 * t + 2];
  // image_utils.cc:385-391: p1 = vertices[tri[2]], p2 = vertices[tri[1]], p3 = vertices[tri[0]]
  const int p1x = __float2int_rn(pos[2 * ic]), p1y = __float2int_rn(pos[2 * ic + 1]);
  const int p2x = __float2int_rn(pos[2 * ib]), p2y = __float2int_rn(pos[2 * ib + 1]);
  const int p3x = __float2int_rn(pos[2 * ia]), p3y = __float2int_rn(pos[2 * ia + 1]);
  const float v1 = values[ic], v2 = values[ib], v3 = values[ia];

  const int xmin = min3(p1x, p2x, p3x), ymin = min3(p1y, p2y, p3y);
  const int xmax = max3(p1x, p2x, p3x), ymax = max3(p1y, p2y, p3y);
  EdgeFn e23, e31, e12;
  e23.init(p2x, p2y, p3x, p3y);  // w1
  e31.init(p3x, p3y, p1x, p1y);  // w2
  e12.init(p1x, p1y, p2x, p2y);  // w3

  const int wg = ((xmax - xmin) / 4 + 1) * 4;  // columns visited by the 4-wide sweep
  const int hg = ymax - ymin + 1;
  const long long npix = (long long)m.rows * m.cols;
  const unsigned long long key_hi = (unsigned long long)(t + 1) << 32;
  for (int idx = lane; idx < wg * hg; idx += 32) {
    const int y = ymin + idx / wg;
    const int x = xmin + idx % wg;
    const float w1 = e23.at(x, y), w2 = e31.at(x, y), w3 = e12.at(x, y);
    if (w1 >= 0.0f && w2 >= 0.0f && w3 >= 0.0f) {
      const float norm = __fadd_rn(w1, __fadd_rn(w2, w3));
      const float num =
          __fadd_rn(__fmul_rn(v1, w1), __fadd_rn(__fmul_rn(v2, w2), __fmul_rn(v3, w3)));
      const float val = __fdiv_rn(num, norm);
      const long long addr = (long long)y * m.cols + x;  // linear, as the reference's pointer math
      if (addr >= 0 && addr < npix)
        atomicMax(&m.win[addr], key_hi | (unsigned long long)__float_as_uint(val));
    }
  }
}

__global__ void __launch_bounds__(256) k_raster_resolve(MeshDev m, int fill_nan) {
  const long long n = (long long)m.rows * m.cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = m.win[i];
    if (k != 0ull) {
      m.img[i] = __uint_as_float((unsigned int)(k & 0xffffffffull));
    } else if (fill_nan) {
      m.img[i] = __uint_as_float(0x7fc00000u);  // std::numeric_limits<float>::quiet_NaN()
    }
  }
}

__global__ void k_values_from_graph(GraphDev g, float scale, float* __restrict__ values) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  values[g.vperm[v]] = __fmul_rn(g.vprim[v].x, scale);  // flame.cc:376
}

int raster_values_from_graph(cudaStream_t st, const GraphDev& g, float scale, float* values) {
  if (g.V == 0) return 0;
  k_values_from_graph<<<(g.V + 255) / 256, 256, 0, st>>>(g, scale, values);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

int raster_project(cudaStream_t st, const MeshDev& m, const float* pos, const float* values,
                   int fill_nan) {
  const size_t n = (size_t)m.rows * m.cols;
  FB_CUDA(cudaMemsetAsync(m.win, 0, n * sizeof(unsigned long long), st));
  if (m.T > 0) {
    const int warps_per_block = 256 / 32;
    k_raster_scatter<<<(m.T + warps_per_block - 1) / warps_per_block, 256, 0, st>>>(m, pos, values);
    count_launch();
  }
  const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
  k_raster_resolve<<<blocks, 256, 0, st>>>(m, fill_nan);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb
