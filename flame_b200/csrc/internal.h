// Internal declarations shared by the translation units of libflame_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "flame_b200.h"

namespace fb {

void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_kernel_launches;
inline void count_launch(int n = 1) { g_kernel_launches.fetch_add(n, std::memory_order_relaxed); }

#define FB_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::fb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
      return FLAME_B200_ECUDA;                                                              \
    }                                                                                       \
  } while (0)

// ---- device-side layout (see DESIGN.md "HBM layout") ----
// Vertices are stored in an INTERNAL order: Morton-sorted, cut into chunks of Vc consecutive
// vertices (one chunk per CTA), and sorted by descending degree inside each chunk so that the
// lanes of a warp walk equally long incidence lists. Each vertex keeps a PRIVATE copy of the
// state of its incident edges (q1..q3 and the static alpha, beta, dx, dy), one record per
// incidence, in ascending ORIGINAL edge id k = 0..deg-1. Incidence k of the i-th vertex of chunk c
// lives at slot chunk_base[c] + koff[c][k] + i, so a warp's k-th loads are contiguous 16-byte words.
struct GraphDev {
  int V, E, Vc, n_chunks, max_deg;
  float4* vprim;    // [V] x, w1, w2, data_term
  float4* vbar;     // [V] x_bar, w1_bar, w2_bar, 0   (current)
  float4* vbar2;    // [V] ping-pong partner for the streaming kernel
  float4* vprev;    // [V] x_prev, w1_prev, w2_prev, 0
  float* vwt;       // [V] data_weight
  float4* inc_q;    // [2E] q1, q2, q3, bits((neighbour internal index << 1) | is_target)
  float4* inc_s;    // [2E] alpha, beta, dx, dy   (dx = pos_source.x - pos_target.x)
  int* deg;         // [V] internal order (device-built: the layout kernels' cursors)
  int* koff;        // [n_chunks][max_deg + 1]
  int* chunk_base;  // [n_chunks + 1]
  int* vperm;       // [V] internal -> caller vertex index
  int* vinv;        // [V] caller -> internal vertex index                      (device-built)
  int* inc_nbr;     // [2E] (neighbour internal index << 1) | is_target          (device-built)
  int* inc_edge;    // [2E] caller edge index of the incidence                   (device-built)
  int* edges;       // [2E] caller (source, target) pairs
  int* edge_slot;   // [2E] (caller edge, role) -> incidence slot                (device-built)
};

struct __align__(16) TriStatic {  // per triangle, written by k_tri_static; 80 bytes; depends on topology + positions
  int A1, B1, C1, A2;             // integer edge functions w1 = e23, w2 = e31, w3 = e12
  int B2, C2, A3, B3;
  int C3, xmin, xend, ymin;       // sweep columns [xmin, xend) (4-px groups), rows [ymin, ymax]
  int ymax, fx1;                  // last column the triangle may write (before clipping to the image)
  int flags;                      // 1 writes nothing, 2 sweep wraps past the right border, 4 degenerate, 8 wide list
  int i1;                         // caller vertex ids of p1, p2, p3 (= triangle's vertices in reverse order): the
  int i2, i3, pad0, pad1;         //     shading values are read without going through tris[]
};

struct MeshDev {
  int T, rows, cols, bin_cap;  // bin_cap: triangle ids a screen tile's bin holds
  const int* tris;             // 3T caller vertex indices
  const unsigned char* tri_ok; // T: tri_valid && all three vtx_valid (nullable)
  const unsigned char* tri_filt; // T: display filters of flame_b200_mesh_filter (nullable)
  const float* pos;            // 2V caller order
  float* values;               // V caller order
  float* img;                  // rows*cols
  TriStatic* stat;             // T
  int* counters;               // [0] wide-list length; followed in memory by tile_cnt (zeroed together)
  int* tile_cnt;               // n_tiles: triangles that asked for the tile's bin (may exceed bin_cap)
  int* bins;                   // n_tiles x bin_cap triangle ids
  int* wide;                   // <= T triangles reachable from every tile: many tiles, wrapping, or a full bin
};

// ---- PD (pd.cu) ----
struct ResidentPlan {
  int cluster;        // CTAs per graph (0 = does not fit -> streaming)
  int variant;        // index of the (threads, edges/thread, vertices/thread) build of k_pd_resident
  int threads;
  size_t smem_bytes;
  int Vc, maxInc, maxEdges, max_deg;
};
#define FB_RESIDENT_MAX_DEG 64
#define FB_RESIDENT_VARIANTS 3

int pd_pack(cudaStream_t st, const GraphDev& g, const float* raw_f, int alpha_on_device, int fresh);
int pd_unpack(cudaStream_t st, const GraphDev& g, float* raw_f);
int pd_unpack_xw(cudaStream_t st, const GraphDev& g, float* out /*3V: x | w1 | w2, caller order*/);
int pd_resync_gather(cudaStream_t st, float* raw_f, int V, int E, const float* carry, int Vold,
                     int Eold, const int* vtx_from, const int* edge_from, int has_init, int sticky);
int pd_set_data(cudaStream_t st, const GraphDev& g, const float* data_term, const float* data_weight);
int pd_reset(cudaStream_t st, const GraphDev* gs_dev, int n, int maxV, int maxE);
int pd_run_stream(cudaStream_t st, const GraphDev* gs_dev, int n, int max_chunks, int threads,
                  const flame_b200_pd_params& p, int n_iters);
size_t pd_resident_smem(int Vc, int maxInc, int maxEdges, int max_deg);
int pd_resident_variant(int Vc, int maxEdges, int* threads);  // -1: none fits
int pd_run_resident(cudaStream_t st, const GraphDev* gs_dev, int n, const ResidentPlan& plan,
                    const flame_b200_pd_params& p, int n_iters);
int pd_resident_init();  // function attributes, once per device
int pd_resident_capacity(const ResidentPlan& plan, int* n_clusters);
int pd_cost(cudaStream_t st, const GraphDev& g, const flame_b200_pd_params& p, int exact_order,
            float* terms_scratch, double* out_dev /*2 doubles*/);

// raw staging layout (floats), caller order: see capi.cu
enum RawF {
  RF_POS = 0,  // 2V
  RF_X, RF_W1, RF_W2, RF_XB, RF_W1B, RF_W2B, RF_XP, RF_W1P, RF_W2P, RF_DATA, RF_WT,  // V each
  RF_ALPHA, RF_BETA, RF_Q1, RF_Q2, RF_Q3,                                           // E each
  RF_COUNT
};
inline size_t raw_f_offset(int field, int V, int E) {
  // pos first (2V), then 11 V-arrays, then 5 E-arrays
  if (field == RF_POS) return 0;
  if (field <= RF_WT) return (size_t)2 * V + (size_t)(field - RF_X) * V;
  return (size_t)2 * V + (size_t)11 * V + (size_t)(field - RF_ALPHA) * E;
}
inline size_t raw_f_total(int V, int E) { return (size_t)13 * V + (size_t)5 * E; }

// ---- dense projection (raster.cu) ----
int raster_project(cudaStream_t st, const MeshDev* ms_dev, int n, int maxT, int max_rows, int max_cols, int fill_nan,
                   int rebuild);
int raster_tiles(int rows, int cols);  // screen tiles of an image (size of tile_cnt)
int raster_bin_cap(int T, int rows, int cols);
int raster_tri_filter(cudaStream_t st, const MeshDev& m, const flame_b200_tri_filter_params& p,
                      const float Kinv[9], unsigned char* out, float* angle_out = nullptr);
int raster_coverage(cudaStream_t st, const MeshDev& m, unsigned long long* d_count);
int raster_vertex_normals(cudaStream_t st, const MeshDev& m, int V, const float Kinv[9], int* scratch,
                          float* d_normals);
// values[vperm[v]] = vprim[v].x * scale  (vtx_idepths_ = x * graph_scale_, flame.cc:376)
int raster_values_from_graph(cudaStream_t st, const GraphDev* gs_dev, const MeshDev* ms_dev, int n,
                             int maxV, float scale);
#define FB_MAX_IMAGE_SIDE 2896  // |edge function| < 2^24: the reference's fp32 evaluation is exact up to here

}  // namespace fb
