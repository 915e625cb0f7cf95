// NLTGV2-L1 primal-dual (Chambolle-Pock) solver kernels for sm_100a.
//
// Replaces flame::optimizers::nltgv2_l1_graph_regularizer::step and friends
// (src/flame/optimizers/nltgv2_l1_graph_regularizer.{h,cc}). See DESIGN.md for the layout and
// the roofline model. Arithmetic reproduces the reference's scalar fp32 evaluation order
// bit-for-bit: every multiply/add below is an explicitly rounded __fmul_rn/__fadd_rn (the
// reference's x86-64 Release build has no FMA), and the per-vertex gather walks the
// incident edges in ascending ORIGINAL edge id, which is the order the reference's sequential
// edge scatter (cc:118-142) applies its rounded += / -= to each vertex.
//
// Three kernels share the arithmetic helpers below:
//   k_pd_stream       state in HBM, one launch per iteration, VERTEX-CENTRIC: one thread owns one vertex
//                     and a private copy of the dual variables of its incident edges; per iteration it
//                     recomputes the dual update of every incident edge from the x_bar/w_bar of the two
//                     endpoints (both endpoints compute the same rounded values, so the two copies stay
//                     bit-identical), applies the primal contributions in order, then proxL1 + the
//                     extragradient step into the other half of a double buffer. Any graph size.
//   k_pd_resident     one thread-block cluster per graph, state on chip for all iterations: edge-parallel
//                     dual phase (q in registers, contributions pushed to both endpoints' shared-memory
//                     slots, remote ones through DSMEM), vertex-parallel primal phase, two cluster barriers
//                     per iteration. The default whenever the graph fits a cluster.
//   k_pd_resident_mb  the same with local halo copies and point-to-point mbarrier synchronisation
//                     (experimental, slower on B200; FLAME_B200_PD_SYNC=p2p).
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "internal.h"

namespace cg = cooperative_groups;

namespace fb {

// ------------------------------------------------------------------------------------------
// Exact-arithmetic helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

// proxNLTGV2Conj (h:171-176): q / (|q| > 1 ? |q| : 1). For finite q, q/|q| is exactly +-1 and
// q/1 is exactly q, so the quotient IS this clamp — two FMNMX, no divide. (A NaN q, on which the
// reference FLAME_ASSERTs and exits, is not propagated by fminf/fmaxf; inputs are validated finite.)
__device__ __forceinline__ float prox_conj(float q) { return fminf(fmaxf(q, -1.0f), 1.0f); }

// dualStep body for one edge (cc:98-110). bi/bj = (x_bar, w1_bar, w2_bar, -) of source/target.
__device__ __forceinline__ float4 dual_update(float4 q, const float4 s, const float4 bi,
                                              const float4 bj, const float step_q) {
  const float alpha = s.x, beta = s.y, dx = s.z, dy = s.w;
  float K1x = mul(alpha, sub(bi.x, bj.x));
  K1x = sub(K1x, mul(mul(alpha, dx), bi.y));
  K1x = sub(K1x, mul(mul(alpha, dy), bi.z));
  q.x = prox_conj(add(q.x, mul(step_q, K1x)));
  const float K2x = mul(beta, sub(bi.y, bj.y));
  q.y = prox_conj(add(q.y, mul(step_q, K2x)));
  const float K3x = mul(beta, sub(bi.z, bj.z));
  q.z = prox_conj(add(q.z, mul(step_q, K3x)));
  return q;
}

// dualStep for one incidence: `own` is this vertex, `nb` the neighbour; the edge's source is `nb`
// when this vertex is the target. Selecting the operands first keeps ONE copy of the arithmetic.
__device__ __forceinline__ float4 dual_update_inc(const float4 q, const float4 s, const float4 own,
                                                  const float4 nb, const int is_target,
                                                  const float step_q) {
  float4 bi, bj;
  bi.x = is_target ? nb.x : own.x;
  bi.y = is_target ? nb.y : own.y;
  bi.z = is_target ? nb.z : own.z;
  bj.x = is_target ? own.x : nb.x;
  bj.y = is_target ? own.y : nb.y;
  bj.z = is_target ? own.z : nb.z;
  bi.w = bj.w = 0.0f;
  return dual_update(q, s, bi, bj, step_q);
}

// primalStep edge-loop body (cc:125-141) as seen from ONE endpoint.
__device__ __forceinline__ void primal_apply(float& x, float& w1, float& w2, const float4 q,
                                             const float4 s, const int is_target,
                                             const float step_x) {
  const float a = mul(mul(q.x, step_x), s.x);
  const float b2 = mul(mul(q.y, step_x), s.y);
  const float b3 = mul(mul(q.z, step_x), s.y);
  if (!is_target) {
    x = sub(x, a);
    w1 = add(w1, mul(a, s.z));
    w2 = add(w2, mul(a, s.w));
    w1 = sub(w1, b2);
    w2 = sub(w2, b3);
  } else {
    x = add(x, a);
    w1 = add(w1, b2);
    w2 = add(w2, b3);
  }
}

// proxL1 (h:179-197) followed by extraGradientStep (cc:158-171) for one vertex.
struct VertexOut {
  float4 prim;  // x, w1, w2, data
  float4 bar;   // x_bar, w1_bar, w2_bar, 0
};
__device__ __forceinline__ VertexOut vertex_finish(float x, float w1, float w2, const float data,
                                                   const float wt, const float x_old,
                                                   const float w1_old, const float w2_old,
                                                   const flame_b200_pd_params& p) {
  const float thresh = mul(p.step_x, mul(p.data_factor, wt));
  const float diff = sub(x, data);
  float nx;
  if (diff > thresh) {
    nx = sub(x, thresh);
  } else if (diff < -thresh) {
    nx = add(x, thresh);
  } else {
    nx = data;
  }
  nx = (nx < p.x_min) ? p.x_min : nx;
  nx = (nx > p.x_max) ? p.x_max : nx;
  float xb = add(nx, mul(p.theta, sub(nx, x_old)));
  xb = (xb < p.x_min) ? p.x_min : xb;
  xb = (xb > p.x_max) ? p.x_max : xb;
  VertexOut o;
  o.prim = make_float4(nx, w1, w2, data);
  o.bar = make_float4(xb, add(w1, mul(p.theta, sub(w1, w1_old))),
                      add(w2, mul(p.theta, sub(w2, w2_old))), 0.0f);
  return o;
}

// ------------------------------------------------------------------------------------------
// pack / unpack between the caller-order SoA staging image and the internal layout
// ------------------------------------------------------------------------------------------
static inline int nblk(int n, int t) { return (n + t - 1) / t; }

// ------------------------------------------------------------------------------------------
// Device-side incidence layout. The host only orders the VERTICES (vperm, chunk tables: O(V)); the
// O(E) part - which incidence slot every (edge, endpoint) gets - is built here:
//   k_layout_begin    vinv = vperm^-1, per-vertex cursors = 0
//   k_layout_scatter  one thread per edge claims a slot at each endpoint (atomic cursor: arbitrary k)
//   k_layout_finish   one thread per vertex sorts its <= deg claimed (edge, role) keys into ascending
//                     ORIGINAL edge id - the order in which the reference's sequential edge loop
//                     (cc:118-142) applies its rounded += / -= to that vertex - and writes the
//                     incidence records (q, alpha/beta/dx/dy, neighbour, edge <-> slot maps) and the
//                     vertex records.
// The final layout does not depend on the order in which the atomics were served.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ const int* koff_row(const GraphDev& g, int ch) {
  return g.koff + (size_t)ch * (g.max_deg + 1);
}
__device__ __forceinline__ int inc_slot(const GraphDev& g, int v, int k) {
  const int ch = v / g.Vc;
  return g.chunk_base[ch] + koff_row(g, ch)[k] + (v - ch * g.Vc);
}

__global__ void k_layout_begin(GraphDev g) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  g.vinv[g.vperm[v]] = v;
  g.deg[v] = 0;
}

__global__ void k_layout_scatter(GraphDev g) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= g.E) return;
  const int s = g.vinv[g.edges[2 * o]], d = g.vinv[g.edges[2 * o + 1]];
  g.inc_edge[inc_slot(g, s, atomicAdd(&g.deg[s], 1))] = 2 * o;      // key = (edge << 1) | is_target
  g.inc_edge[inc_slot(g, d, atomicAdd(&g.deg[d], 1))] = 2 * o + 1;
}

// fresh != 0: only pos / data_term / data_weight were uploaded; the rest follows the fresh-vertex
// rule of Flame::syncGraph (flame.cc:2039-2048): x = x_bar = x_prev = data_term, w* = 0, q* = 0, beta = 1.
__global__ void k_layout_finish(GraphDev g, const float* __restrict__ raw, int alpha_on_device, int fresh) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  const int V = g.V;
  const size_t E = g.E;
  {  // vertex records
    const int o = g.vperm[v];
    const float* b = raw + 2 * (size_t)V;
    if (fresh) {
      const float d = b[9 * V + o];
      g.vprim[v] = make_float4(d, 0.0f, 0.0f, d);
      g.vbar[v] = make_float4(d, 0.0f, 0.0f, 0.0f);
      g.vprev[v] = make_float4(d, 0.0f, 0.0f, 0.0f);
    } else {
      g.vprim[v] = make_float4(b[0 * V + o], b[1 * V + o], b[2 * V + o], b[9 * V + o]);
      g.vbar[v] = make_float4(b[3 * V + o], b[4 * V + o], b[5 * V + o], 0.0f);
      g.vprev[v] = make_float4(b[6 * V + o], b[7 * V + o], b[8 * V + o], 0.0f);
    }
    g.vwt[v] = b[10 * V + o];
  }
  const int ch = v / g.Vc, i = v - ch * g.Vc;
  const int base = g.chunk_base[ch] + i;
  const int* __restrict__ koff = koff_row(g, ch);
  const int d = g.deg[v];
  // sort the claimed keys (typical degree 6); element k lives at base + koff[k]. Up to 24 keys are fetched with
  // independent loads and insertion-sorted in a thread-local array (one global round trip instead of one per
  // comparison); hubs sort in place in global memory.
  constexpr int LOCAL_KEYS = 24;
  if (d <= LOCAL_KEYS) {
    int keys[LOCAL_KEYS];
#pragma unroll
    for (int k = 0; k < LOCAL_KEYS; ++k) keys[k] = k < d ? g.inc_edge[base + koff[k]] : 0x7fffffff;
    for (int k = 1; k < d; ++k) {
      const int key = keys[k];
      int j = k - 1;
      while (j >= 0 && keys[j] > key) {
        keys[j + 1] = keys[j];
        --j;
      }
      keys[j + 1] = key;
    }
    for (int k = 0; k < d; ++k) g.inc_edge[base + koff[k]] = keys[k];
  } else {
    for (int k = 1; k < d; ++k) {
      const int key = g.inc_edge[base + koff[k]];
      int j = k - 1;
      while (j >= 0) {
        const int kj = g.inc_edge[base + koff[j]];
        if (kj <= key) break;
        g.inc_edge[base + koff[j + 1]] = kj;
        --j;
      }
      g.inc_edge[base + koff[j + 1]] = key;
    }
  }
  const float* eb = raw + 13 * (size_t)V;
  // eight incidences at a time, staged (keys -> endpoints -> positions / neighbour index -> records) so that the
  // dependent global loads of a stage are all in flight together
  for (int k0 = 0; k0 < d; k0 += 8) {
    int slot[8], o[8], os[8], od[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      slot[u] = base + koff[min(k0 + u, d - 1)];
      o[u] = g.inc_edge[slot[u]];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      os[u] = g.edges[2 * (o[u] >> 1)];
      od[u] = g.edges[2 * (o[u] >> 1) + 1];
    }
    float dx[8], dy[8];
    int ent[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      dx[u] = sub(raw[2 * os[u]], raw[2 * od[u]]);          // pos_i.x - pos_j.x (cc:100)
      dy[u] = sub(raw[2 * os[u] + 1], raw[2 * od[u] + 1]);
      const int role = o[u] & 1;
      ent[u] = (g.vinv[role ? os[u] : od[u]] << 1) | role;  // (neighbour internal index << 1) | is_target
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (k0 + u >= d) break;
      const int e = o[u] >> 1, role = o[u] & 1;
      float alpha;
      if (alpha_on_device) {
        // flame.cc:2089-2092, 2102: alpha = 1 / sqrt(dx*dx + dy*dy)
        const float len = __fsqrt_rn(add(mul(dx[u], dx[u]), mul(dy[u], dy[u])));
        alpha = __fdiv_rn(1.0f, len);
      } else {
        alpha = eb[0 * E + e];
      }
      if (fresh) {
        g.inc_s[slot[u]] = make_float4(alpha, 1.0f, dx[u], dy[u]);
        g.inc_q[slot[u]] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(ent[u]));
      } else {
        g.inc_s[slot[u]] = make_float4(alpha, eb[1 * E + e], dx[u], dy[u]);
        g.inc_q[slot[u]] = make_float4(eb[2 * E + e], eb[3 * E + e], eb[4 * E + e], __int_as_float(ent[u]));
      }
      g.inc_nbr[slot[u]] = ent[u];
      g.inc_edge[slot[u]] = e;
      g.edge_slot[2 * e + role] = slot[u];  // inverse map: (edge, role) -> incidence slot
    }
  }
}

__global__ void k_unpack_vertices(GraphDev g, float* __restrict__ raw) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  const int o = g.vperm[v];
  const int V = g.V;
  float* b = raw + 2 * (size_t)V;
  const float4 p = g.vprim[v], br = g.vbar[v], pv = g.vprev[v];
  b[0 * V + o] = p.x; b[1 * V + o] = p.y; b[2 * V + o] = p.z;
  b[3 * V + o] = br.x; b[4 * V + o] = br.y; b[5 * V + o] = br.z;
  b[6 * V + o] = pv.x; b[7 * V + o] = pv.y; b[8 * V + o] = pv.z;
  b[9 * V + o] = p.w;
  b[10 * V + o] = g.vwt[v];
}

// every edge has exactly one source-role incidence: that copy is written back
__global__ void k_unpack_incidences(GraphDev g, float* __restrict__ raw) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= 2 * g.E) return;
  const float4 q = g.inc_q[slot];
  if (__float_as_int(q.w) & 1) return;
  const int o = g.inc_edge[slot];
  const size_t E = g.E;
  float* b = raw + 13 * (size_t)g.V;
  const float4 s = g.inc_s[slot];
  b[0 * E + o] = s.x; b[1 * E + o] = s.y;
  b[2 * E + o] = q.x; b[3 * E + o] = q.y; b[4 * E + o] = q.z;
}

__global__ void k_set_data(GraphDev g, const float* __restrict__ data, const float* __restrict__ wt) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  const int o = g.vperm[v];
  if (data) g.vprim[v].w = data[o];
  if (wt) g.vwt[v] = wt[o];
}

// Fresh-graph rule of Flame::syncGraph (flame.cc:2039-2048): x = x_bar = x_prev = data_term,
// w = w_bar = w_prev = 0, q = 0. blockIdx.y selects the graph of a batch.
__global__ void k_graph_reset(const GraphDev* __restrict__ gs) {
  const GraphDev g = gs[blockIdx.y];
  const int n = max(g.V, 2 * g.E);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < g.V) {
      const float d = g.vprim[i].w;
      g.vprim[i] = make_float4(d, 0.0f, 0.0f, d);
      g.vbar[i] = make_float4(d, 0.0f, 0.0f, 0.0f);
      g.vprev[i] = make_float4(d, 0.0f, 0.0f, 0.0f);
    }
    if (i < 2 * g.E) {
      const float w = g.inc_q[i].w;
      g.inc_q[i] = make_float4(0.0f, 0.0f, 0.0f, w);
    }
  }
}

// Device-side Flame::syncGraph state carry-over (flame.cc:1984-2048, 2084-2105): builds the raw
// (caller-order) state of the NEW graph from the unpacked state of the OLD one.
//   carried vertex (vtx_from >= 0): x, w, x_bar, w_bar, x_prev, w_prev are kept (:2003-2008 touch only
//     pos / data_term / data_weight), optional sticky-obstacle reset of x (:2016-2019);
//   new vertex: x = x_bar = x_prev = x_init (or data_term), w = 0 (:2046-2048, VertexData defaults);
//   carried edge: q kept; new edge: q = 0 (EdgeData defaults); beta = 1 for all (:2103); alpha is
//     recomputed from the new positions by the pack kernel (:2102).
__global__ void k_resync_gather(float* __restrict__ raw, int V, int E, const float* __restrict__ carry,
                                int Vold, int Eold, const int* __restrict__ vtx_from,
                                const int* __restrict__ edge_from, int has_init, int sticky) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V) {
    float* b = raw + 2 * (size_t)V;
    const float* bo = carry + 2 * (size_t)Vold;
    const int from = vtx_from[i];
    const float data = b[9 * (size_t)V + i];
    if (from >= 0) {
      float x = bo[from];
      if (sticky && sub(x, data) > 0.25f) x = data;
      b[i] = x;
#pragma unroll
      for (int f = 1; f < 9; ++f) b[f * (size_t)V + i] = bo[f * (size_t)Vold + from];
    } else {
      const float x = has_init ? b[i] : data;
      b[0 * (size_t)V + i] = x; b[1 * (size_t)V + i] = 0.0f; b[2 * (size_t)V + i] = 0.0f;
      b[3 * (size_t)V + i] = x; b[4 * (size_t)V + i] = 0.0f; b[5 * (size_t)V + i] = 0.0f;
      b[6 * (size_t)V + i] = x; b[7 * (size_t)V + i] = 0.0f; b[8 * (size_t)V + i] = 0.0f;
    }
  }
  if (i < E) {
    float* e = raw + 13 * (size_t)V;
    const float* eo = carry + 13 * (size_t)Vold;
    const int from = edge_from[i];
    e[1 * (size_t)E + i] = 1.0f;
    e[2 * (size_t)E + i] = from >= 0 ? eo[2 * (size_t)Eold + from] : 0.0f;
    e[3 * (size_t)E + i] = from >= 0 ? eo[3 * (size_t)Eold + from] : 0.0f;
    e[4 * (size_t)E + i] = from >= 0 ? eo[4 * (size_t)Eold + from] : 0.0f;
  }
}

int pd_resync_gather(cudaStream_t st, float* raw_f, int V, int E, const float* carry, int Vold,
                     int Eold, const int* vtx_from, const int* edge_from, int has_init, int sticky) {
  const int n = V > E ? V : E;
  if (n == 0) return 0;
  k_resync_gather<<<nblk(n, 256), 256, 0, st>>>(raw_f, V, E, carry, Vold, Eold, vtx_from, edge_from,
                                                has_init, sticky);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

int pd_pack(cudaStream_t st, const GraphDev& g, const float* raw_f, int alpha_on_device, int fresh) {
  if (g.V == 0) return 0;
  k_layout_begin<<<nblk(g.V, 256), 256, 0, st>>>(g);
  count_launch();
  if (g.E > 0) {
    k_layout_scatter<<<nblk(g.E, 256), 256, 0, st>>>(g);
    count_launch();
  }
  k_layout_finish<<<nblk(g.V, 128), 128, 0, st>>>(g, raw_f, alpha_on_device || fresh, fresh);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

int pd_unpack(cudaStream_t st, const GraphDev& g, float* raw_f) {
  if (g.V > 0) {
    k_unpack_vertices<<<nblk(g.V, 256), 256, 0, st>>>(g, raw_f);
    count_launch();
  }
  if (g.E > 0) {
    k_unpack_incidences<<<nblk(2 * g.E, 256), 256, 0, st>>>(g, raw_f);
    count_launch();
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

// x, w1, w2 only, caller order, into out[0..3V) - what Flame::update reads back per frame (flame.cc:372-380)
__global__ void k_unpack_xw(GraphDev g, float* __restrict__ out) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  const int o = g.vperm[v];
  const float4 p = g.vprim[v];
  out[o] = p.x;
  out[g.V + o] = p.y;
  out[2 * (size_t)g.V + o] = p.z;
}

int pd_unpack_xw(cudaStream_t st, const GraphDev& g, float* out) {
  if (g.V == 0) return 0;
  k_unpack_xw<<<nblk(g.V, 256), 256, 0, st>>>(g, out);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

int pd_set_data(cudaStream_t st, const GraphDev& g, const float* data_term, const float* data_weight) {
  if (g.V > 0) {
    k_set_data<<<nblk(g.V, 256), 256, 0, st>>>(g, data_term, data_weight);
    count_launch();
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

int pd_reset(cudaStream_t st, const GraphDev* gs_dev, int n, int maxV, int maxE) {
  const int m = std::max(maxV, 2 * maxE);
  if (m == 0 || n == 0) return 0;
  k_graph_reset<<<dim3(std::min(nblk(m, 256), 1024), n), 256, 0, st>>>(gs_dev);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// STREAM mode: state in HBM / L2, ONE kernel per iteration (x_bar double-buffered), any size.
// grid = (chunks, graphs); a CTA sweeps one chunk of Vc vertices.
// ------------------------------------------------------------------------------------------
#define STREAM_T 128
// grid = (chunks * ceil(Vc / STREAM_T), graphs): warps are independent (no CTA-level sync), so small
// CTAs keep the occupancy up; incidences are processed in groups of STREAM_U with all loads of a
// group (private q / static data, then the neighbour gathers) issued before its arithmetic.
template <int STREAM_U, int MINB>
__global__ void __launch_bounds__(STREAM_T, MINB)
k_pd_stream(const GraphDev* __restrict__ gs, flame_b200_pd_params p, int parity, int write_prev,
            int ctas_per_chunk) {
  const GraphDev g = gs[blockIdx.y];
  const int c = blockIdx.x / ctas_per_chunk;
  if (c >= g.n_chunks) return;
  const float4* __restrict__ cur = parity ? g.vbar2 : g.vbar;
  float4* __restrict__ nxt = parity ? g.vbar : g.vbar2;
  const int v0 = c * g.Vc, nv = min(g.Vc, g.V - v0);
  const int i = (blockIdx.x % ctas_per_chunk) * STREAM_T + threadIdx.x;
  if (i >= nv) return;
  const int base = g.chunk_base[c];
  const int* __restrict__ koff = koff_row(g, c);
  const int v = v0 + i;
  const float4 pr = g.vprim[v];
  const float4 own = cur[v];
  const float wt = g.vwt[v];
  float x = pr.x, w1 = pr.y, w2 = pr.z;
  const int d = g.deg[v];
  for (int k0 = 0; k0 < d; k0 += STREAM_U) {
    float4 q[STREAM_U], s[STREAM_U], nb[STREAM_U];
    int slot[STREAM_U];
#pragma unroll
    for (int u = 0; u < STREAM_U; ++u)
      if (k0 + u < d) {
        slot[u] = base + koff[k0 + u] + i;
        q[u] = __ldcs(&g.inc_q[slot[u]]);  // streamed once per iteration: do not displace the gathered x_bar
        s[u] = __ldcs(&g.inc_s[slot[u]]);
      }
#pragma unroll
    for (int u = 0; u < STREAM_U; ++u)
      if (k0 + u < d) nb[u] = __ldg(&cur[__float_as_int(q[u].w) >> 1]);
#pragma unroll
    for (int u = 0; u < STREAM_U; ++u)
      if (k0 + u < d) {
        const int is_target = __float_as_int(q[u].w) & 1;
        const float4 qn = dual_update_inc(q[u], s[u], own, nb[u], is_target, p.step_q);
        primal_apply(x, w1, w2, qn, s[u], is_target, p.step_x);
        __stcs(&g.inc_q[slot[u]], make_float4(qn.x, qn.y, qn.z, q[u].w));
      }
  }
  const VertexOut o = vertex_finish(x, w1, w2, pr.w, wt, pr.x, pr.y, pr.z, p);
  g.vprim[v] = o.prim;
  nxt[v] = o.bar;
  if (write_prev) g.vprev[v] = make_float4(pr.x, pr.y, pr.z, 0.0f);
}

__global__ void k_copy_bar(const GraphDev* __restrict__ gs) {
  const GraphDev g = gs[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < g.V; i += gridDim.x * blockDim.x)
    g.vbar[i] = g.vbar2[i];
}

int pd_run_stream(cudaStream_t st, const GraphDev* gs_dev, int n, int max_chunks, int threads,
                  const flame_b200_pd_params& p, int n_iters) {
  if (n_iters <= 0 || n <= 0 || max_chunks <= 0) return 0;
  const int ctas_per_chunk = (threads + STREAM_T - 1) / STREAM_T;  // threads = max chunk width (rounded)
  const dim3 grid(max_chunks * ctas_per_chunk, n);
  int variant = 2;  // 64 registers, 8 CTAs/SM measured best on B200 (FLAME_B200_STREAM_U=4|46|3 for experiments)
  if (const char* sv = getenv("FLAME_B200_STREAM_U")) variant = atoi(sv);
  for (int it = 0; it < n_iters; ++it) {
    if (variant == 2)
      k_pd_stream<2, 8><<<grid, STREAM_T, 0, st>>>(gs_dev, p, it & 1, it == n_iters - 1, ctas_per_chunk);
    else if (variant == 46)
      k_pd_stream<4, 6><<<grid, STREAM_T, 0, st>>>(gs_dev, p, it & 1, it == n_iters - 1, ctas_per_chunk);
    else if (variant == 3)
      k_pd_stream<3, 6><<<grid, STREAM_T, 0, st>>>(gs_dev, p, it & 1, it == n_iters - 1, ctas_per_chunk);
    else
      k_pd_stream<4, 5><<<grid, STREAM_T, 0, st>>>(gs_dev, p, it & 1, it == n_iters - 1, ctas_per_chunk);
  }
  count_launch(n_iters);
  if (n_iters & 1) {  // result is in vbar2: bring it home
    k_copy_bar<<<dim3(std::min(max_chunks, 64), n), 256, 0, st>>>(gs_dev);
    count_launch();
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// RESIDENT mode: one thread-block cluster per graph; the PD state stays on chip for all n_iters
// iterations. Per iteration, two phases separated by cluster barriers:
//   A (edge-parallel, balanced): every EDGE is owned by one thread of the CTA that holds its source
//     vertex, with q1..q3 in REGISTERS for the whole solve. The thread reads x_bar/w_bar of both
//     endpoints (source: ld.shared; target: DSMEM through a pre-mapped shared::cluster address),
//     performs the dual update once, and PUSHES the rounded primal contributions of the edge to the
//     incidence slots of both endpoints: (-a, a*dx, a*dy, -b2 | -b3) to the source's slot in its own
//     shared memory and (a, b2, b3, -0 | -0) to the target's slot, which may live in another CTA of
//     the cluster (st.shared::cluster). "+ -0.0f" is the exact identity of IEEE addition, so the
//     target's slot needs no branch in phase B.
//   barrier.cluster (release/acquire: remote stores visible)
//   B (vertex-parallel): a thread owns up to VPT vertices (x, w1, w2 in registers); it adds the
//     contributions of their incidence slots in ascending ORIGINAL edge id — five rounded adds per
//     incidence, bit-identical to the reference's sequential edge scatter (cc:125-141) — then proxL1
//     and the extragradient step, and publishes x_bar/w_bar.
//   barrier.cluster
// Nobody reads x_bar between the two barriers, so ONE x_bar buffer suffices.
//
// The template parameters are the register budgets the kernel is built for: at most MAXT threads, SPT owned
// edges and VPT vertices per thread; the block size is chosen at launch (pd_resident_variant). <640,6,2> holds
// a 10 000-vertex Delaunay graph in a cluster of NINE CTAs with 576 threads each (two such clusters fit every
// GPC with >= 18 SMs: 15 frames in flight on 135 of the 148 SMs); <640,4,1> is the low-latency variant for
// 16-CTA clusters (one frame spread over 16 SMs); <1024,4,2> takes chunks up to 2 048 vertices.
// ------------------------------------------------------------------------------------------
struct ResidentArgs {
  const GraphDev* gs;
  flame_b200_pd_params p;
  int n_iters;
  int Vc, maxInc, maxEdges, max_deg;  // batch maxima: identical shared-memory carve in every CTA
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ void st_shared_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Shared-memory slots of the contribution records. In global memory incidence k of the chunk's i-th vertex
// is slot koff[k] + i; in shared memory the k-blocks start at multiples of FB_KPAD records (kpad[k] + i), so
// that the 16-byte bank group of a vertex's records is its index mod 8 for EVERY k: a quarter-warp working on
// eight consecutive vertices never conflicts, whichever incidences it touches.
#define FB_KPAD 8
#define FB_DEAL_INTS (72 + 64 + 64 + 72 + 72)  // dealing tables: supply, demand, list offsets, hole / overflow bases
__host__ __device__ inline int kpad_round(int n) { return (n + FB_KPAD - 1) & ~(FB_KPAD - 1); }
__host__ __device__ inline int resident_inc_records(int maxInc, int max_deg) {
  return ((maxInc + FB_KPAD * (max_deg + 1)) + 1) & ~1;
}
// The records double as load-phase scratch (two owned-edge lists, the per-vertex scan, two small rank
// tables): the region holds max(records, that scratch).
__host__ __device__ inline int resident_p4_records(int maxInc, int maxEdges, int Vc, int max_deg) {
  const int scratch = (2 * maxEdges + Vc + 1 + 2 * (max_deg + 2) + maxEdges + (maxEdges + 1) / 2 + FB_DEAL_INTS + 3) / 4;
  const int rec = resident_inc_records(maxInc, max_deg);
  return rec > scratch ? rec : scratch;
}

#ifdef FB_PROFILE_LOAD
__device__ unsigned long long fb_dbg[32];
#define FB_STAMP(n)                                                                  \
  do {                                                                               \
    __syncthreads();                                                                 \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                                       \
      unsigned long long t_;                                                         \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                         \
      fb_dbg[n] = t_;                                                                \
    }                                                                                \
  } while (0)
extern "C" int flame_b200_debug_stamps(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, fb_dbg, sizeof(fb_dbg));
}
#else
#define FB_STAMP(n)
#endif

template <int MAXT, int SPT, int VPT, int PD_MLP>
__global__ void __launch_bounds__(MAXT, 1) k_pd_resident(ResidentArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int c = cluster.block_rank();
  const GraphDev g = a.gs[blockIdx.x / C];
  const int tid = threadIdx.x;
  const int NT = blockDim.x;  // <= MAXT: the planner launches as many threads as balance the chunk (pd_resident_variant)
  const int Vc = g.Vc;  // this graph's chunk width (a.Vc >= Vc is the carve width)

  extern __shared__ __align__(128) unsigned char smem[];
  float4* s_bar = reinterpret_cast<float4*>(smem);  // x_bar, w1_bar, w2_bar of the chunk's vertices
  float4* s_p4 = s_bar + a.Vc;                      // per incidence slot: (sx, t1, t2, u1)
  float4* s_s = s_p4 + resident_p4_records(a.maxInc, a.maxEdges, a.Vc, a.max_deg);  // per owned edge: (alpha, beta, dx, dy)
  float* s_p1 = reinterpret_cast<float*>(s_s + a.maxEdges);                          // per incidence slot: u2
  int* s_kpad = reinterpret_cast<int*>(s_p1 + resident_inc_records(a.maxInc, a.max_deg));  // a.max_deg + 1 (+1 pad: read in pairs)
  int* s_koff = s_kpad + ((a.max_deg + 2) & ~1);                                     // a.max_deg + 1
  int* s_cnt = s_koff + a.max_deg + 1;                                               // 1
  // load phase only (aliases of the records):
  unsigned int* s_list = reinterpret_cast<unsigned int*>(s_p4);  // owned edges vertex by vertex (vertex << 16 | k)
  int* s_scan = reinterpret_cast<int*>(s_list + a.maxEdges);     // first owned edge of every vertex (a.Vc + 1)
  unsigned int* s_rank = reinterpret_cast<unsigned int*>(s_scan + a.Vc + 1);  // the same edges, rank-major
  int* s_hist = reinterpret_cast<int*>(s_rank + a.maxEdges);     // #vertices per owned-edge count (max_deg + 2)
  int* s_roff = s_hist + a.max_deg + 2;                          // first list position of every rank (max_deg + 2)
  int* s_tmp = s_roff + a.max_deg + 2;                           // per list position: (target group, supply index, demand index)
  int* s_sup = s_tmp + a.maxEdges;                               // [8][9] edges per (source group, target group | remote)
  int* s_dem = s_sup + 72;                                       // [8][8] positions per (source group, wanted target group)
  int* s_off = s_dem + 64;                                       // [8][8] start of every position list
  int* s_hb = s_off + 64;                                        // [8][9] first hole index of every position list
  int* s_ob = s_hb + 72;                                         // [8][9] first overflow index of every edge queue
  unsigned short* s_pos = reinterpret_cast<unsigned short*>(s_ob + 72);  // position lists (a.maxEdges)

  const bool have = c < g.n_chunks;
  const int v0 = c * Vc;
  const int nv = have ? min(Vc, g.V - v0) : 0;
  const int base = have ? g.chunk_base[c] : 0;
  const uint32_t bar_addr = smem_u32(s_bar);
  const uint32_t p4_addr = smem_u32(s_p4);
  const float neg0 = __int_as_float(0x80000000);

  FB_STAMP(0);
  // ---- load ----
  if (have)
    for (int k = tid; k <= g.max_deg; k += NT) s_koff[k] = koff_row(g, c)[k];
  for (int k = tid; k < g.max_deg + 2; k += NT) s_hist[k] = 0;
  for (int k = tid; k < 72 + 64; k += NT) s_sup[k] = 0;  // s_sup and s_dem are adjacent
  __syncthreads();
  if (tid == 0) {
    int tot = 0;
    for (int k = 0; k <= g.max_deg; ++k) {
      s_kpad[k] = tot;
      if (k < g.max_deg) tot += have ? kpad_round(s_koff[k + 1] - s_koff[k]) : 0;
    }
  }
  // vertex registers: thread t owns vertices t, t + NT, ... of the chunk (nv <= NT * VPT)
  float x[VPT], w1[VPT], w2[VPT], data[VPT], wt[VPT];
  int deg[VPT];
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    const int i = tid + v * NT;
    x[v] = w1[v] = w2[v] = data[v] = wt[v] = 0.f;
    deg[v] = 0;
    if (i < nv) {
      const float4 pr = g.vprim[v0 + i];
      x[v] = pr.x; w1[v] = pr.y; w2[v] = pr.z; data[v] = pr.w;
      wt[v] = g.vwt[v0 + i];
      deg[v] = g.deg[v0 + i];
      s_bar[i] = g.vbar[v0 + i];
    }
  }
  FB_STAMP(1);
  // Owned edges = source-role incidences of this chunk's vertices. They are dealt to the threads RANK-MAJOR:
  // first every vertex's first owned edge (in vertex order), then every vertex's second one, ... The host
  // orders a chunk by (degree, owned count) descending, so the vertices that have an r-th owned edge are a
  // few contiguous runs and a quarter-warp's eight edges have (almost always) eight consecutive sources:
  // their x_bar reads and their source-record stores fall into eight different 16-byte bank groups (FB_KPAD).
  // Cut edges are wherever their source vertex falls - spread over the chunk by the same sort.
  //   per-vertex counts + histogram -> exclusive scan (warp 0) -> vertex-major list -> rank-major list
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    const int i = tid + v * NT;
    if (i < nv) {
      int cnt = 0;
      for (int k = 0; k < deg[v]; ++k) cnt += (g.inc_nbr[base + s_koff[k] + i] & 1) ^ 1;
      s_scan[i] = cnt;
      atomicAdd(&s_hist[cnt], 1);
    }
  }
  __syncthreads();
  if (tid < 32) {
    int run = 0;
    for (int b = 0; b < nv; b += 32) {
      const int i = b + tid;
      const int d0 = (i < nv) ? s_scan[i] : 0;
      int inc = d0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (tid >= o) inc += n;
      }
      if (i < nv) s_scan[i] = run + inc - d0;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (tid == 0) {
      s_scan[nv] = run;
      *s_cnt = run;
      // s_roff[r] = #owned edges of rank < r = sum over r' < r of #vertices with more than r' owned edges
      int more = 0;
      for (int r = g.max_deg; r >= 0; --r) {  // s_hist[r] := #vertices with > r owned edges
        const int h = s_hist[r];
        s_hist[r] = more;
        more += h;
      }
      int at = 0;
      for (int r = 0; r <= g.max_deg; ++r) {
        s_roff[r] = at;
        at += s_hist[r];
      }
    }
  }
  __syncthreads();
  const int ne = *s_cnt;  // <= NT * SPT (planner)
  FB_STAMP(2);
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    const int i = tid + v * NT;
    if (i < nv) {
      int at = s_scan[i];
      for (int k = 0; k < deg[v]; ++k)
        if ((g.inc_nbr[base + s_koff[k] + i] & 1) == 0) s_list[at++] = ((unsigned int)i << 16) | (unsigned int)k;
    }
  }
  __syncthreads();
  FB_STAMP(3);
  for (int r = tid >> 5; r <= g.max_deg && s_hist[r] > 0; r += NT >> 5) {  // one warp per rank
    int run = s_roff[r];
    const unsigned int lt = (1u << (tid & 31)) - 1u;
    for (int b = 0; b < nv; b += 32) {
      const int i = b + (tid & 31);
      const int first = (i < nv) ? s_scan[i] : 0;
      const bool act = (i < nv) && (s_scan[i + 1] - first > r);
      const unsigned int m = __ballot_sync(0xffffffffu, act);
      if (act) s_rank[run + __popc(m & lt)] = s_list[first + r];
      run += __popc(m);
    }
  }
  __syncthreads();
  FB_STAMP(4);
  // Balanced dealing. A warp's 16-byte shared-memory access costs max(4, most addresses in one 16-byte bank
  // group) wavefronts. Sources: position p of the rank-major list keeps the source GROUP (vertex index mod 8)
  // it has - consecutive vertices, so every warp-sized block holds each group 4 times. Targets: position p
  // WANTS target group (lane + 2 * (lane / 8) + block parity) mod 8 - each group 4 times per block, and over
  // two blocks every source group asks for all eight target groups. Edges of a source group are matched to the
  // positions of that group that want their target group (any order); what does not match - surplus edges and
  // the cut edges, whose targets are in another SM - fills the positions left over. Only the assignment of
  // edges to threads changes, neither the arithmetic nor the accumulation order (that is by record slot).
  for (int pp = tid; pp < ne; pp += NT) {
    const unsigned int rec = s_rank[pp];
    const int i = (int)(rec >> 16), k = (int)(rec & 0xFFFFu);
    const int nbr = g.inc_nbr[base + s_koff[k] + i] >> 1;
    const int ag = i & 7;
    const int tg = (nbr / Vc == c) ? ((nbr - c * Vc) & 7) : 8;
    const int lane = pp & 31;
    const int want = (lane + 2 * (lane >> 3) + ((pp >> 5) & 1)) & 7;
    const int m = atomicAdd(&s_sup[ag * 9 + tg], 1);
    const int q = atomicAdd(&s_dem[ag * 8 + want], 1);
    s_tmp[pp] = (tg << 28) | (m << 14) | q;
  }
  __syncthreads();
  if (tid < 8) {
    int hb = 0, ob = 0;
    for (int d = 0; d < 8; ++d) {
      const int dem = s_dem[tid * 8 + d], sup = s_sup[tid * 9 + d];
      s_hb[tid * 9 + d] = hb;
      s_ob[tid * 9 + d] = ob;
      hb += max(0, dem - sup);
      ob += max(0, sup - dem);
    }
    s_hb[tid * 9 + 8] = hb;
    s_ob[tid * 9 + 8] = ob;
  } else if (tid == 32) {
    int at = 0;
    for (int kk = 0; kk < 64; ++kk) {
      s_off[kk] = at;
      at += s_dem[kk];
    }
  }
  __syncthreads();
  for (int pp = tid; pp < ne; pp += NT) {
    const int ag = (int)(s_rank[pp] >> 16) & 7;
    const int lane = pp & 31;
    const int want = (lane + 2 * (lane >> 3) + ((pp >> 5) & 1)) & 7;
    s_pos[s_off[ag * 8 + want] + (s_tmp[pp] & 0x3FFF)] = (unsigned short)pp;
  }
  __syncthreads();
  for (int pp = tid; pp < ne; pp += NT) {
    const unsigned int rec = s_rank[pp];
    const int ag = (int)(rec >> 16) & 7;
    const int tg = (int)((unsigned int)s_tmp[pp] >> 28), m = (s_tmp[pp] >> 14) & 0x3FFF;
    const int dem = tg < 8 ? s_dem[ag * 8 + tg] : 0;
    int at;
    if (m < dem) {
      at = s_off[ag * 8 + tg] + m;
    } else {
      const int o = s_ob[ag * 9 + tg] + (m - dem);  // index among this source group's unmatched edges
      int d = 0;
      while (d < 7 && s_hb[ag * 9 + d + 1] <= o) ++d;  // the list whose holes contain hole number o
      at = s_off[ag * 8 + d] + s_sup[ag * 9 + d] + (o - s_hb[ag * 9 + d]);
    }
    s_list[s_pos[at]] = rec;
  }
  __syncthreads();
  FB_STAMP(5);
  float q1[SPT], q2[SPT], q3[SPT];
  uint32_t src[SPT];      // (local index of the source vertex << 16) | shared-memory record slot of the source
  uint32_t dst_bar[SPT];  // address of the target's x_bar record: shared::cta, or shared::cluster | 1 if remote
  uint32_t dst_p4[SPT];   // address of the target endpoint's p4 record (same address space as dst_bar)
  // staged so that the dependent global loads of all SPT edges overlap: (q, s, edge id) -> target slot -> rest
  int e_id[SPT];
#pragma unroll
  for (int u = 0; u < SPT; ++u) {
    const int j = tid + u * NT;
    q1[u] = q2[u] = q3[u] = 0.f;
    src[u] = 0;
    dst_bar[u] = bar_addr;
    dst_p4[u] = p4_addr;
    e_id[u] = 0;
    if (j < ne) {
      const unsigned int rec = s_list[j];
      const int i = (int)(rec >> 16), k = (int)(rec & 0xFFFFu);
      const int slot = s_koff[k] + i;  // global slot (relative to the chunk's base)
      src[u] = (rec & 0xFFFF0000u) | (unsigned int)(s_kpad[k] + i);
      const float4 q = g.inc_q[base + slot];
      q1[u] = q.x; q2[u] = q.y; q3[u] = q.z;
      dst_bar[u] = (uint32_t)(__float_as_int(q.w) >> 1);  // the target vertex, until the last stage
      e_id[u] = g.inc_edge[base + slot];
      s_s[j] = g.inc_s[base + slot];
    }
  }
#pragma unroll
  for (int u = 0; u < SPT; ++u)
    if (tid + u * NT < ne) e_id[u] = g.edge_slot[2 * e_id[u] + 1];  // global slot of the target's incidence
#pragma unroll
  for (int u = 0; u < SPT; ++u) {
    const int j = tid + u * NT;
    if (j < ne) {
      const int nbr = (int)dst_bar[u];
      const int oc = nbr / Vc;  // the target vertex lives in chunk oc, and so do its slots
      // bit 0 of dst_bar = "target lives in another CTA": only those go through the DSMEM path
      // (ld/st.shared::cluster is several times slower than ld/st.shared even for the own CTA)
      const bool remote = (oc != c);
      const int ti = nbr - oc * Vc;
      const uint32_t tb = bar_addr + (uint32_t)ti * 16u;
      // the target's record: its global slot is koff_oc[k2] + ti -> shared-memory slot kpad_oc[k2] + ti
      const int tl = e_id[u] - g.chunk_base[oc] - ti;  // = koff_oc[k2]
      int tpad = 0;
      if (!remote) {
        int k2 = 0;
        while (s_koff[k2] != tl) ++k2;
        tpad = s_kpad[k2];
      } else {
        const int* row = koff_row(g, oc);
        int prev = 0;
        for (int k2 = 0; prev != tl; ++k2) {
          const int nxt = row[k2 + 1];
          tpad += kpad_round(nxt - prev);
          prev = nxt;
        }
      }
      const uint32_t tp = p4_addr + (uint32_t)(tpad + ti) * 16u;
      dst_bar[u] = remote ? (mapa(tb, (uint32_t)oc) | 1u) : tb;
      dst_p4[u] = remote ? mapa(tp, (uint32_t)oc) : tp;
    }
  }
  FB_STAMP(6);
  __syncthreads();  // the load-phase lists (aliases of s_p4) are dead from here on
  for (int slot = tid; slot < s_kpad[g.max_deg]; slot += NT)
    s_p1[slot] = neg0;  // target slots keep -0.0f forever; source slots are rewritten every iteration
  cluster.sync();

  FB_STAMP(7);
  const flame_b200_pd_params p = a.p;

  // ---- phase A: one dual update per edge, contributions pushed to both endpoints ----
  // Rounds of PD_MLP edges: all operand loads of a round are issued before its arithmetic so the
  // shared-memory / DSMEM latencies overlap.
  auto phase_a = [&]() {
#pragma unroll
    for (int u0 = 0; u0 < SPT; u0 += PD_MLP) {
      float4 s[PD_MLP], bi[PD_MLP], bj[PD_MLP];
#pragma unroll
      for (int v = 0; v < PD_MLP; ++v) {
        const int u = u0 + v;
        if (u < SPT && tid + u * NT < ne) {
          s[v] = s_s[tid + u * NT];
          bi[v] = s_bar[src[u] >> 16];
          bj[v] = (dst_bar[u] & 1u) ? ld_cluster_f4(dst_bar[u] & ~1u) : ld_shared_f4(dst_bar[u]);
        }
      }
#pragma unroll
      for (int v = 0; v < PD_MLP; ++v) {
        const int u = u0 + v;
        if (u < SPT && tid + u * NT < ne) {
          const float4 qu = dual_update(make_float4(q1[u], q2[u], q3[u], 0.f), s[v], bi[v], bj[v], p.step_q);
          q1[u] = qu.x; q2[u] = qu.y; q3[u] = qu.z;
          // primalStep contributions (cc:125-141), rounded as the reference: a = (q1*step_x)*alpha ...
          const float pa = mul(mul(qu.x, p.step_x), s[v].x);
          const float b2 = mul(mul(qu.y, p.step_x), s[v].y);
          const float b3 = mul(mul(qu.z, p.step_x), s[v].y);
          const uint32_t sl = src[u] & 0xFFFFu;
          // source i: x -= a; w1 += a*dx; w2 += a*dy; w1 -= b2; w2 -= b3
          s_p4[sl] = make_float4(-pa, mul(pa, s[v].z), mul(pa, s[v].w), -b2);
          s_p1[sl] = -b3;
          // target j: x += a; w1 += b2; w2 += b3   (u1 = u2 = -0.0f: exact no-ops)
          if (dst_bar[u] & 1u) st_cluster_f4(dst_p4[u], make_float4(pa, b2, b3, neg0));
          else st_shared_f4(dst_p4[u], make_float4(pa, b2, b3, neg0));
        }
      }
    }
  };

  // ---- phase B: ordered accumulation + prox + extragradient ----
  // Shared-memory byte addresses of the vertex's first-incidence records: incidence k is koff[k] records on.
  uint32_t vp4[VPT], vp1[VPT];
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    vp4[v] = p4_addr + (uint32_t)(tid + v * NT) * 16u;
    vp1[v] = smem_u32(s_p1) + (uint32_t)(tid + v * NT) * 4u;
  }
  const uint32_t koff_addr = smem_u32(s_kpad);
  auto phase_b = [&](const bool last) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = tid + v * NT;
      if (i < nv) {
        const float x_old = x[v], w1_old = w1[v], w2_old = w2[v];
        float xx = x_old, ww1 = w1_old, ww2 = w2_old;
        const int d = deg[v];
        // plain loop, two incidences per trip: the 24-32 resident warps hide the shared-memory latency, and
        // a hand-rolled software pipeline costs more issue slots (register copies) than it saves
        int k = 0;
        for (; k + 1 < d; k += 2) {
          uint32_t ko0, ko1;
          asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(ko0), "=r"(ko1) : "r"(koff_addr + (uint32_t)k * 4u));
          const float4 a4 = ld_shared_f4(vp4[v] + ko0 * 16u);
          float a1, b1;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a1) : "r"(vp1[v] + ko0 * 4u));
          const float4 b4 = ld_shared_f4(vp4[v] + ko1 * 16u);
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(b1) : "r"(vp1[v] + ko1 * 4u));
          xx = add(xx, a4.x);
          ww1 = add(ww1, a4.y);
          ww2 = add(ww2, a4.z);
          ww1 = add(ww1, a4.w);
          ww2 = add(ww2, a1);
          xx = add(xx, b4.x);
          ww1 = add(ww1, b4.y);
          ww2 = add(ww2, b4.z);
          ww1 = add(ww1, b4.w);
          ww2 = add(ww2, b1);
        }
        if (k < d) {
          const uint32_t ko0 = (uint32_t)s_kpad[k];
          const float4 a4 = ld_shared_f4(vp4[v] + ko0 * 16u);
          float a1;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a1) : "r"(vp1[v] + ko0 * 4u));
          xx = add(xx, a4.x);
          ww1 = add(ww1, a4.y);
          ww2 = add(ww2, a4.z);
          ww1 = add(ww1, a4.w);
          ww2 = add(ww2, a1);
        }
        const VertexOut o = vertex_finish(xx, ww1, ww2, data[v], wt[v], x_old, w1_old, w2_old, p);
        x[v] = o.prim.x; w1[v] = ww1; w2[v] = ww2;
        s_bar[i] = o.bar;
        if (last) g.vprev[v0 + i] = make_float4(x_old, w1_old, w2_old, 0.0f);
      }
    }
  };

  // Measured alternatives (all bit-exact, all slower on B200): splitting the barriers into
  // arrive / wait around the work that needs no other CTA (interior vertices, local edges) - the extra
  // __syncthreads and the second pass over the vertex registers cost more than the hidden barrier latency.
  for (int it = 0; it < a.n_iters; ++it) {
    phase_a();
    cluster_arrive();  // release: this CTA's contributions, local and remote, are visible
    cluster_wait();
    phase_b(it == a.n_iters - 1);
    cluster_arrive();  // x_bar published (nobody reads it between the two barriers: one buffer suffices)
    cluster_wait();
  }

  FB_STAMP(8);
  // ---- write back (both private copies of every edge's q) ----
#pragma unroll
  for (int v = 0; v < VPT; ++v) {
    const int i = tid + v * NT;
    if (i < nv) {
      g.vbar[v0 + i] = s_bar[i];
      g.vprim[v0 + i] = make_float4(x[v], w1[v], w2[v], data[v]);
    }
  }
  // (staged like the load: edge ids -> target slots -> stores; the neighbour word of inc_q stays as it is)
  int wb_s[SPT], wb_t[SPT];
#pragma unroll
  for (int u = 0; u < SPT; ++u) {
    wb_s[u] = wb_t[u] = 0;
    if (tid + u * NT < ne) {
      const int ps = (int)(src[u] & 0xFFFFu) - (int)(src[u] >> 16);  // = kpad[k] of the source incidence
      int k = 0;
      while (s_kpad[k] != ps) ++k;
      wb_s[u] = base + s_koff[k] + (int)(src[u] >> 16);
      wb_t[u] = g.inc_edge[wb_s[u]];
    }
  }
#pragma unroll
  for (int u = 0; u < SPT; ++u)
    if (tid + u * NT < ne) wb_t[u] = g.edge_slot[2 * wb_t[u] + 1];
#pragma unroll
  for (int u = 0; u < SPT; ++u) {
    if (tid + u * NT < ne) {
      float* qs = reinterpret_cast<float*>(g.inc_q + wb_s[u]);
      float* qt = reinterpret_cast<float*>(g.inc_q + wb_t[u]);
      *reinterpret_cast<float2*>(qs) = make_float2(q1[u], q2[u]);
      qs[2] = q3[u];
      *reinterpret_cast<float2*>(qt) = make_float2(q1[u], q2[u]);
      qt[2] = q3[u];
    }
  }
  FB_STAMP(9);
}

size_t pd_resident_smem(int Vc, int maxInc, int maxEdges, int max_deg) {
  return (size_t)Vc * 16 + (size_t)resident_p4_records(maxInc, maxEdges, Vc, max_deg) * 16 + (size_t)maxEdges * 16 +
         (size_t)resident_inc_records(maxInc, max_deg) * 4 + (size_t)(2 * (max_deg + 2) + 2) * 4 + 128;
}

// The register budgets the kernel is built for: (most threads, owned edges per thread, vertices per thread).
struct ResidentVariant { int threads, spt, vpt; };
static const ResidentVariant kVariants[FB_RESIDENT_VARIANTS] = {{640, 4, 1}, {640, 6, 2}, {1024, 4, 2}};

// Smallest variant that holds chunks of Vc vertices / maxEdges owned edges; -1 = none. *threads = the block
// size to launch: the fewest warps that give every thread vpt vertices and spt edges - a thread count that
// leaves the last vertex round or the last edge round half empty makes the full warps wait at the cluster
// barrier for nothing (config D, 1 112-vertex chunks: 576 threads x (2 vertices, 6 edges) 0.80 ms per wave,
// 768 x (2, 5) 0.86 ms, 384 x (3, 9) 0.84 ms).
int pd_resident_variant(int Vc, int maxEdges, int* threads) {
  int forced = -1, forced_threads = 0;
  if (const char* sv = getenv("FLAME_B200_VARIANT")) forced = atoi(sv);
  if (const char* sv = getenv("FLAME_B200_THREADS")) forced_threads = atoi(sv);
  for (int i = 0; i < FB_RESIDENT_VARIANTS; ++i) {
    if (forced >= 0 && i != forced) continue;
    const ResidentVariant& r = kVariants[i];
    if (Vc <= r.threads * r.vpt && maxEdges <= r.threads * r.spt) {
      int t = std::max((Vc + r.vpt - 1) / r.vpt, (maxEdges + r.spt - 1) / r.spt);
      t = std::max(64, (t + 31) / 32 * 32);
      if (forced_threads >= t && forced_threads <= r.threads && forced_threads % 32 == 0) t = forced_threads;
      if (threads) *threads = std::min(t, r.threads);
      return i;
    }
  }
  return -1;
}

typedef void (*ResidentKernel)(ResidentArgs);
static ResidentKernel resident_kernel(int variant) {
  switch (variant) {
    case 0: return k_pd_resident<640, 4, 1, 2>;
    case 1: return k_pd_resident<640, 6, 2, 1>;
    default: return k_pd_resident<1024, 4, 2, 1>;
  }
}

// cudaFuncSetAttribute applies per device: done once for every device a context is created on.
int pd_resident_init() {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  FB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev >= 0 && dev < 64 && done[dev]) return 0;
  for (int i = 0; i < FB_RESIDENT_VARIANTS; ++i) {
    FB_CUDA(cudaFuncSetAttribute(resident_kernel(i), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    FB_CUDA(cudaFuncSetAttribute(resident_kernel(i), cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  }
  if (dev >= 0 && dev < 64) done[dev] = true;
  return 0;
}

static void resident_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, int n, const ResidentPlan& plan,
                            cudaStream_t st) {
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(n * plan.cluster);
  cfg->blockDim = dim3(plan.threads);
  cfg->dynamicSmemBytes = plan.smem_bytes;
  cfg->stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}

int pd_run_resident(cudaStream_t st, const GraphDev* gs_dev, int n, const ResidentPlan& plan,
                    const flame_b200_pd_params& p, int n_iters) {
  if (int rc = pd_resident_init()) return rc;
  ResidentArgs a;
  a.gs = gs_dev;
  a.p = p;
  a.n_iters = n_iters;
  a.Vc = plan.Vc;
  a.maxInc = plan.maxInc;
  a.maxEdges = plan.maxEdges;
  a.max_deg = plan.max_deg;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  resident_config(&cfg, attr, n, plan, st);
  FB_CUDA(cudaLaunchKernelEx(&cfg, resident_kernel(plan.variant), a));
  count_launch();
  return 0;
}

// How many frame clusters of this plan the device keeps resident at once (one wave of pd_run_resident).
int pd_resident_capacity(const ResidentPlan& plan, int* n_clusters) {
  if (int rc = pd_resident_init()) return rc;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  resident_config(&cfg, attr, 64, plan, nullptr);
  FB_CUDA(cudaOccupancyMaxActiveClusters(n_clusters, resident_kernel(plan.variant), &cfg));
  return 0;
}

// ------------------------------------------------------------------------------------------
// Energies: smoothnessCost (cc:51-71), dataCost (cc:73-85)
// ------------------------------------------------------------------------------------------
// terms layout (floats, caller order): [0, 2E) smoothness pairs (t1, t2) per edge, [2E, 2E+V) data.
__global__ void k_cost_terms(GraphDev g, float* __restrict__ terms) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  const int c = v / g.Vc, i = v - c * g.Vc;
  const int base = g.chunk_base[c];
  const int* koff = koff_row(g, c);
  const float4 pi = g.vprim[v];
  const int d = g.deg[v];
  for (int k = 0; k < d; ++k) {
    const int slot = base + koff[k] + i;
    const int ent = __float_as_int(g.inc_q[slot].w);
    if (ent & 1) continue;  // the source endpoint evaluates the edge
    const float4 s = g.inc_s[slot];
    const float4 pj = g.vprim[ent >> 1];
    // cc:64-67: alpha*|x_i - x_j - w1_i*dx - w2_i*dy|;  beta*|w1_i-w1_j| + beta*|w2_i-w2_j|
    float a = sub(sub(sub(pi.x, pj.x), mul(pi.y, s.z)), mul(pi.z, s.w));
    a = (a >= 0) ? a : -a;
    float b1 = sub(pi.y, pj.y);
    b1 = (b1 >= 0) ? b1 : -b1;
    float b2 = sub(pi.z, pj.z);
    b2 = (b2 >= 0) ? b2 : -b2;
    const int o = g.inc_edge[slot];
    terms[2 * (size_t)o] = mul(s.x, a);
    terms[2 * (size_t)o + 1] = add(mul(s.y, b1), mul(s.y, b2));
  }
  float dd = mul(sub(pi.x, pi.w), g.vwt[v]);  // cc:79-80
  dd = (dd > 0) ? dd : -dd;
  terms[2 * (size_t)g.E + g.vperm[v]] = dd;
}

// Sequential fp32 accumulation in the reference's order (the block stages, thread 0 adds).
__global__ void k_cost_sum_exact(const float* __restrict__ terms, int nE2, int nV, float data_factor,
                                 double* out) {
  __shared__ float buf[1024];
  float acc[2] = {0.0f, 0.0f};
  const float* base[2] = {terms, terms + nE2};
  const int len[2] = {nE2, nV};
  for (int part = 0; part < 2; ++part) {
    float cost = 0.0f;
    for (int off = 0; off < len[part]; off += 1024) {
      const int n = min(1024, len[part] - off);
      for (int i = threadIdx.x; i < n; i += blockDim.x) buf[i] = base[part][off + i];
      __syncthreads();
      if (threadIdx.x == 0)
        for (int i = 0; i < n; ++i) cost = add(cost, buf[i]);
      __syncthreads();
    }
    acc[part] = cost;
  }
  if (threadIdx.x == 0) {
    out[0] = (double)mul(data_factor, acc[0]);  // cc:70
    out[1] = (double)acc[1];
  }
}

__global__ void k_cost_sum_fast(const float* __restrict__ terms, int nE2, int nV, double* out) {
  double s0 = 0.0, s1 = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nE2 + nV; i += gridDim.x * blockDim.x) {
    if (i < nE2) s0 += (double)terms[i];
    else s1 += (double)terms[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out[0], s0);
    atomicAdd(&out[1], s1);
  }
}

int pd_cost(cudaStream_t st, const GraphDev& g, const flame_b200_pd_params& p, int exact_order,
            float* terms, double* out_dev) {
  if (g.V == 0) {
    FB_CUDA(cudaMemsetAsync(out_dev, 0, 2 * sizeof(double), st));
    return 0;
  }
  k_cost_terms<<<nblk(g.V, 256), 256, 0, st>>>(g, terms);
  count_launch();
  if (exact_order) {
    k_cost_sum_exact<<<1, 256, 0, st>>>(terms, 2 * g.E, g.V, p.data_factor, out_dev);
    count_launch();
  } else {
    FB_CUDA(cudaMemsetAsync(out_dev, 0, 2 * sizeof(double), st));
    k_cost_sum_fast<<<std::min(nblk(2 * g.E + g.V, 256), 296), 256, 0, st>>>(terms, 2 * g.E, g.V,
                                                                            out_dev);
    count_launch();
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace fb
