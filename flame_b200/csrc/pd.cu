// NLTGV2-L1 primal-dual (Chambolle-Pock) solver kernels for sm_100a.
//
// Replaces flame::optimizers::nltgv2_l1_graph_regularizer::step and friends
// (src/flame/optimizers/nltgv2_l1_graph_regularizer.{h,cc}). See DESIGN.md for the layout and
// the roofline model. Arithmetic reproduces the reference's scalar fp32 evaluation order
// bit-for-bit: every multiply/add below is an explicitly rounded __fmul_rn/__fadd_rn (the
// reference's x86-64 Release build has no FMA), and the per-vertex gather walks the
// incident edges in ascending ORIGINAL edge id, which is the order the reference's sequential
// edge scatter (cc:118-142) applies its rounded += / -= to each vertex.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

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
// q/1 is exactly q, so the quotient is this clamp (NaN propagates identically); no divide.
__device__ __forceinline__ float prox_conj(float q) {
  return q > 1.0f ? 1.0f : (q < -1.0f ? -1.0f : q);
}

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
// pack / unpack between the caller-order SoA staging image and the internal records
// ------------------------------------------------------------------------------------------
__global__ void k_pack_vertices(GraphDev g, const float* __restrict__ raw) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  const int o = g.vperm[v];
  const int V = g.V;
  const float* b = raw + 2 * (size_t)V;
  g.vprim[v] = make_float4(b[0 * V + o], b[1 * V + o], b[2 * V + o], b[9 * V + o]);
  g.vbar[v] = make_float4(b[3 * V + o], b[4 * V + o], b[5 * V + o], 0.0f);
  g.vprev[v] = make_float4(b[6 * V + o], b[7 * V + o], b[8 * V + o], 0.0f);
  g.vwt[v] = b[10 * V + o];
}

__global__ void k_pack_edges(GraphDev g, const float* __restrict__ raw, int alpha_on_device) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.E) return;
  const int o = g.eperm[e];
  const int V = g.V, E = g.E;
  const int2 sd = g.ends[e];
  const int os = g.vperm[sd.x], od = g.vperm[sd.y];
  const float dx = sub(raw[2 * os], raw[2 * od]);          // pos_i.x - pos_j.x (cc:100)
  const float dy = sub(raw[2 * os + 1], raw[2 * od + 1]);
  const float* b = raw + 13 * (size_t)V;
  float alpha;
  if (alpha_on_device) {
    // flame.cc:2089-2092, 2102: alpha = 1 / sqrt(dx*dx + dy*dy)
    const float len = __fsqrt_rn(add(mul(dx, dx), mul(dy, dy)));
    alpha = __fdiv_rn(1.0f, len);
  } else {
    alpha = b[0 * (size_t)E + o];
  }
  EdgeRec r;
  r.q = make_float4(b[2 * (size_t)E + o], b[3 * (size_t)E + o], b[4 * (size_t)E + o], 0.0f);
  r.s = make_float4(alpha, b[1 * (size_t)E + o], dx, dy);
  g.erec[e] = r;
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

__global__ void k_unpack_edges(GraphDev g, float* __restrict__ raw) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.E) return;
  const int o = g.eperm[e];
  const size_t E = g.E;
  float* b = raw + 13 * (size_t)g.V;
  const EdgeRec r = g.erec[e];
  b[0 * E + o] = r.s.x; b[1 * E + o] = r.s.y;
  b[2 * E + o] = r.q.x; b[3 * E + o] = r.q.y; b[4 * E + o] = r.q.z;
}

__global__ void k_set_data(GraphDev g, const float* __restrict__ data, const float* __restrict__ wt) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= g.V) return;
  const int o = g.vperm[v];
  if (data) g.vprim[v].w = data[o];
  if (wt) g.vwt[v] = wt[o];
}

static inline int nblk(int n, int t) { return (n + t - 1) / t; }

// Fresh-graph rule of Flame::syncGraph (flame.cc:2039-2048): x = x_bar = x_prev = data_term,
// w = w_bar = w_prev = 0, q = 0. blockIdx.y selects the graph of a batch.
__global__ void k_graph_reset(const GraphDev* __restrict__ gs) {
  const GraphDev g = gs[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max(g.V, g.E); i += gridDim.x * blockDim.x) {
    if (i < g.V) {
      const float d = g.vprim[i].w;
      g.vprim[i] = make_float4(d, 0.0f, 0.0f, d);
      g.vbar[i] = make_float4(d, 0.0f, 0.0f, 0.0f);
      g.vprev[i] = make_float4(d, 0.0f, 0.0f, 0.0f);
    }
    if (i < g.E) g.erec[i].q = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  }
}

int pd_reset(cudaStream_t st, const GraphDev* gs_dev, int n, int maxV, int maxE) {
  const int m = std::max(maxV, maxE);
  if (m == 0 || n == 0) return 0;
  k_graph_reset<<<dim3(std::min(nblk(m, 256), 1024), n), 256, 0, st>>>(gs_dev);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

int pd_pack(cudaStream_t st, const GraphDev& g, const float* raw_f, int alpha_on_device, int) {
  if (g.V > 0) {
    k_pack_vertices<<<nblk(g.V, 256), 256, 0, st>>>(g, raw_f);
    count_launch();
  }
  if (g.E > 0) {
    k_pack_edges<<<nblk(g.E, 256), 256, 0, st>>>(g, raw_f, alpha_on_device);
    count_launch();
  }
  FB_CUDA(cudaGetLastError());
  return 0;
}

int pd_unpack(cudaStream_t st, const GraphDev& g, float* raw_f) {
  if (g.V > 0) {
    k_unpack_vertices<<<nblk(g.V, 256), 256, 0, st>>>(g, raw_f);
    count_launch();
  }
  if (g.E > 0) {
    k_unpack_edges<<<nblk(g.E, 256), 256, 0, st>>>(g, raw_f);
    count_launch();
  }
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

// ------------------------------------------------------------------------------------------
// STREAM mode: state in HBM / L2, one edge pass + one vertex pass per iteration, any size.
// blockIdx.y selects the graph of a batch.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_dual_stream(const GraphDev* __restrict__ gs, float step_q) {
  const GraphDev g = gs[blockIdx.y];
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < g.E; e += gridDim.x * blockDim.x) {
    const int2 sd = g.ends[e];
    const float4 s = g.erec[e].s;
    float4 q = g.erec[e].q;
    const float4 bi = g.vbar[sd.x];
    const float4 bj = g.vbar[sd.y];
    g.erec[e].q = dual_update(q, s, bi, bj, step_q);
  }
}

__global__ void __launch_bounds__(256)
k_primal_stream(const GraphDev* __restrict__ gs, flame_b200_pd_params p, int write_prev) {
  const GraphDev g = gs[blockIdx.y];
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < g.V; v += gridDim.x * blockDim.x) {
    const float4 pr = g.vprim[v];
    float x = pr.x, w1 = pr.y, w2 = pr.z;
    const int k0 = g.inc_off[v], k1 = g.inc_off[v + 1];
    for (int k = k0; k < k1; ++k) {
      const int ent = g.inc[k];
      const EdgeRec* r = &g.erec[ent >> 1];
      const float4 q = r->q;
      const float4 s = r->s;
      primal_apply(x, w1, w2, q, s, ent & 1, p.step_x);
    }
    const VertexOut o = vertex_finish(x, w1, w2, pr.w, g.vwt[v], pr.x, pr.y, pr.z, p);
    g.vprim[v] = o.prim;
    g.vbar[v] = o.bar;
    if (write_prev) g.vprev[v] = make_float4(pr.x, pr.y, pr.z, 0.0f);
  }
}

// ------------------------------------------------------------------------------------------
// RESIDENT mode: one thread-block cluster per graph; the whole PD state lives in the cluster's
// distributed shared memory for all n_iters iterations. Edge (dual) and vertex (primal +
// extragradient) phases are separated by barrier.cluster; cross-CTA operands are read through
// DSMEM (ld.shared::cluster via mapa).
// ------------------------------------------------------------------------------------------
struct ResidentArgs {
  const GraphDev* gs;
  flame_b200_pd_params p;
  int n_iters;
  int Vc, maxE, maxInc;
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

// Packed location of a vertex / edge inside the cluster: owner CTA in the top bits.
#define LOC_SHIFT 20
#define LOC_MASK ((1u << LOC_SHIFT) - 1u)

__global__ void __launch_bounds__(1024, 1) k_pd_resident(ResidentArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int c = cluster.block_rank();
  const GraphDev g = a.gs[blockIdx.x / C];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int Vc = a.Vc;

  extern __shared__ __align__(128) unsigned char smem[];
  // identical carve in every CTA of the cluster (remote addressing relies on it)
  EdgeRec* s_erec = reinterpret_cast<EdgeRec*>(smem);
  float4* s_vbar = reinterpret_cast<float4*>(s_erec + a.maxE);
  float4* s_vprim = s_vbar + Vc;
  uint2* s_ends = reinterpret_cast<uint2*>(s_vprim + Vc);
  float* s_vwt = reinterpret_cast<float*>(s_ends + a.maxE);
  int* s_incoff = reinterpret_cast<int*>(s_vwt + Vc);
  uint32_t* s_inc = reinterpret_cast<uint32_t*>(s_incoff + Vc + 1);
  int* s_estart = reinterpret_cast<int*>(s_inc + a.maxInc);  // C + 1 entries (<= 17)

  const int v0 = min(c * Vc, g.V), v1 = min(v0 + Vc, g.V), nv = v1 - v0;
  const int e0 = g.esrc_off[v0], e1 = g.esrc_off[v1], ne = e1 - e0;
  if (tid <= C) s_estart[tid] = g.esrc_off[min(tid * Vc, g.V)];
  __syncthreads();

  // ---- load this CTA's partition ----
  for (int i = tid; i < nv; i += nt) {
    s_vbar[i] = g.vbar[v0 + i];
    s_vprim[i] = g.vprim[v0 + i];
    s_vwt[i] = g.vwt[v0 + i];
  }
  const int k0 = g.inc_off[v0];
  for (int i = tid; i <= nv; i += nt) s_incoff[i] = g.inc_off[v0 + i] - k0;
  const int nk = g.inc_off[v1] - k0;
  for (int k = tid; k < nk; k += nt) {
    const int ent = g.inc[k0 + k];
    const int e = ent >> 1;
    int owner = 0;
    while (owner + 1 < C && s_estart[owner + 1] <= e) ++owner;
    s_inc[k] = (uint32_t)(ent & 1) | ((uint32_t)(e - s_estart[owner]) << 1) |
               ((uint32_t)owner << (LOC_SHIFT + 1));
  }
  for (int i = tid; i < ne; i += nt) {
    s_erec[i] = g.erec[e0 + i];
    const int2 sd = g.ends[e0 + i];
    const int cs = sd.x / Vc, cd = sd.y / Vc;
    s_ends[i] = make_uint2(((uint32_t)cs << LOC_SHIFT) | (uint32_t)(sd.x - cs * Vc),
                           ((uint32_t)cd << LOC_SHIFT) | (uint32_t)(sd.y - cd * Vc));
  }
  cluster.sync();

  const uint32_t vbar_base = smem_u32(s_vbar);
  const uint32_t erec_base = smem_u32(s_erec);
  const flame_b200_pd_params p = a.p;

  for (int it = 0; it < a.n_iters; ++it) {
    // ---- dual step: one thread per owned edge ----
    for (int i = tid; i < ne; i += nt) {
      const uint2 sd = s_ends[i];
      const float4 bi = ld_cluster_f4(mapa(vbar_base + (sd.x & LOC_MASK) * 16u, sd.x >> LOC_SHIFT));
      const float4 bj = ld_cluster_f4(mapa(vbar_base + (sd.y & LOC_MASK) * 16u, sd.y >> LOC_SHIFT));
      s_erec[i].q = dual_update(s_erec[i].q, s_erec[i].s, bi, bj, p.step_q);
    }
    cluster.sync();
    // ---- primal step + prox + extragradient: one thread per owned vertex ----
    const bool last = (it == a.n_iters - 1);
    for (int i = tid; i < nv; i += nt) {
      const float4 pr = s_vprim[i];
      float x = pr.x, w1 = pr.y, w2 = pr.z;
      const int ka = s_incoff[i], kb = s_incoff[i + 1];
      for (int k = ka; k < kb; ++k) {
        const uint32_t ent = s_inc[k];
        const uint32_t addr =
            mapa(erec_base + ((ent >> 1) & LOC_MASK) * 32u, ent >> (LOC_SHIFT + 1));
        const float4 q = ld_cluster_f4(addr);
        const float4 s = ld_cluster_f4(addr + 16u);
        primal_apply(x, w1, w2, q, s, ent & 1u, p.step_x);
      }
      const VertexOut o = vertex_finish(x, w1, w2, pr.w, s_vwt[i], pr.x, pr.y, pr.z, p);
      s_vprim[i] = o.prim;
      s_vbar[i] = o.bar;
      if (last) g.vprev[v0 + i] = make_float4(pr.x, pr.y, pr.z, 0.0f);
    }
    cluster.sync();
  }

  // ---- write back ----
  for (int i = tid; i < nv; i += nt) {
    g.vbar[v0 + i] = s_vbar[i];
    g.vprim[v0 + i] = s_vprim[i];
  }
  for (int i = tid; i < ne; i += nt) g.erec[e0 + i].q = s_erec[i].q;
}

size_t pd_resident_smem(int C, int Vc, int maxE, int maxInc) {
  return (size_t)maxE * 32 + (size_t)Vc * 32 + (size_t)maxE * 8 + (size_t)Vc * 4 +
         (size_t)(Vc + 1) * 4 + (size_t)maxInc * 4 + (size_t)(C + 1) * 4 + 64;
}

int pd_resident_init() {
  static bool done = false;
  if (done) return 0;
  FB_CUDA(cudaFuncSetAttribute(k_pd_resident, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               227 * 1024));
  FB_CUDA(cudaFuncSetAttribute(k_pd_resident, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  done = true;
  return 0;
}

// Chooses the cluster size for one graph from its (host) offset tables.
ResidentPlan pd_plan_resident(int V, int E, const int* esrc_off, const int* inc_off) {
  ResidentPlan best{};
  best.cluster = 0;
  int forced = 0;
  if (const char* s = getenv("FLAME_B200_CLUSTER")) forced = atoi(s);
  int threads = 1024;
  if (const char* s = getenv("FLAME_B200_THREADS")) threads = atoi(s);
  if (V >= (1 << LOC_SHIFT) || E >= (1 << LOC_SHIFT)) return best;
  for (int C = 1; C <= 16; C *= 2) {
    if (forced && C != forced) continue;
    const int Vc = (V + C - 1) / C;
    if (Vc == 0) continue;
    int maxE = 0, maxInc = 0;
    for (int c = 0; c < C; ++c) {
      const int a = std::min(c * Vc, V), b = std::min(a + Vc, V);
      maxE = std::max(maxE, esrc_off[b] - esrc_off[a]);
      maxInc = std::max(maxInc, inc_off[b] - inc_off[a]);
    }
    maxE = std::max(maxE, 1);
    maxInc = std::max(maxInc, 1);
    const size_t smem = pd_resident_smem(C, Vc, maxE, maxInc);
    if (smem > 227 * 1024) continue;
    best.cluster = C;
    best.threads = threads;
    best.smem_bytes = smem;
    best.Vc = Vc;
    best.maxE = maxE;
    best.maxInc = maxInc;
    // heuristic: stop growing once every thread has at most ~2 edges
    if (forced || (E + C - 1) / C <= 2 * threads) break;
  }
  return best;
}

int pd_run_resident(cudaStream_t st, const GraphDev* gs_dev, int n, const ResidentPlan& plan,
                    const flame_b200_pd_params& p, int n_iters) {
  if (int rc = pd_resident_init()) return rc;
  ResidentArgs a;
  a.gs = gs_dev;
  a.p = p;
  a.n_iters = n_iters;
  a.Vc = plan.Vc;
  a.maxE = plan.maxE;
  a.maxInc = plan.maxInc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n * plan.cluster);
  cfg.blockDim = dim3(plan.threads);
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = plan.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  FB_CUDA(cudaLaunchKernelEx(&cfg, k_pd_resident, a));
  count_launch();
  return 0;
}

// gs_dev: device array of n GraphDev; maxV/maxE = maxima over the batch.
int pd_run_stream(cudaStream_t st, const GraphDev* gs_dev, int n, int maxV, int maxE,
                       const flame_b200_pd_params& p, int n_iters) {
  if (n_iters <= 0 || n <= 0) return 0;
  const dim3 ge(std::max(1, std::min(nblk(maxE, 256), 65535)), n);
  const dim3 gv(std::max(1, std::min(nblk(maxV, 256), 65535)), n);
  for (int it = 0; it < n_iters; ++it) {
    k_dual_stream<<<ge, 256, 0, st>>>(gs_dev, p.step_q);
    k_primal_stream<<<gv, 256, 0, st>>>(gs_dev, p, it == n_iters - 1);
  }
  count_launch(2 * n_iters);
  FB_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// Energies: smoothnessCost (cc:51-71), dataCost (cc:73-85)
// ------------------------------------------------------------------------------------------
// terms layout (floats, caller order): [0, 2E) smoothness pairs (t1, t2) per edge, [2E, 2E+V) data.
__global__ void k_cost_terms(GraphDev g, float* __restrict__ terms) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < g.E) {
    const int2 sd = g.ends[i];
    const float4 s = g.erec[i].s;
    const float4 pi = g.vprim[sd.x], pj = g.vprim[sd.y];
    // cc:64-67: alpha*|x_i - x_j - w1_i*dx - w2_i*dy|;  beta*|w1_i-w1_j| + beta*|w2_i-w2_j|
    float a = sub(sub(sub(pi.x, pj.x), mul(pi.y, s.z)), mul(pi.z, s.w));
    a = (a >= 0) ? a : -a;
    float b1 = sub(pi.y, pj.y);
    b1 = (b1 >= 0) ? b1 : -b1;
    float b2 = sub(pi.z, pj.z);
    b2 = (b2 >= 0) ? b2 : -b2;
    const int o = g.eperm[i];
    terms[2 * (size_t)o] = mul(s.x, a);
    terms[2 * (size_t)o + 1] = add(mul(s.y, b1), mul(s.y, b2));
  }
  if (i < g.V) {
    const float4 pr = g.vprim[i];
    float d = mul(sub(pr.x, pr.w), g.vwt[i]);  // cc:79-80
    d = (d > 0) ? d : -d;
    terms[2 * (size_t)g.E + g.vperm[i]] = d;
  }
}

// Sequential fp32 accumulation in the reference's order (one warp stages, lane 0 adds).
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
  const int n = std::max(g.V, g.E);
  if (n == 0) {
    FB_CUDA(cudaMemsetAsync(out_dev, 0, 2 * sizeof(double), st));
    return 0;
  }
  k_cost_terms<<<nblk(n, 256), 256, 0, st>>>(g, terms);
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
