// C ABI + host runtime of libflame_b200.so (context, slots, staging, host-side graph prep).
// The compute is in pd.cu / raster.cu / stereo.cu; nothing here falls back to the CPU.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <memory>
#include <mutex>

#include "internal.h"

namespace fb {

static thread_local std::string t_last_error;
void set_error(const std::string& msg) { t_last_error = msg; }
std::atomic<uint64_t> g_kernel_launches{0};

int stereo_set_camera(void** state, int width, int height, const float K[9], const float Kinv[9]);
int stereo_frame_upload(void** state, cudaStream_t st, int fslot, const uint8_t* img_pad,
                        const float* gradx_pad, const float* grady_pad, int pad, const float pose[7]);
int stereo_update(void* state, cudaStream_t st, const flame_b200_track_params* params, int fnew,
                  int curr_pf, flame_b200_feature* feats, int32_t n_feats,
                  flame_b200_track_diag* diag, int32_t* n_updates);
int stereo_frame_upload_raw(void** state, cudaStream_t st, int fslot, const uint8_t* img, int pad,
                            const float pose[7]);
int stereo_frame_download(void* state, cudaStream_t st, int fslot, uint8_t* img_pad, float* gradx_pad,
                          float* grady_pad);
int stereo_features_upload(void* state, cudaStream_t st, const flame_b200_feature* feats, int32_t n_feats);
int stereo_features_track(void* state, cudaStream_t st, const flame_b200_track_params* params, int fnew,
                          int curr_pf);
int stereo_features_download(void* state, cudaStream_t st, flame_b200_feature* feats,
                             flame_b200_track_diag* diag, int32_t* n_updates, int commit);
void stereo_destroy(void* state);
int stereo_graph_project(void* state, cudaStream_t st, const GraphDev& g, float* d_pos,
                         const float q[4], const float t[3], float scale, float* d_out,
                         const flame_b200_project_rules* rules, int fnew, unsigned char* d_marks);
int stereo_detect(void* state, cudaStream_t st, const flame_b200_detect_params* p, int fref, int fprev,
                  const float* curr_xy, int n_curr, float* out_xy, int max_out, int32_t* n_out);

// ------------------------------------------------------------------------------------------
struct Slot {
  GraphDev g{};
  bool loaded = false;
  // one device blob: [raw floats (13V + 5E) | ints: vperm V | deg V | koff n_chunks*(max_deg+1) |
  //                   chunk_base n_chunks+1 | inc_nbr 2E | inc_edge 2E | edges 2E],
  // mirrored by one pinned host blob -> single H2D.
  unsigned char* d_blob = nullptr;
  size_t d_blob_cap = 0;
  unsigned char* h_blob = nullptr;
  size_t h_blob_cap = 0;
  cudaEvent_t blob_free = nullptr;  // recorded after the last async op that reads/writes h_blob
  size_t stage_extra = 0;           // bytes the next graph upload reserves behind its own staging image
  size_t stage_used = 0;            // end of the last graph upload's staging image inside h_blob
  std::vector<int32_t> h_tris, h_edges;  // topology computed by process_frame_submit itself (flame_b200_delaunay)
  int32_t h_nt = -1, h_ne = -1;
  // asynchronous whole-frame pipeline (flame_b200_process_frame_submit / _wait)
  cudaStream_t stream = nullptr;    // the slot's own stream: frames in different slots overlap
  cudaEvent_t done = nullptr;       // recorded after the frame's last device->host copy
  cudaEvent_t uploaded = nullptr;   // recorded on the slot's stream after the upload + layout kernels
  bool in_flight = false;
  bool launched = false;            // the frame's wave has been issued
  float* out_map = nullptr;         // host destination of the dense map of the frame in flight
  int out_V = 0;                    // vertex count of the frame in flight (x | w1 | w2 staged in h_blob)
  size_t out_off = 0;               // where in h_blob
  int capV = 0, capE = 0;
  float* d_pos = nullptr;           // 2V caller order (kept for the dense projection)
  float* d_values = nullptr;        // V caller order
  ResidentPlan plan{};              // chosen at upload (cluster == 0 -> streaming only)
  bool stale = false;               // graph_project moved the vertices: edge records need a resync
  float* d_carry = nullptr;         // graph_resync: unpacked state of the previous graph
  size_t carry_cap = 0;             // floats
  // mesh
  MeshDev m{};
  int capT = 0;
  size_t cap_pix = 0;
  float* own_img = nullptr;         // the slot's own dense map while m.img is redirected (flame_b200_mesh_redirect)
  float* values_dst = nullptr;      // where mesh_project additionally copies the vertex values (peer memory)
  int cap_tiles = 0;                // screen tiles the bin tables are sized for
  size_t cap_bins = 0;              // ints in m.bins
  bool bins_dirty = true;           // topology / positions changed since the tile bins were built
  int mesh_V = -1;                  // vertex count the triangle indices were validated against at bind time
  int* d_tris = nullptr;            // owning pointers behind the const views in `m`
  unsigned char* d_tri_ok = nullptr;
  unsigned char* d_tri_filt = nullptr;  // flame_b200_mesh_filter result (capT bytes)
  int* d_vt = nullptr;                  // flame_b200_mesh_normals scratch: vertex -> triangle lists
  size_t vt_cap = 0;                    // ints
  float* d_mpos = nullptr;          // interpolate_mesh(): caller-supplied pos / values
  float* d_mvalues = nullptr;
  int capMV = 0;
  float* d_raw() const { return reinterpret_cast<float*>(d_blob); }
};

}  // namespace fb

struct flame_b200_ctx {
  int device = 0;
  int n_slots = 0;
  cudaStream_t stream = nullptr;
  std::vector<fb::Slot> slots;
  fb::GraphDev* d_graphs = nullptr;  // n_slots entries, mirrors slots[i].g
  fb::MeshDev* d_meshes = nullptr;   // n_slots entries, mirrors slots[i].m
  double* d_cost = nullptr;
  double* h_cost = nullptr;          // pinned, 2 doubles
  void* stereo = nullptr;
  // asynchronous dense projection (flame_b200_ctx_set_async_projection): the projection kernels run on
  // `side` so that the next frames' solver launches on `stream` overlap them
  bool async_proj = false;
  int prio_lo = 0;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_vals = nullptr;
  struct Pending { int lo, hi; cudaEvent_t done; };  // projections of slots [lo, hi) issued on `side`
  std::vector<Pending> pending;                      // in issue order (the side stream is in-order)
  std::vector<cudaEvent_t> ev_pool;
  // host scratch for graph prep
  std::vector<uint32_t> keys;
  std::vector<int> order, koff_tmp, pre_deg, pre_src, sort_tmp;
  std::vector<uint64_t> degw, degm;  // per vertex (degree | owned edges << 32): caller order, Z-curve order
  int own_cnt[FB_RESIDENT_MAX_DEG + 2];
  int pd_pref = 0;                   // FLAME_B200_PREF_*: how prep_plan sizes the frame clusters
  int busy_sms[17] = {0};            // [C] = C x clusters of C CTAs the device keeps resident (full shared memory)
  // frame pipeline: frames submitted to consecutive slots are solved as one wave (wave_flush)
  std::vector<int> wave;
  flame_b200_pd_params wave_params{};
  int wave_iters = 0;
  float wave_scale = 1.0f;
  cudaStream_t down = nullptr;       // device->host copies of finished waves
  cudaEvent_t ev_wave = nullptr;
  int cap_cluster = 0, cap_variant = -1, cap_frames = 1;  // cached pd_resident_capacity of the last plan
  size_t cap_smem = 0;
  // host-side accounting of the frame pipeline (flame_b200_ctx_stats)
  double t_prep = 0, t_stage = 0, t_enqueue = 0, t_wait = 0, t_tri = 0;
  uint64_t n_frames = 0, h2d_bytes = 0, d2h_bytes = 0;
};

namespace fb {

static inline uint32_t part1by1(uint32_t x) {
  x &= 0x0000ffffu;
  x = (x ^ (x << 8)) & 0x00ff00ffu;
  x = (x ^ (x << 4)) & 0x0f0f0f0fu;
  x = (x ^ (x << 2)) & 0x33333333u;
  x = (x ^ (x << 1)) & 0x55555555u;
  return x;
}

struct PrepOut {
  int Vc = 0, n_chunks = 0, max_deg = 0;
  ResidentPlan plan{};
};

// uploaded ints: [vperm V | koff n_chunks*(max_deg+1) | chunk_base n_chunks+1 | edges 2E]
static size_t ints_uploaded(int V, int E, int n_chunks, int max_deg) {
  return (size_t)V + (size_t)n_chunks * (max_deg + 1) + (size_t)n_chunks + 1 + (size_t)2 * E;
}
// device-built ints behind them: [vinv V | deg V | inc_nbr 2E | inc_edge 2E | edge_slot 2E],
// then the resync maps [vtx_from V | edge_from E]
static size_t blob_bytes(int V, int E, int n_chunks, int max_deg) {
  return (raw_f_total(V, E) + ints_uploaded(V, E, n_chunks, max_deg) + (size_t)2 * V + (size_t)6 * E +
          (size_t)V + E) * 4 + 256;
}

// Host side of the graph preparation: only the VERTEX order and the launch plan - O(V) work plus one
// pass over the edge list for validation and degrees. (The launch configuration of the resident solver -
// cluster size, shared-memory carve - has to be known on the host, which is why this part is not a kernel.)
// The O(E) incidence layout is built on the device (pd.cu: k_layout_*).
//  1. Morton (Z-curve) order of the vertices -> spatially compact chunks of Vc vertices;
//  2. cluster size C / chunk width Vc from the shared-memory footprint (0 -> streaming chunks).
static int prep_plan(flame_b200_ctx* c, int V, int E, const float* pos, const int32_t* edges,
                     PrepOut* out) {
  // per vertex: degree in the low half, owned (source-role) edges in the high half of one 64-bit word
  c->degw.assign(V, 0);
  uint64_t* const dg = c->degw.data();
  for (int e = 0; e < E; ++e) {
    const int s = edges[2 * e], d = edges[2 * e + 1];
    if ((unsigned)s >= (unsigned)V || (unsigned)d >= (unsigned)V) {
      set_error("graph_upload: edge endpoint out of range");
      return FLAME_B200_EINVAL;
    }
    dg[s] += 0x100000001ull;
    dg[d] += 1;
  }
  int max_deg = 0;
  for (int v = 0; v < V; ++v) max_deg = std::max(max_deg, (int)(uint32_t)dg[v]);

  // 1. Z-curve order at cell granularity: a 64 x 64 grid over the bounding box, one stable counting sort by
  //    the Morton code of the cell (vertices of a cell keep their input order)
  float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (int v = 0; v < V; ++v) {
    const float x = pos[2 * v], y = pos[2 * v + 1];
    xmin = x < xmin ? x : xmin;
    xmax = x > xmax ? x : xmax;
    ymin = y < ymin ? y : ymin;
    ymax = y > ymax ? y : ymax;
  }
  const float span = std::max(std::max(xmax - xmin, ymax - ymin), 1e-6f);
  const float sc = 63.999f / span;
  static uint16_t spread[64];  // bits of a 6-bit coordinate moved to the even positions
  if (!spread[63])
    for (int i = 63; i >= 0; --i) spread[i] = (uint16_t)part1by1((uint32_t)i);
  c->keys.resize(V);
  int cnt[4097];
  std::memset(cnt, 0, sizeof(cnt));
  for (int v = 0; v < V; ++v) {
    const float fx = (pos[2 * v] - xmin) * sc, fy = (pos[2 * v + 1] - ymin) * sc;
    const int qx = (fx >= 0.0f && fx < 64.0f) ? (int)fx : 0;  // NaN-safe
    const int qy = (fy >= 0.0f && fy < 64.0f) ? (int)fy : 0;
    const uint32_t key = (uint32_t)spread[qx] | ((uint32_t)spread[qy] << 1);
    c->keys[v] = key;
    cnt[key + 1]++;
  }
  for (int b = 0; b < 4096; ++b) cnt[b + 1] += cnt[b];
  c->order.resize(V);
  for (int v = 0; v < V; ++v) c->order[cnt[c->keys[v]]++] = v;
  const int* morton = c->order.data();  // morton[i] = caller vertex at Z-curve position i

  // 2. plan: per-chunk incidence / owned-edge counts from prefix sums over the Morton order. The packed counts
  //    are gathered into Z-curve order once (c->degm, reused by prep_fill): everything after is sequential.
  c->degm.resize(V);
  c->pre_deg.resize(V + 1);
  c->pre_src.resize(V + 1);
  c->pre_deg[0] = c->pre_src[0] = 0;
  {
    uint64_t* const dm = c->degm.data();
    int a_deg = 0, a_src = 0;
    for (int i = 0; i < V; ++i) {
      const uint64_t w = dg[morton[i]];
      dm[i] = w;
      a_deg += (int)(uint32_t)w;
      a_src += (int)(w >> 32);
      c->pre_deg[i + 1] = a_deg;
      c->pre_src[i + 1] = a_src;
    }
  }
  int forced = 0;
  if (const char* sv = getenv("FLAME_B200_CLUSTER")) forced = atoi(sv);
  ResidentPlan plan{};
  int best_busy = -1;
  if (V > 0 && V < (1 << 20) && max_deg <= FB_RESIDENT_MAX_DEG) {
    // throughput (default): the smallest cluster that holds the graph - any size, not only powers of two -
    // so that as many frames as possible are resident at once; latency: the smallest cluster of the
    // 640-thread build (a frame spread over more SMs finishes sooner), falling back to the former
    for (int pass = (c->pd_pref == FLAME_B200_PREF_LATENCY ? 0 : 1); pass < 2 && !plan.cluster; ++pass)
      for (int C = 1; C <= 16; ++C) {
        if (forced && C != forced) continue;
        const int Vc = (V + C - 1) / C;
        int maxInc = 1, maxEdges = 1;
        for (int k = 0; k < C; ++k) {
          const int a = std::min(k * Vc, V), b = std::min((k + 1) * Vc, V);
          maxInc = std::max(maxInc, c->pre_deg[b] - c->pre_deg[a]);
          maxEdges = std::max(maxEdges, c->pre_src[b] - c->pre_src[a]);
        }
        const size_t smem = pd_resident_smem(Vc, maxInc, maxEdges, max_deg);
        if (smem > 227 * 1024) continue;
        int thr = 0;
        const int variant = pd_resident_variant(Vc, maxEdges, &thr);
        if (variant < 0) continue;
        if (pass == 0 && variant != 0) continue;
        // throughput: the cluster size that keeps the most SMs busy (clusters resident on the device x CTAs per
        // cluster - a cluster may not straddle GPCs, so this is not monotonic: 15 x 9 = 135 SMs beats 15 x 8
        // and 12 x 10); ties and the latency pass: the first (smallest) that fits
        const int busy = (pass == 1 && !forced) ? c->busy_sms[C] : 0;
        if (busy < best_busy || (plan.cluster && busy == best_busy)) continue;
        best_busy = busy;
        plan.cluster = C;
        plan.variant = variant;
        plan.Vc = Vc;
        plan.maxInc = maxInc;
        plan.maxEdges = maxEdges;
        plan.max_deg = max_deg;
        plan.smem_bytes = smem;
        plan.threads = thr;
        if (pass == 0 || forced) break;
      }
  }
  const int Vc = plan.cluster ? plan.Vc : std::min(std::max(V, 1), 1024);
  const int n_chunks = V > 0 ? (V + Vc - 1) / Vc : 0;
  out->Vc = Vc;
  out->n_chunks = n_chunks;
  out->max_deg = max_deg;
  out->plan = plan;
  return 0;
}

//  3. stable sort by descending degree inside each chunk (uniform warps) -> vperm; the slot tables
//     koff / chunk_base: incidence k of the i-th vertex of chunk c lives at chunk_base[c] + koff[c][k] + i.
static void prep_fill(flame_b200_ctx* c, int V, int E, const int32_t* edges, const PrepOut& pr,
                      int* ints, int** p_vperm, int** p_koff, int** p_chunk_base, int** p_edges) {
  const int Vc = pr.Vc, n_chunks = pr.n_chunks, max_deg = pr.max_deg;
  const int* morton = c->order.data();
  int* vperm = ints;
  int* koff = vperm + V;
  int* chunk_base = koff + (size_t)n_chunks * (max_deg + 1);
  int* edges_out = chunk_base + n_chunks + 1;
  *p_vperm = vperm; *p_koff = koff; *p_chunk_base = chunk_base; *p_edges = edges_out;
  if (E > 0) std::memcpy(edges_out, edges, (size_t)2 * E * sizeof(int));
  c->koff_tmp.assign(2 * (size_t)(max_deg + 2), 0);
  int* cnt = c->koff_tmp.data();         // cnt[d]   = #vertices of degree d in the chunk
  int* start = cnt + (max_deg + 2);      // start[d] = #vertices of degree > d (block start of d)
  const uint64_t* const dm = c->degm.data();  // (degree | owned << 32) at every Z-curve position (prep_plan)
  chunk_base[0] = 0;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int a = ch * Vc, b = std::min(a + Vc, V);
    std::fill(cnt, cnt + max_deg + 2, 0);
    for (int i = a; i < b; ++i) cnt[(uint32_t)dm[i]]++;
    int run = 0;
    for (int d = max_deg; d >= 0; --d) {
      start[d] = run;
      run += cnt[d];
    }
    // incidence k of a vertex exists iff deg > k; those vertices are the first start[k] of the chunk
    int* ko = koff + (size_t)ch * (max_deg + 1);
    int tot = 0;
    for (int k = 0; k <= max_deg; ++k) {
      ko[k] = tot;
      tot += start[k];
    }
    chunk_base[ch + 1] = chunk_base[ch] + ko[max_deg];
    for (int d = 0; d <= max_deg; ++d) cnt[d] = start[d];  // write cursors
    if (!pr.plan.cluster) {
      for (int i = a; i < b; ++i) vperm[a + cnt[(uint32_t)dm[i]]++] = morton[i];
      continue;
    }
    // resident solver: secondary key = number of owned (source-role) edges, descending, so that the vertices
    // that have an r-th owned edge form a few contiguous runs (pd.cu deals the edges rank-major). Two stable
    // counting sorts: by owned count first (into positions of the chunk), by degree second.
    c->sort_tmp.resize(b - a);
    int* own_cnt = c->own_cnt;
    std::fill(own_cnt, own_cnt + max_deg + 2, 0);
    for (int i = a; i < b; ++i) own_cnt[dm[i] >> 32]++;
    for (int d = max_deg, run2 = 0; d >= 0; --d) {
      const int n = own_cnt[d];
      own_cnt[d] = run2;
      run2 += n;
    }
    for (int i = a; i < b; ++i) c->sort_tmp[own_cnt[dm[i] >> 32]++] = i;
    for (int j = 0; j < b - a; ++j) {
      const int i = c->sort_tmp[j];
      vperm[a + cnt[(uint32_t)dm[i]]++] = morton[i];
    }
  }
}

static int check_slot(flame_b200_ctx* c, int slot, int n = 1) {
  if (!c) {
    set_error("null context");
    return FLAME_B200_EINVAL;
  }
  if (slot < 0 || n < 1 || slot + n > c->n_slots) {
    set_error("slot range out of bounds");
    return FLAME_B200_EINVAL;
  }
  return 0;
}

static int use_device(flame_b200_ctx* c) {
  FB_CUDA(cudaSetDevice(c->device));
  return 0;
}

static int wait_blob(Slot& s) {
  if (s.blob_free) FB_CUDA(cudaEventSynchronize(s.blob_free));
  return 0;
}

// Grow-only buffers. A pointer is nulled and its capacity zeroed right after the free, and the new
// capacity is committed only once the allocation succeeded: a failed cudaMalloc leaves the slot empty
// but consistent (no dangling pointer for a later call, no double free in ctx_destroy).
template <typename T>
static int regrow(T** p, size_t bytes) {
  if (*p) {
    cudaError_t e = cudaFree(*p);
    *p = nullptr;
    if (e != cudaSuccess) {
      set_error(std::string("cudaFree: ") + cudaGetErrorString(e));
      return FLAME_B200_ECUDA;
    }
  }
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes);
  if (e != cudaSuccess) {
    *p = nullptr;
    set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? FLAME_B200_ENOMEM : FLAME_B200_ECUDA;
  }
  return 0;
}

static int ensure_blob(Slot& s, size_t need) {
  if (need > s.d_blob_cap) {
    const size_t n = std::max(need, s.d_blob_cap + s.d_blob_cap / 2);
    s.d_blob_cap = 0;
    if (int rc = regrow(&s.d_blob, n)) return rc;
    s.d_blob_cap = n;
  }
  if (need > s.h_blob_cap) {
    const size_t n = std::max(need, s.h_blob_cap + s.h_blob_cap / 2);
    s.h_blob_cap = 0;
    if (s.h_blob) FB_CUDA(cudaFreeHost(s.h_blob));
    s.h_blob = nullptr;
    FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s.h_blob), n));
    s.h_blob_cap = n;
  }
  if (!s.blob_free) FB_CUDA(cudaEventCreateWithFlags(&s.blob_free, cudaEventDisableTiming));
  return 0;
}

static int ensure_graph(Slot& s, int V, int E) {
  if (V > s.capV) {
    const int n = std::max(V, s.capV + s.capV / 2);
    s.capV = 0;
    s.loaded = false;
    if (int rc = regrow(&s.g.vprim, (size_t)n * sizeof(float4))) return rc;
    if (int rc = regrow(&s.g.vbar, (size_t)n * sizeof(float4))) return rc;
    if (int rc = regrow(&s.g.vbar2, (size_t)n * sizeof(float4))) return rc;
    if (int rc = regrow(&s.g.vprev, (size_t)n * sizeof(float4))) return rc;
    if (int rc = regrow(&s.g.vwt, (size_t)n * sizeof(float))) return rc;
    if (int rc = regrow(&s.d_pos, (size_t)n * 2 * sizeof(float))) return rc;
    if (int rc = regrow(&s.d_values, (size_t)n * sizeof(float))) return rc;
    s.capV = n;
  }
  if (E > s.capE) {
    const int n = std::max(E, s.capE + s.capE / 2);
    s.capE = 0;
    s.loaded = false;
    if (int rc = regrow(&s.g.inc_q, (size_t)2 * n * sizeof(float4))) return rc;
    if (int rc = regrow(&s.g.inc_s, (size_t)2 * n * sizeof(float4))) return rc;
    s.capE = n;
  }
  return 0;
}

static void fill_or_copy(float* dst, const float* src, size_t n, float fill) {
  if (src) std::memcpy(dst, src, n * sizeof(float));
  else std::fill(dst, dst + n, fill);
}

// The device-side descriptor tables are written by a one-thread kernel whose argument IS the
// descriptor: captured at launch, so there is no pinned staging copy a later call could overwrite
// while an earlier copy is still queued.
template <typename T>
__global__ void k_publish(T* dst, T v) { *dst = v; }

static int publish_mesh(flame_b200_ctx* c, int slot) {
  k_publish<MeshDev><<<1, 1, 0, c->stream>>>(c->d_meshes + slot, c->slots[slot].m);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

static int publish_graph(flame_b200_ctx* c, int slot) {
  k_publish<GraphDev><<<1, 1, 0, c->stream>>>(c->d_graphs + slot, c->slots[slot].g);
  count_launch();
  FB_CUDA(cudaGetLastError());
  return 0;
}

// Orders the pending projections that touch slots [lo, hi) - and, the projection stream being in-order,
// everything issued before them - ahead of whatever is issued next on the main stream (device-side
// wait, the host does not block).
static int join_range(flame_b200_ctx* c, int lo, int hi) {
  int last = -1;
  for (int i = (int)c->pending.size() - 1; i >= 0; --i)
    if (c->pending[i].lo < hi && lo < c->pending[i].hi) {
      last = i;
      break;
    }
  if (last < 0) return 0;
  FB_CUDA(cudaStreamWaitEvent(c->stream, c->pending[last].done, 0));
  for (int i = 0; i <= last; ++i) c->ev_pool.push_back(c->pending[i].done);
  c->pending.erase(c->pending.begin(), c->pending.begin() + last + 1);
  return 0;
}
static int join_side(flame_b200_ctx* c) { return join_range(c, 0, c->n_slots); }

}  // namespace fb

using namespace fb;

extern "C" {

const char* flame_b200_last_error(void) { return t_last_error.c_str(); }
int flame_b200_version(void) { return FLAME_B200_VERSION; }
uint64_t flame_b200_kernel_launches(void) { return g_kernel_launches.load(); }

void flame_b200_pd_default_params(flame_b200_pd_params* p) {
  // nltgv2_l1_graph_regularizer.h:121-129
  p->data_factor = 0.1f;
  p->step_x = 0.001f;
  p->step_q = 125.0f;
  p->theta = 0.25f;
  p->x_min = 0.0f;
  p->x_max = 10.0f;
}

void flame_b200_track_default_params(flame_b200_track_params* p) {
  // inverse_depth_filter.h:48-68, line_stereo.h:47-59, inverse_depth_meas_model.h:44-52, params.h:58-67
  p->fparams.win_size = 5;
  p->fparams.search_sigma = 2.0f;
  p->fparams.min_grad_mag = 5.0f;
  p->fparams.idepth_min = 1e-3f;
  p->fparams.idepth_max = 2.0f;
  p->fparams.epilength_min = 3.0f;
  p->fparams.epilength_max = 32.0f;
  p->fparams.process_var_factor = 1.01f;
  p->fparams.process_fail_var_factor = 1.1f;
  p->fparams.sparams.max_cost = 1300.0f;
  p->fparams.sparams.do_subpixel = 1;
  p->fparams.sparams.sample_dist = 1.0f;
  p->fparams.sparams.second_best_factor = 1.5f;
  p->zparams.win_size = 5;
  p->zparams.pixel_var = 16.0f;
  p->zparams.epipolar_line_var = 1.0f;
  p->min_baseline = 0.01f;
  p->rescale_factor_min = 0.7f;
  p->rescale_factor_max = 1.4f;
  p->idepth_var_max = 0.5f * 0.5f;
  p->max_dropouts = 5;
  p->outlier_sigma_thresh = 3.0f;
  p->do_meas_fusion = 1;
  p->do_letterbox = 0;
}

int flame_b200_ctx_create(int device, int n_slots, flame_b200_ctx** out) {
  if (!out || n_slots < 1) {
    set_error("ctx_create: bad arguments");
    return FLAME_B200_EINVAL;
  }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) +
              " (libflame_b200 has no CPU fallback)");
    return FLAME_B200_ENODEV;
  }
  if (device < 0 || device >= ndev) {
    set_error("ctx_create: device ordinal out of range");
    return FLAME_B200_EINVAL;
  }
  cudaDeviceProp prop;
  FB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error(std::string("device ") + prop.name + " is not sm_100: kernels are built for sm_100a only");
    return FLAME_B200_ENODEV;
  }
  FB_CUDA(cudaSetDevice(device));
  std::unique_ptr<flame_b200_ctx> c(new flame_b200_ctx());
  c->device = device;
  c->n_slots = n_slots;
  c->slots.resize(n_slots);
  // the main stream carries the solver, whose clusters need whole SMs of one GPC free at the same time: it
  // gets the highest priority so that pending cluster launches are served before the many small CTAs of the
  // projection (second stream, lowest priority), which then fill the SMs the clusters leave idle
  int prio_lo = 0, prio_hi = 0;
  FB_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  c->prio_lo = prio_lo;
  FB_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
  FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_graphs), sizeof(GraphDev) * n_slots));
  FB_CUDA(cudaMemset(c->d_graphs, 0, sizeof(GraphDev) * n_slots));
  FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_meshes), sizeof(MeshDev) * n_slots));
  FB_CUDA(cudaMemset(c->d_meshes, 0, sizeof(MeshDev) * n_slots));
  FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->d_cost), 2 * sizeof(double)));
  FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&c->h_cost), 2 * sizeof(double)));
  if (int rc = pd_resident_init()) return rc;
  for (int C = 1; C <= 16; ++C) {
    ResidentPlan pl{};
    pl.cluster = C;
    pl.variant = 1;  // one CTA per SM for every build at these shared-memory sizes
    pl.threads = 640;
    pl.smem_bytes = 200 * 1024;
    int n = 0;
    if (int rc = pd_resident_capacity(pl, &n)) return rc;
    c->busy_sms[C] = n * C;
  }
  *out = c.release();
  return 0;
}

void flame_b200_ctx_destroy(flame_b200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->side) cudaStreamSynchronize(c->side);
  cudaStreamSynchronize(c->stream);
  for (auto& sl : c->slots)
    if (sl.stream) cudaStreamSynchronize(sl.stream);
  if (c->down) {
    cudaStreamSynchronize(c->down);
    cudaStreamDestroy(c->down);
  }
  if (c->ev_wave) cudaEventDestroy(c->ev_wave);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_vals) cudaEventDestroy(c->ev_vals);
  for (auto& pnd : c->pending) cudaEventDestroy(pnd.done);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  for (auto& s : c->slots) {
    if (s.own_img) s.m.img = s.own_img;  // never free a redirected (foreign) destination
    for (void* p : {(void*)s.g.vprim, (void*)s.g.vbar, (void*)s.g.vbar2, (void*)s.g.vprev,
                    (void*)s.g.vwt, (void*)s.g.inc_q, (void*)s.g.inc_s, (void*)s.d_blob,
                    (void*)s.d_pos, (void*)s.d_values,
                    (void*)s.d_tris, (void*)s.d_tri_ok, (void*)s.d_tri_filt, (void*)s.m.img, (void*)s.m.stat,
                    (void*)s.m.bins, (void*)s.m.wide, (void*)s.m.counters, (void*)s.d_mpos, (void*)s.d_mvalues,
                    (void*)s.d_carry, (void*)s.d_vt})
      if (p) cudaFree(p);
    if (s.h_blob) cudaFreeHost(s.h_blob);
    if (s.blob_free) cudaEventDestroy(s.blob_free);
    if (s.done) cudaEventDestroy(s.done);
    if (s.uploaded) cudaEventDestroy(s.uploaded);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  stereo_destroy(c->stereo);
  cudaFree(c->d_graphs);
  cudaFree(c->d_meshes);
  cudaFree(c->d_cost);
  cudaFreeHost(c->h_cost);
  cudaStreamDestroy(c->stream);
  delete c;
}

static int wave_flush(flame_b200_ctx* c);  // frame pipeline, below

int flame_b200_ctx_sync(flame_b200_ctx* c) {
  if (!c) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  if (int rc = wave_flush(c)) return rc;
  if (int rc = join_side(c)) return rc;
  for (auto& sl : c->slots)
    if (sl.stream) FB_CUDA(cudaStreamSynchronize(sl.stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  if (c->down) FB_CUDA(cudaStreamSynchronize(c->down));
  return 0;
}

int flame_b200_ctx_set_async_projection(flame_b200_ctx* c, int on) {
  if (!c) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  if (int rc = join_side(c)) return rc;
  if (on && !c->side) {
    FB_CUDA(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, c->prio_lo));
    FB_CUDA(cudaEventCreateWithFlags(&c->ev_vals, cudaEventDisableTiming));
  }
  c->async_proj = on != 0;
  return 0;
}

int flame_b200_ctx_join(flame_b200_ctx* c) {
  if (!c) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return join_side(c);
}

void* flame_b200_ctx_stream(flame_b200_ctx* c) { return c ? (void*)c->stream : nullptr; }

int flame_b200_ctx_set_pd_preference(flame_b200_ctx* c, int pref) {
  if (!c || (pref != FLAME_B200_PREF_THROUGHPUT && pref != FLAME_B200_PREF_LATENCY)) {
    set_error("ctx_set_pd_preference: bad arguments");
    return FLAME_B200_EINVAL;
  }
  c->pd_pref = pref;
  return 0;
}

int flame_b200_host_alloc(size_t bytes, void** out) {
  if (!out) return FLAME_B200_EINVAL;
  FB_CUDA(cudaMallocHost(out, bytes));
  return 0;
}
int flame_b200_host_free(void* p) {
  if (p) FB_CUDA(cudaFreeHost(p));
  return 0;
}

}  // extern "C"

// fresh != 0: g carries only pos / edges / data_term / data_weight (flame_b200_graph_init): the
// fresh-vertex defaults are generated on the device and only 4V floats + the index tables cross PCIe.
struct SyncMaps {
  const int32_t* vtx_from;
  const int32_t* edge_from;
  const float* x_init;
  int sticky;
};

// sync != nullptr (flame_b200_graph_resync): like fresh, plus the state of the slot's current graph is
// unpacked into d_carry first and gathered into the new raw image on the device.
static int graph_upload_impl(flame_b200_ctx* c, int slot, const flame_b200_graph_in* g, int fresh,
                             const SyncMaps* sync = nullptr) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!g || g->V < 0 || g->E < 0 || (g->V > 0 && (!g->pos || !g->data_term || !g->x)) ||
      (g->E > 0 && !g->edges)) {
    set_error("graph_upload: missing required array");
    return FLAME_B200_EINVAL;
  }
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  const int V = g->V, E = g->E;
  PrepOut prep;
  timespec tp0, tp1;
  clock_gettime(CLOCK_MONOTONIC, &tp0);
  if (int rc = prep_plan(c, V, E, g->pos, g->edges, &prep)) return rc;
  clock_gettime(CLOCK_MONOTONIC, &tp1);
  c->t_prep += (double)(tp1.tv_sec - tp0.tv_sec) + 1e-9 * (double)(tp1.tv_nsec - tp0.tv_nsec);
  if (int rc = wait_blob(s)) return rc;
  const int Vold = s.g.V, Eold = s.g.E;
  if (sync) {  // before any buffer of the old graph is touched
    const size_t need = raw_f_total(Vold, Eold) + 1;
    if (need > s.carry_cap) {
      if (s.d_carry) FB_CUDA(cudaFree(s.d_carry));
      s.d_carry = nullptr;
      const size_t n = std::max(need, s.carry_cap + s.carry_cap / 2);
      FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.d_carry), n * sizeof(float)));
      s.carry_cap = n;
    }
    if (int rc = pd_unpack(c->stream, s.g, s.d_carry)) return rc;
  }
  const size_t graph_stage = (blob_bytes(V, E, prep.n_chunks, prep.max_deg) + 255) & ~(size_t)255;
  if (int rc = ensure_blob(s, graph_stage + s.stage_extra)) return rc;
  s.stage_used = graph_stage;
  if (int rc = ensure_graph(s, V, E)) return rc;

  float* hf = reinterpret_cast<float*>(s.h_blob);
  int* hi = reinterpret_cast<int*>(hf + raw_f_total(V, E));
  int *h_vperm, *h_koff, *h_chunk_base, *h_edges;
  clock_gettime(CLOCK_MONOTONIC, &tp0);
  prep_fill(c, V, E, g->edges, prep, hi, &h_vperm, &h_koff, &h_chunk_base, &h_edges);
  clock_gettime(CLOCK_MONOTONIC, &tp1);
  c->t_prep += (double)(tp1.tv_sec - tp0.tv_sec) + 1e-9 * (double)(tp1.tv_nsec - tp0.tv_nsec);
  s.plan = prep.plan;

  if (V > 0) std::memcpy(hf + raw_f_offset(RF_POS, V, E), g->pos, (size_t)2 * V * sizeof(float));
  if (!fresh) {
    fill_or_copy(hf + raw_f_offset(RF_X, V, E), g->x, V, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_W1, V, E), g->w1, V, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_W2, V, E), g->w2, V, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_XB, V, E), g->x_bar ? g->x_bar : g->x, V, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_W1B, V, E), g->w1_bar, V, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_W2B, V, E), g->w2_bar, V, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_XP, V, E), g->x_prev ? g->x_prev : g->x, V, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_W1P, V, E), g->w1_prev, V, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_W2P, V, E), g->w2_prev, V, 0.0f);
  }
  if (sync && sync->x_init) fill_or_copy(hf + raw_f_offset(RF_X, V, E), sync->x_init, V, 0.0f);
  fill_or_copy(hf + raw_f_offset(RF_DATA, V, E), g->data_term, V, 0.0f);
  fill_or_copy(hf + raw_f_offset(RF_WT, V, E), g->data_weight, V, 1.0f);
  if (!fresh) {
    fill_or_copy(hf + raw_f_offset(RF_ALPHA, V, E), g->alpha, E, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_BETA, V, E), g->beta, E, 1.0f);
    fill_or_copy(hf + raw_f_offset(RF_Q1, V, E), g->q1, E, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_Q2, V, E), g->q2, E, 0.0f);
    fill_or_copy(hf + raw_f_offset(RF_Q3, V, E), g->q3, E, 0.0f);
  }

  const size_t n_up = ints_uploaded(V, E, prep.n_chunks, prep.max_deg);
  const size_t bytes = (raw_f_total(V, E) + n_up) * 4;
  // device-built tables, then the resync maps [vtx_from V | edge_from E]
  const size_t o_dev = bytes;
  const size_t o_maps = o_dev + ((size_t)2 * V + (size_t)6 * E) * 4, maps_bytes = sync ? ((size_t)V + E) * 4 : 0;
  if (sync) {
    int* hm = reinterpret_cast<int*>(s.h_blob + o_maps);
    if (V > 0) std::memcpy(hm, sync->vtx_from, (size_t)V * 4);
    if (E > 0) std::memcpy(hm + V, sync->edge_from, (size_t)E * 4);
  }
  if (o_maps + maps_bytes > s.d_blob_cap || o_maps + maps_bytes > s.h_blob_cap) {
    set_error("internal: staging blob too small");
    return FLAME_B200_ENOMEM;
  }
  if (!fresh) {
    FB_CUDA(cudaMemcpyAsync(s.d_blob, s.h_blob, bytes, cudaMemcpyHostToDevice, c->stream));
  } else {  // pos | data_term, data_weight | index tables
    const size_t o_data = raw_f_offset(RF_DATA, V, E) * 4, o_int = raw_f_total(V, E) * 4;
    if (V > 0) {
      FB_CUDA(cudaMemcpyAsync(s.d_blob, s.h_blob, (size_t)2 * V * 4, cudaMemcpyHostToDevice, c->stream));
      FB_CUDA(cudaMemcpyAsync(s.d_blob + o_data, s.h_blob + o_data, (size_t)2 * V * 4,
                              cudaMemcpyHostToDevice, c->stream));
    }
    FB_CUDA(cudaMemcpyAsync(s.d_blob + o_int, s.h_blob + o_int, bytes - o_int, cudaMemcpyHostToDevice,
                            c->stream));
    if (sync && sync->x_init && V > 0) {
      const size_t o_x = raw_f_offset(RF_X, V, E) * 4;
      FB_CUDA(cudaMemcpyAsync(s.d_blob + o_x, s.h_blob + o_x, (size_t)V * 4, cudaMemcpyHostToDevice,
                              c->stream));
    }
  }
  if (maps_bytes)
    FB_CUDA(cudaMemcpyAsync(s.d_blob + o_maps, s.h_blob + o_maps, maps_bytes, cudaMemcpyHostToDevice,
                            c->stream));
  FB_CUDA(cudaEventRecord(s.blob_free, c->stream));
  int* di = reinterpret_cast<int*>(s.d_raw() + raw_f_total(V, E));
  s.g.V = V;
  s.g.E = E;
  s.g.Vc = prep.Vc;
  s.g.n_chunks = prep.n_chunks;
  s.g.max_deg = prep.max_deg;
  s.g.vperm = di + (h_vperm - hi);
  s.g.koff = di + (h_koff - hi);
  s.g.chunk_base = di + (h_chunk_base - hi);
  s.g.edges = di + (h_edges - hi);
  int* dd = reinterpret_cast<int*>(s.d_blob + o_dev);
  s.g.vinv = dd;
  s.g.deg = dd + V;
  s.g.inc_nbr = s.g.deg + V;
  s.g.inc_edge = s.g.inc_nbr + 2 * (size_t)E;
  s.g.edge_slot = s.g.inc_edge + 2 * (size_t)E;
  if (V > 0)
    FB_CUDA(cudaMemcpyAsync(s.d_pos, s.d_raw(), (size_t)2 * V * sizeof(float),
                            cudaMemcpyDeviceToDevice, c->stream));
  if (sync) {
    const int* dm = reinterpret_cast<const int*>(s.d_blob + o_maps);
    if (int rc = pd_resync_gather(c->stream, s.d_raw(), V, E, s.d_carry, Vold, Eold, dm, dm + V,
                                  sync->x_init != nullptr, sync->sticky)) return rc;
    fresh = 0;
  }
  if (int rc = pd_pack(c->stream, s.g, s.d_raw(), g->alpha == nullptr, fresh)) return rc;
  s.loaded = true;
  s.stale = false;
  s.bins_dirty = true;  // the dense projection bins triangles by vertex position
  return publish_graph(c, slot);
}

extern "C" {

int flame_b200_graph_upload(flame_b200_ctx* c, int slot, const flame_b200_graph_in* g) {
  return graph_upload_impl(c, slot, g, 0);
}

int flame_b200_graph_init(flame_b200_ctx* c, int slot, int32_t V, const float* pos,
                          const float* idepth, const float* weight, int32_t E,
                          const int32_t* edges) {
  flame_b200_graph_in g;
  std::memset(&g, 0, sizeof(g));
  g.V = V;
  g.E = E;
  g.pos = pos;
  g.edges = edges;
  g.data_term = idepth;
  g.data_weight = weight;
  g.x = idepth;  // x = x_bar = x_prev = idepth (flame.cc:2045-2048)
  return graph_upload_impl(c, slot, &g, 1);
}

static int graph_project_impl(flame_b200_ctx* c, int slot, const float q_wxyz[4], const float t[3],
                              float graph_scale, const flame_b200_project_rules* rules, int fnew, float* pos_out,
                              float* idepth_out, uint8_t* remove_out) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!q_wxyz || !t || !(graph_scale > 0.0f)) {
    set_error("graph_project: bad arguments");
    return FLAME_B200_EINVAL;
  }
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded) {
    set_error("graph_project: slot has no graph");
    return FLAME_B200_ESTATE;
  }
  const int V = s.g.V;
  if (int rc = wait_blob(s)) return rc;
  // the raw staging area (>= 13V floats) is scratch between uploads: [u_new 2V | idepth_new V | marks V bytes]
  unsigned char* d_marks = reinterpret_cast<unsigned char*>(s.d_raw() + 3 * (size_t)V);
  if (int rc = stereo_graph_project(c->stereo, c->stream, s.g, s.d_pos, q_wxyz, t, graph_scale, s.d_raw(), rules,
                                    fnew, d_marks)) return rc;
  s.stale = true;
  s.bins_dirty = true;
  if (V == 0) return 0;
  const size_t bytes = (size_t)3 * V * 4 + (rules ? (size_t)V : 0);
  FB_CUDA(cudaMemcpyAsync(s.h_blob, s.d_blob, bytes, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaEventRecord(s.blob_free, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  const float* hf = reinterpret_cast<const float*>(s.h_blob);
  if (pos_out) std::memcpy(pos_out, hf, (size_t)2 * V * 4);
  if (idepth_out) std::memcpy(idepth_out, hf + 2 * (size_t)V, (size_t)V * 4);
  if (remove_out && rules) std::memcpy(remove_out, s.h_blob + (size_t)3 * V * 4, (size_t)V);
  return 0;
}

int flame_b200_graph_project(flame_b200_ctx* c, int slot, const float q_wxyz[4], const float t[3],
                             float graph_scale, float* pos_out, float* idepth_out) {
  return graph_project_impl(c, slot, q_wxyz, t, graph_scale, nullptr, -1, pos_out, idepth_out, nullptr);
}

int flame_b200_graph_project_marked(flame_b200_ctx* c, int slot, const float q_wxyz[4], const float t[3],
                                    float graph_scale, const flame_b200_project_rules* rules, int fnew,
                                    float* pos_out, float* idepth_out, uint8_t* remove_out) {
  if (!rules) {
    set_error("graph_project_marked: rules missing");
    return FLAME_B200_EINVAL;
  }
  return graph_project_impl(c, slot, q_wxyz, t, graph_scale, rules, fnew, pos_out, idepth_out, remove_out);
}

int flame_b200_detect_features(flame_b200_ctx* c, const flame_b200_detect_params* params, int fref, int fprev,
                               const float* curr_xy, int32_t n_curr, float* features_xy, int32_t max_features,
                               int32_t* n_features) {
  if (!c || !params || (n_curr > 0 && !curr_xy) || !n_features) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return stereo_detect(c->stereo, c->stream, params, fref, fprev, curr_xy, n_curr, features_xy, max_features,
                       n_features);
}

int flame_b200_graph_resync(flame_b200_ctx* c, int slot, const flame_b200_graph_sync* y) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!y || y->V < 0 || y->E < 0 || (y->V > 0 && (!y->pos || !y->data_term || !y->vtx_from)) ||
      (y->E > 0 && (!y->edges || !y->edge_from))) {
    set_error("graph_resync: missing required array");
    return FLAME_B200_EINVAL;
  }
  Slot& s = c->slots[slot];
  if (!s.loaded) {
    set_error("graph_resync: slot has no graph to carry state from (use flame_b200_graph_init)");
    return FLAME_B200_ESTATE;
  }
  for (int v = 0; v < y->V; ++v)
    if (y->vtx_from[v] < -1 || y->vtx_from[v] >= s.g.V) {
      set_error("graph_resync: vtx_from out of range");
      return FLAME_B200_EINVAL;
    }
  for (int e = 0; e < y->E; ++e)
    if (y->edge_from[e] < -1 || y->edge_from[e] >= s.g.E) {
      set_error("graph_resync: edge_from out of range");
      return FLAME_B200_EINVAL;
    }
  flame_b200_graph_in g;
  std::memset(&g, 0, sizeof(g));
  g.V = y->V;
  g.E = y->E;
  g.pos = y->pos;
  g.edges = y->edges;
  g.data_term = y->data_term;
  g.data_weight = y->data_weight;
  g.x = y->data_term;
  SyncMaps m{y->vtx_from, y->edge_from, y->x_init, y->check_sticky_obstacles != 0};
  return graph_upload_impl(c, slot, &g, 1, &m);
}

int flame_b200_graph_set_data(flame_b200_ctx* c, int slot, const float* data_term,
                              const float* data_weight) {
  if (int rc = check_slot(c, slot)) return rc;
  if (int rc = use_device(c)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded) {
    set_error("graph_set_data: slot has no graph");
    return FLAME_B200_ESTATE;
  }
  const int V = s.g.V;
  if (int rc = wait_blob(s)) return rc;
  float* hf = reinterpret_cast<float*>(s.h_blob);
  if (data_term) std::memcpy(hf, data_term, (size_t)V * 4);
  if (data_weight) std::memcpy(hf + V, data_weight, (size_t)V * 4);
  FB_CUDA(cudaMemcpyAsync(s.d_raw(), hf, (size_t)2 * V * 4, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(cudaEventRecord(s.blob_free, c->stream));
  return pd_set_data(c->stream, s.g, data_term ? s.d_raw() : nullptr,
                     data_weight ? s.d_raw() + V : nullptr);
}

int flame_b200_graph_reset(flame_b200_ctx* c, int slot0, int n_slots) {
  if (int rc = check_slot(c, slot0, n_slots)) return rc;
  if (int rc = use_device(c)) return rc;
  int maxV = 0, maxE = 0;
  for (int i = slot0; i < slot0 + n_slots; ++i) {
    if (!c->slots[i].loaded) {
      set_error("graph_reset: slot has no graph");
      return FLAME_B200_ESTATE;
    }
    maxV = std::max(maxV, c->slots[i].g.V);
    maxE = std::max(maxE, c->slots[i].g.E);
  }
  return pd_reset(c->stream, c->d_graphs + slot0, n_slots, maxV, maxE);
}

int flame_b200_graph_download(flame_b200_ctx* c, int slot, const flame_b200_graph_out* out) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!out) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded) {
    set_error("graph_download: slot has no graph");
    return FLAME_B200_ESTATE;
  }
  const int V = s.g.V, E = s.g.E;
  if (int rc = wait_blob(s)) return rc;
  if (int rc = pd_unpack(c->stream, s.g, s.d_raw())) return rc;
  float* vdst[9] = {out->x, out->w1, out->w2, out->x_bar, out->w1_bar, out->w2_bar,
                    out->x_prev, out->w1_prev, out->w2_prev};
  float* edst[5] = {out->alpha, out->beta, out->q1, out->q2, out->q3};
  int vlo = 9, vhi = -1, elo = 5, ehi = -1;
  for (int i = 0; i < 9; ++i)
    if (vdst[i]) { vlo = std::min(vlo, i); vhi = std::max(vhi, i); }
  for (int i = 0; i < 5; ++i)
    if (edst[i]) { elo = std::min(elo, i); ehi = std::max(ehi, i); }
  float* hf = reinterpret_cast<float*>(s.h_blob);
  if (vhi >= 0 && V > 0) {
    const size_t off = raw_f_offset(RF_X + vlo, V, E), n = (size_t)(vhi - vlo + 1) * V;
    FB_CUDA(cudaMemcpyAsync(hf + off, s.d_raw() + off, n * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  if (ehi >= 0 && E > 0) {
    const size_t off = raw_f_offset(RF_ALPHA + elo, V, E), n = (size_t)(ehi - elo + 1) * E;
    FB_CUDA(cudaMemcpyAsync(hf + off, s.d_raw() + off, n * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  FB_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < 9; ++i)
    if (vdst[i] && V > 0) std::memcpy(vdst[i], hf + raw_f_offset(RF_X + i, V, E), (size_t)V * 4);
  for (int i = 0; i < 5; ++i)
    if (edst[i] && E > 0) std::memcpy(edst[i], hf + raw_f_offset(RF_ALPHA + i, V, E), (size_t)E * 4);
  return 0;
}

int flame_b200_pd_run(flame_b200_ctx* c, int slot0, int n_slots, const flame_b200_pd_params* params,
                      int n_iters, int mode) {
  if (int rc = check_slot(c, slot0, n_slots)) return rc;
  if (!params || n_iters < 0 || mode < FLAME_B200_PD_AUTO || mode > FLAME_B200_PD_RESIDENT) {
    set_error("pd_run: bad arguments");
    return FLAME_B200_EINVAL;
  }
  if (int rc = use_device(c)) return rc;
  int maxV = 0, max_chunks = 0, maxVc = 0;
  ResidentPlan plan{};
  bool fits = true;
  for (int i = slot0; i < slot0 + n_slots; ++i) {
    const Slot& s = c->slots[i];
    if (!s.loaded) {
      set_error("pd_run: slot has no graph");
      return FLAME_B200_ESTATE;
    }
    if (s.stale) {
      set_error("pd_run: vertices were moved by graph_project; call graph_resync (or upload) first");
      return FLAME_B200_ESTATE;
    }
    maxV = std::max(maxV, s.g.V);
    max_chunks = std::max(max_chunks, s.g.n_chunks);
    maxVc = std::max(maxVc, s.g.Vc);
    if (s.g.V == 0) continue;
    if (s.plan.cluster == 0) fits = false;
    plan.cluster = std::max(plan.cluster, s.plan.cluster);
    plan.Vc = std::max(plan.Vc, s.plan.Vc);
    plan.maxInc = std::max(plan.maxInc, s.plan.maxInc);
    plan.maxEdges = std::max(plan.maxEdges, s.plan.maxEdges);
    plan.max_deg = std::max(plan.max_deg, s.plan.max_deg);
    plan.threads = std::max(plan.threads, s.plan.threads);
  }
  if (n_iters == 0 || maxV == 0) return 0;
  if (fits) {
    plan.smem_bytes = pd_resident_smem(plan.Vc, plan.maxInc, plan.maxEdges, plan.max_deg);
    plan.variant = pd_resident_variant(plan.Vc, plan.maxEdges, &plan.threads);
    if (plan.smem_bytes > 227 * 1024 || plan.variant < 0) fits = false;
  }
  if (mode == FLAME_B200_PD_RESIDENT && !fits) {
    set_error("pd_run: graph does not fit the resident (cluster shared memory) mode");
    return FLAME_B200_EINVAL;
  }
  if (fits && mode != FLAME_B200_PD_STREAM)
    return pd_run_resident(c->stream, c->d_graphs + slot0, n_slots, plan, *params, n_iters);
  const int threads = (maxVc + 31) / 32 * 32;  // widest chunk; the kernel covers it with 128-thread CTAs
  return pd_run_stream(c->stream, c->d_graphs + slot0, n_slots, max_chunks, threads, *params, n_iters);
}

// Host-only: the vertex order and launch plan a graph upload would compute (no device, no context).
int flame_b200_graph_plan(int32_t V, const float* pos, int32_t E, const int32_t* edges, int preference,
                          const int32_t* busy_sms, flame_b200_graph_layout* out, int32_t* vperm,
                          int32_t koff_capacity, int32_t* koff, int32_t* chunk_base) {
  if (V < 0 || E < 0 || !out || (V > 0 && !pos) || (E > 0 && !edges)) {
    set_error("graph_plan: bad arguments");
    return FLAME_B200_EINVAL;
  }
  std::unique_ptr<flame_b200_ctx> c(new flame_b200_ctx());
  c->pd_pref = preference;
  for (int i = 0; i <= 16; ++i) c->busy_sms[i] = busy_sms ? busy_sms[i] : 0;
  PrepOut prep;
  if (int rc = prep_plan(c.get(), V, E, pos, edges, &prep)) return rc;
  out->chunk_width = prep.Vc;
  out->n_chunks = prep.n_chunks;
  out->max_degree = prep.max_deg;
  out->cluster = prep.plan.cluster;
  out->variant = prep.plan.cluster ? prep.plan.variant : -1;
  out->threads = prep.plan.threads;
  out->max_incidences = prep.plan.maxInc;
  out->max_owned_edges = prep.plan.maxEdges;
  out->shared_bytes = (int64_t)prep.plan.smem_bytes;
  if (!vperm && !koff && !chunk_base) return 0;
  const size_t need = (size_t)prep.n_chunks * (prep.max_deg + 1);
  if (koff && (size_t)koff_capacity < need) {
    set_error("graph_plan: koff capacity too small (needs n_chunks * (max_degree + 1))");
    return FLAME_B200_ENOMEM;
  }
  std::vector<int> ints(ints_uploaded(V, E, prep.n_chunks, prep.max_deg));
  int *h_vperm, *h_koff, *h_chunk_base, *h_edges;
  prep_fill(c.get(), V, E, edges, prep, ints.data(), &h_vperm, &h_koff, &h_chunk_base, &h_edges);
  if (vperm && V > 0) std::memcpy(vperm, h_vperm, (size_t)V * sizeof(int));
  if (koff && need) std::memcpy(koff, h_koff, need * sizeof(int));
  if (chunk_base) std::memcpy(chunk_base, h_chunk_base, ((size_t)prep.n_chunks + 1) * sizeof(int));
  return 0;
}

int flame_b200_pd_wave(flame_b200_ctx* c, int slot, int32_t* frames) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!frames) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  const Slot& s = c->slots[slot];
  if (!s.loaded) {
    set_error("pd_wave: slot has no graph");
    return FLAME_B200_ESTATE;
  }
  *frames = 0;
  if (s.g.V == 0 || s.plan.cluster == 0) return 0;
  ResidentPlan plan = s.plan;
  plan.smem_bytes = pd_resident_smem(plan.Vc, plan.maxInc, plan.maxEdges, plan.max_deg);
  plan.variant = pd_resident_variant(plan.Vc, plan.maxEdges, &plan.threads);
  if (plan.smem_bytes > 227 * 1024 || plan.variant < 0) return 0;
  int n = 0;
  if (int rc = pd_resident_capacity(plan, &n)) return rc;
  *frames = n;
  return 0;
}

int flame_b200_pd_cost(flame_b200_ctx* c, int slot, const flame_b200_pd_params* params,
                       int exact_order, float* smoothness_cost, float* data_cost) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!params) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded) {
    set_error("pd_cost: slot has no graph");
    return FLAME_B200_ESTATE;
  }
  if (int rc = wait_blob(s)) return rc;
  if (int rc = pd_cost(c->stream, s.g, *params, exact_order, s.d_raw(), c->d_cost)) return rc;
  FB_CUDA(cudaMemcpyAsync(c->h_cost, c->d_cost, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  if (smoothness_cost)
    *smoothness_cost = exact_order ? (float)c->h_cost[0] : params->data_factor * (float)c->h_cost[0];
  if (data_cost) *data_cost = (float)c->h_cost[1];
  return 0;
}

// ---- dense projection ---------------------------------------------------------------------
static inline size_t mesh_stage_bytes(int T) { return ((size_t)T * 13 + 64 + 255) & ~(size_t)255; }

static int mesh_bind_impl(flame_b200_ctx* c, Slot& s, int32_t T, const int32_t* tris, int32_t V,
                          const uint8_t* vtx_valid, const uint8_t* tri_valid, int rows, int cols,
                          bool behind_graph = false) {
  if (T < 0 || rows <= 0 || cols <= 0 || (T > 0 && !tris)) {
    set_error("mesh: bad arguments");
    return FLAME_B200_EINVAL;
  }
  if (T >= (1 << 26)) {  // raster.cu packs (triangle id << 5 | chunk position) into one int
    set_error("mesh: more than 2^26 - 1 triangles");
    return FLAME_B200_EINVAL;
  }
  for (int i = 0; i < 3 * T; ++i)
    if (tris[i] < 0 || tris[i] >= V) {
      set_error("mesh: triangle vertex index out of range");
      return FLAME_B200_EINVAL;
    }
  if (rows > FB_MAX_IMAGE_SIDE || cols > FB_MAX_IMAGE_SIDE) {
    set_error("mesh: image side above 2896 px (the reference's fp32 edge functions stop being exact there)");
    return FLAME_B200_EINVAL;
  }
  if (T > s.capT) {
    const int n = std::max(T, s.capT + s.capT / 2);
    s.capT = 0;
    if (int rc = regrow(&s.d_tris, (size_t)n * 3 * sizeof(int))) return rc;
    if (int rc = regrow(&s.d_tri_ok, (size_t)n)) return rc;
    if (int rc = regrow(&s.d_tri_filt, (size_t)n)) return rc;
    if (int rc = regrow(&s.m.stat, (size_t)n * sizeof(TriStatic))) return rc;
    if (int rc = regrow(&s.m.wide, (size_t)n * sizeof(int))) return rc;
    s.capT = n;
  }
  if (s.own_img) {  // a new bind ends a redirection of the dense map
    s.m.img = s.own_img;
    s.own_img = nullptr;
  }
  s.values_dst = nullptr;
  const size_t npix = (size_t)rows * cols;
  if (npix > s.cap_pix) {
    s.cap_pix = 0;
    if (int rc = regrow(&s.m.img, npix * sizeof(float))) return rc;
    s.cap_pix = npix;
  }
  const int n_tiles = raster_tiles(rows, cols);
  if (n_tiles > s.cap_tiles) {
    s.cap_tiles = 0;
    // [counters 8 | tile_cnt n_tiles]
    if (int rc = regrow(&s.m.counters, ((size_t)n_tiles + 8) * sizeof(int))) return rc;
    s.cap_tiles = n_tiles;
  }
  s.m.tile_cnt = s.m.counters + 8;
  s.m.bin_cap = raster_bin_cap(T, rows, cols);
  const size_t bin_ints = (size_t)n_tiles * s.m.bin_cap;
  if (bin_ints > s.cap_bins) {
    s.cap_bins = 0;
    if (int rc = regrow(&s.m.bins, bin_ints * sizeof(int))) return rc;
    s.cap_bins = bin_ints;
  }
  s.bins_dirty = true;
  s.m.tris = s.d_tris;
  s.m.tri_ok = s.d_tri_ok;
  s.m.tri_filt = nullptr;  // a new triangulation starts unfiltered
  s.m.T = T;
  s.m.rows = rows;
  s.m.cols = cols;
  if (T > 0) {
    // staged through the slot's pinned blob (tris + combined validity): at its start after waiting for the
    // blob's previous use, or - behind_graph - in the space the graph upload of the SAME frame reserved
    // behind its own staging image (Slot::stage_extra), which needs no wait
    const size_t need = mesh_stage_bytes(T);
    size_t off = 0;
    if (behind_graph && s.stage_used + need <= s.h_blob_cap) {
      off = s.stage_used;
    } else {
      if (int rc = wait_blob(s)) return rc;
      if (need > s.h_blob_cap) {
        s.h_blob_cap = 0;
        if (s.h_blob) FB_CUDA(cudaFreeHost(s.h_blob));
        s.h_blob = nullptr;
        FB_CUDA(cudaMallocHost(reinterpret_cast<void**>(&s.h_blob), need));
        s.h_blob_cap = need;
      }
    }
    if (!s.blob_free) FB_CUDA(cudaEventCreateWithFlags(&s.blob_free, cudaEventDisableTiming));
    int* ht = reinterpret_cast<int*>(s.h_blob + off);
    std::memcpy(ht, tris, (size_t)T * 12);
    FB_CUDA(cudaMemcpyAsync(s.d_tris, ht, (size_t)T * 12, cudaMemcpyHostToDevice, c->stream));
    if (!tri_valid && !vtx_valid) {
      FB_CUDA(cudaMemsetAsync(s.d_tri_ok, 1, (size_t)T, c->stream));
    } else {
      unsigned char* hok = s.h_blob + off + (size_t)T * 12;
      for (int t = 0; t < T; ++t) {
        bool ok = !tri_valid || tri_valid[t];
        if (ok && vtx_valid)
          ok = vtx_valid[tris[3 * t]] && vtx_valid[tris[3 * t + 1]] && vtx_valid[tris[3 * t + 2]];
        hok[t] = ok ? 1 : 0;
      }
      FB_CUDA(cudaMemcpyAsync(s.d_tri_ok, hok, (size_t)T, cudaMemcpyHostToDevice, c->stream));
    }
    FB_CUDA(cudaEventRecord(s.blob_free, c->stream));
  }
  return 0;
}

static int mesh_bind_slot(flame_b200_ctx* c, int slot, int32_t T, const int32_t* tris,
                          const uint8_t* vtx_valid, const uint8_t* tri_valid, int rows, int cols,
                          bool behind_graph) {
  if (int rc = check_slot(c, slot)) return rc;
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded) {
    set_error("mesh_bind: slot has no graph");
    return FLAME_B200_ESTATE;
  }
  if (int rc = mesh_bind_impl(c, s, T, tris, s.g.V, vtx_valid, tri_valid, rows, cols, behind_graph)) return rc;
  s.m.pos = s.d_pos;
  s.m.values = s.d_values;
  s.mesh_V = s.g.V;
  return publish_mesh(c, slot);
}

int flame_b200_mesh_bind(flame_b200_ctx* c, int slot, int32_t T, const int32_t* tris,
                         const uint8_t* vtx_valid, const uint8_t* tri_valid, int rows, int cols) {
  return mesh_bind_slot(c, slot, T, tris, vtx_valid, tri_valid, rows, cols, false);
}

int flame_b200_mesh_project(flame_b200_ctx* c, int slot0, int n_slots, float scale) {
  if (int rc = check_slot(c, slot0, n_slots)) return rc;
  if (int rc = use_device(c)) return rc;
  int maxV = 0, maxT = 0, max_rows = 0, max_cols = 0;
  bool rebuild = false;
  for (int i = slot0; i < slot0 + n_slots; ++i) {
    Slot& s = c->slots[i];
    if (!s.loaded || !s.m.img || s.m.values != s.d_values || s.m.pos != s.d_pos) {
      set_error("mesh_project: slot has no graph / bound mesh");
      return FLAME_B200_ESTATE;
    }
    if (s.mesh_V != s.g.V) {
      set_error("mesh_project: the graph changed size since mesh_bind; bind the new triangulation first");
      return FLAME_B200_ESTATE;
    }
    maxV = std::max(maxV, s.g.V);
    maxT = std::max(maxT, s.m.T);
    max_rows = std::max(max_rows, s.m.rows);
    max_cols = std::max(max_cols, s.m.cols);
    rebuild |= s.bins_dirty;
  }
  // values_from_graph overwrites what a pending projection of the same slots still reads
  if (int rc = join_range(c, slot0, slot0 + n_slots)) return rc;
  if (c->pending.size() >= 64)
    if (int rc = join_side(c)) return rc;
  if (int rc = raster_values_from_graph(c->stream, c->d_graphs + slot0, c->d_meshes + slot0, n_slots,
                                        maxV, scale))
    return rc;
  for (int i = slot0; i < slot0 + n_slots; ++i) {
    Slot& s = c->slots[i];
    if (s.values_dst && s.g.V > 0)  // peer copy of vtx_idepths (UVA: the destination may live on another GPU)
      FB_CUDA(cudaMemcpyAsync(s.values_dst, s.d_values, (size_t)s.g.V * sizeof(float), cudaMemcpyDefault, c->stream));
  }
  cudaStream_t rs = c->stream;
  if (c->async_proj) {
    FB_CUDA(cudaEventRecord(c->ev_vals, c->stream));
    FB_CUDA(cudaStreamWaitEvent(c->side, c->ev_vals, 0));
    rs = c->side;
  }
  if (rebuild)  // re-arm [wide-list counter | per-tile counts] of every mesh of the batch (all are rebuilt)
    for (int i = slot0; i < slot0 + n_slots; ++i) {
      Slot& s = c->slots[i];
      FB_CUDA(cudaMemsetAsync(s.m.counters, 0, ((size_t)raster_tiles(s.m.rows, s.m.cols) + 8) * sizeof(int), rs));
      s.bins_dirty = false;
    }
  if (int rc = raster_project(rs, c->d_meshes + slot0, n_slots, maxT, max_rows, max_cols, 1, rebuild)) return rc;
  if (c->async_proj) {
    cudaEvent_t done = nullptr;
    if (!c->ev_pool.empty()) {
      done = c->ev_pool.back();
      c->ev_pool.pop_back();
    } else {
      FB_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    }
    FB_CUDA(cudaEventRecord(done, c->side));
    c->pending.push_back({slot0, slot0 + n_slots, done});
  }
  return 0;
}

int flame_b200_mesh_download(flame_b200_ctx* c, int slot, float* out_host) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!out_host) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  if (!s.m.img) {
    set_error("mesh_download: no dense map in this slot");
    return FLAME_B200_ESTATE;
  }
  FB_CUDA(cudaMemcpyAsync(out_host, s.m.img, (size_t)s.m.rows * s.m.cols * sizeof(float),
                          cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

// ---- peer memory for the batched-frame case -------------------------------------------------------
int flame_b200_device_alloc(int device, size_t bytes, void** out) {
  if (!out) return FLAME_B200_EINVAL;
  FB_CUDA(cudaSetDevice(device));
  FB_CUDA(cudaMalloc(out, bytes ? bytes : 1));
  return 0;
}

int flame_b200_device_free(void* p) {
  if (p) FB_CUDA(cudaFree(p));
  return 0;
}

int flame_b200_ipc_export(void* dev_ptr, uint8_t handle[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!dev_ptr || !handle) return FLAME_B200_EINVAL;
  cudaIpcMemHandle_t h;
  FB_CUDA(cudaIpcGetMemHandle(&h, dev_ptr));
  std::memcpy(handle, &h, 64);
  return 0;
}

int flame_b200_ipc_open(int device, const uint8_t handle[64], void** out) {
  if (!handle || !out) return FLAME_B200_EINVAL;
  FB_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  FB_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int flame_b200_ipc_close(void* p) {
  if (p) FB_CUDA(cudaIpcCloseMemHandle(p));
  return 0;
}

int flame_b200_mesh_redirect(flame_b200_ctx* c, int slot, void* dense_map_dst, void* values_dst) {
  if (int rc = check_slot(c, slot)) return rc;
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded || !s.m.img || s.m.values != s.d_values) {
    set_error("mesh_redirect: slot has no graph / bound mesh");
    return FLAME_B200_ESTATE;
  }
  if (!s.own_img) s.own_img = s.m.img;
  s.m.img = dense_map_dst ? static_cast<float*>(dense_map_dst) : s.own_img;
  s.values_dst = static_cast<float*>(values_dst);
  if (!dense_map_dst) s.own_img = nullptr;
  return publish_mesh(c, slot);
}

int flame_b200_slot_device_ptrs(flame_b200_ctx* c, int slot, void** dense_map, void** values, int32_t* rows,
                                int32_t* cols, int32_t* V) {
  if (int rc = check_slot(c, slot)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded || !s.m.img || s.m.values != s.d_values) {
    set_error("slot_device_ptrs: slot has no graph / bound mesh");
    return FLAME_B200_ESTATE;
  }
  if (dense_map) *dense_map = s.m.img;
  if (values) *values = s.d_values;
  if (rows) *rows = s.m.rows;
  if (cols) *cols = s.m.cols;
  if (V) *V = s.g.V;
  return 0;
}

void flame_b200_tri_filter_default_params(flame_b200_tri_filter_params* p) {
  if (!p) return;
  p->do_oblique_triangle_filter = 1;  // params.h:70-85
  p->oblique_normal_thresh = 1.39626f;
  p->oblique_idepth_diff_factor = 0.35f;
  p->oblique_idepth_diff_abs = 0.1f;
  p->do_edge_length_filter = 1;
  p->edge_length_thresh = 0.333f;
  p->do_idepth_triangle_filter = 1;
  p->min_triangle_idepth = 0.01f;
}

int flame_b200_mesh_filter(flame_b200_ctx* c, int slot, const flame_b200_tri_filter_params* params,
                           const float Kinv[9], float scale, uint8_t* tri_valid_out, int apply) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!params || !Kinv) {
    set_error("mesh_filter: bad arguments");
    return FLAME_B200_EINVAL;
  }
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded || !s.m.img || s.m.values != s.d_values || s.m.pos != s.d_pos || s.mesh_V != s.g.V) {
    set_error("mesh_filter: slot has no graph / mesh bound to the current graph");
    return FLAME_B200_ESTATE;
  }
  const int T = s.m.T;
  if (int rc = raster_values_from_graph(c->stream, c->d_graphs + slot, c->d_meshes + slot, 1, s.g.V, scale))
    return rc;
  if (int rc = raster_tri_filter(c->stream, s.m, *params, Kinv, s.d_tri_filt)) return rc;
  if (apply) {
    s.m.tri_filt = s.d_tri_filt;
    if (int rc = publish_mesh(c, slot)) return rc;
  }
  if (tri_valid_out && T > 0) {
    if (int rc = wait_blob(s)) return rc;
    if ((size_t)T > s.h_blob_cap) {
      set_error("internal: staging blob too small");
      return FLAME_B200_ENOMEM;
    }
    FB_CUDA(cudaMemcpyAsync(s.h_blob, s.d_tri_filt, (size_t)T, cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaEventRecord(s.blob_free, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    std::memcpy(tri_valid_out, s.h_blob, (size_t)T);
  }
  return 0;
}

int flame_b200_mesh_filter_angles(flame_b200_ctx* c, int slot, const float Kinv[9], float scale,
                                  float* angles_out) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!Kinv || !angles_out) {
    set_error("mesh_filter_angles: bad arguments");
    return FLAME_B200_EINVAL;
  }
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded || !s.m.img || s.m.values != s.d_values || s.m.pos != s.d_pos || s.mesh_V != s.g.V) {
    set_error("mesh_filter_angles: slot has no graph / mesh bound to the current graph");
    return FLAME_B200_ESTATE;
  }
  const int T = s.m.T;
  if (T == 0) return 0;
  if ((size_t)T * 5 > raw_f_total(s.g.V, s.g.E) * 4 || (size_t)T * 4 > s.h_blob_cap) {
    set_error("mesh_filter_angles: more triangles than the slot's scratch holds");
    return FLAME_B200_EINVAL;
  }
  if (int rc = wait_blob(s)) return rc;
  if (int rc = raster_values_from_graph(c->stream, c->d_graphs + slot, c->d_meshes + slot, 1, s.g.V, scale))
    return rc;
  flame_b200_tri_filter_params p;
  flame_b200_tri_filter_default_params(&p);
  // the raw staging area is scratch between uploads: [angles T floats | validity T bytes (discarded)]
  float* d_ang = s.d_raw();
  if (int rc = raster_tri_filter(c->stream, s.m, p, Kinv, reinterpret_cast<unsigned char*>(d_ang + T), d_ang))
    return rc;
  FB_CUDA(cudaMemcpyAsync(s.h_blob, d_ang, (size_t)T * 4, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaEventRecord(s.blob_free, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  std::memcpy(angles_out, s.h_blob, (size_t)T * 4);
  return 0;
}

int flame_b200_mesh_normals(flame_b200_ctx* c, int slot, const float Kinv[9], float scale, float* normals_out) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!Kinv || !normals_out) {
    set_error("mesh_normals: bad arguments");
    return FLAME_B200_EINVAL;
  }
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  if (!s.loaded || !s.m.img || s.m.values != s.d_values || s.m.pos != s.d_pos || s.mesh_V != s.g.V) {
    set_error("mesh_normals: slot has no graph / mesh bound to the current graph");
    return FLAME_B200_ESTATE;
  }
  const int V = s.g.V, T = s.m.T;
  if (V == 0) return 0;
  const size_t need = (size_t)2 * (V + 1) + (size_t)3 * T;
  if (need > s.vt_cap) {
    if (s.d_vt) FB_CUDA(cudaFree(s.d_vt));
    s.d_vt = nullptr;
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.d_vt), need * sizeof(int)));
    s.vt_cap = need;
  }
  if (int rc = wait_blob(s)) return rc;
  if (int rc = raster_values_from_graph(c->stream, c->d_graphs + slot, c->d_meshes + slot, 1, V, scale)) return rc;
  // the raw staging area (>= 13V floats) is scratch between uploads
  if (int rc = raster_vertex_normals(c->stream, s.m, V, Kinv, s.d_vt, s.d_raw())) return rc;
  FB_CUDA(cudaMemcpyAsync(s.h_blob, s.d_blob, (size_t)3 * V * 4, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaEventRecord(s.blob_free, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  std::memcpy(normals_out, s.h_blob, (size_t)3 * V * 4);
  return 0;
}

int flame_b200_mesh_coverage(flame_b200_ctx* c, int slot, int64_t* covered) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!covered) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  if (!s.m.img) {
    set_error("mesh_coverage: no dense map in this slot");
    return FLAME_B200_ESTATE;
  }
  unsigned long long* d_count = reinterpret_cast<unsigned long long*>(c->d_cost);
  if (int rc = raster_coverage(c->stream, s.m, d_count)) return rc;
  FB_CUDA(cudaMemcpyAsync(c->h_cost, c->d_cost, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  *covered = (int64_t)*reinterpret_cast<unsigned long long*>(c->h_cost);
  return 0;
}

int flame_b200_interpolate_mesh(flame_b200_ctx* c, int slot, int32_t T, const int32_t* tris,
                                int32_t V, const float* pos, const float* values, float scale,
                                const uint8_t* vtx_valid, const uint8_t* tri_valid, int rows,
                                int cols, int fill_nan, float* out_host) {
  if (int rc = check_slot(c, slot)) return rc;
  if (!out_host || V < 0 || (V > 0 && !pos)) {
    set_error("interpolate_mesh: bad arguments");
    return FLAME_B200_EINVAL;
  }
  if (int rc = use_device(c)) return rc;
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  Slot& s = c->slots[slot];
  if (!values && (!s.loaded || s.g.V != V)) {
    set_error("interpolate_mesh: values == NULL needs a graph with V vertices in the slot");
    return FLAME_B200_ESTATE;
  }
  if (V > s.capMV) {
    const int n = std::max(V, s.capMV + s.capMV / 2);
    if (s.d_mpos) FB_CUDA(cudaFree(s.d_mpos));
    if (s.d_mvalues) FB_CUDA(cudaFree(s.d_mvalues));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.d_mpos), (size_t)n * 2 * sizeof(float)));
    FB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.d_mvalues), (size_t)n * sizeof(float)));
    s.capMV = n;
  }
  if (int rc = mesh_bind_impl(c, s, T, tris, V, vtx_valid, tri_valid, rows, cols)) return rc;
  if (V > 0)
    FB_CUDA(cudaMemcpyAsync(s.d_mpos, pos, (size_t)2 * V * 4, cudaMemcpyHostToDevice, c->stream));
  s.m.pos = s.d_mpos;
  s.m.values = s.d_mvalues;
  if (int rc = publish_mesh(c, slot)) return rc;
  if (values) {
    if (V > 0)
      FB_CUDA(cudaMemcpyAsync(s.d_mvalues, values, (size_t)V * 4, cudaMemcpyHostToDevice, c->stream));
  } else {
    if (int rc = raster_values_from_graph(c->stream, c->d_graphs + slot, c->d_meshes + slot, 1, s.g.V, scale))
      return rc;
  }
  const size_t npix = (size_t)rows * cols;
  if (!fill_nan)
    FB_CUDA(cudaMemcpyAsync(s.m.img, out_host, npix * 4, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(cudaMemsetAsync(s.m.counters, 0, ((size_t)raster_tiles(rows, cols) + 8) * sizeof(int), c->stream));
  if (int rc = raster_project(c->stream, c->d_meshes + slot, 1, s.m.T, rows, cols, fill_nan, 1)) return rc;
  s.bins_dirty = true;  // bound to caller-supplied positions: never reused
  FB_CUDA(cudaMemcpyAsync(out_host, s.m.img, npix * 4, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

// ---- whole frame in one call ---------------------------------------------------------------------
namespace {
inline double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
// Library calls issue their work on c->stream; inside a frame submission that is the slot's own stream,
// in stream order (no second projection stream: other slots' frames fill the machine instead).
struct StreamScope {
  flame_b200_ctx* c;
  cudaStream_t saved;
  bool saved_async;
  StreamScope(flame_b200_ctx* c_, cudaStream_t st) : c(c_), saved(c_->stream), saved_async(c_->async_proj) {
    c->stream = st;
    c->async_proj = false;
  }
  ~StreamScope() {
    c->stream = saved;
    c->async_proj = saved_async;
  }
};
}  // namespace

// Launches the solver, the projection and the downloads of the frames collected in c->wave (consecutive
// slots, same parameters) as ONE batch on the context's main stream: a single solver launch whose clusters
// fill the device together, one batched projection, and the device->host copies on the context's download
// stream so that the next wave's kernels overlap them.
static int wave_flush(flame_b200_ctx* c) {
  if (c->wave.empty()) return 0;
  const double t0 = now_s();
  const int a = c->wave.front(), n = (int)c->wave.size();
  c->wave.clear();
  if (!c->down) FB_CUDA(cudaStreamCreateWithFlags(&c->down, cudaStreamNonBlocking));
  if (!c->ev_wave) FB_CUDA(cudaEventCreateWithFlags(&c->ev_wave, cudaEventDisableTiming));
  for (int i = a; i < a + n; ++i) FB_CUDA(cudaStreamWaitEvent(c->stream, c->slots[i].uploaded, 0));
  const bool saved_async = c->async_proj;
  c->async_proj = false;  // in stream order: the wave's downloads wait for one event on the main stream
  int rc = flame_b200_pd_run(c, a, n, &c->wave_params, c->wave_iters, FLAME_B200_PD_AUTO);
  if (!rc) rc = flame_b200_mesh_project(c, a, n, c->wave_scale);
  c->async_proj = saved_async;
  if (rc) return rc;
  for (int i = a; i < a + n; ++i) {
    Slot& s = c->slots[i];
    if (s.out_V > 0)  // x | w1 | w2 (caller order) through the raw staging area, scratch between uploads
      if ((rc = pd_unpack_xw(c->stream, s.g, s.d_raw()))) return rc;
  }
  FB_CUDA(cudaEventRecord(c->ev_wave, c->stream));
  FB_CUDA(cudaStreamWaitEvent(c->down, c->ev_wave, 0));
  for (int i = a; i < a + n; ++i) {
    Slot& s = c->slots[i];
    if (s.out_map) {
      FB_CUDA(cudaMemcpyAsync(s.out_map, s.m.img, (size_t)s.m.rows * s.m.cols * sizeof(float),
                              cudaMemcpyDeviceToHost, c->down));
      c->d2h_bytes += (uint64_t)s.m.rows * s.m.cols * 4;
    }
    if (s.out_V > 0) {
      FB_CUDA(cudaMemcpyAsync(s.h_blob + s.out_off, s.d_raw(), (size_t)3 * s.out_V * 4, cudaMemcpyDeviceToHost,
                              c->down));
      c->d2h_bytes += (uint64_t)3 * s.out_V * 4;
    }
    FB_CUDA(cudaEventRecord(s.blob_free, c->down));
    FB_CUDA(cudaEventRecord(s.done, c->down));
    s.launched = true;
  }
  c->t_enqueue += now_s() - t0;
  return 0;
}

static bool same_params(const flame_b200_pd_params& a, const flame_b200_pd_params& b) {
  return std::memcmp(&a, &b, sizeof(a)) == 0;
}

int flame_b200_process_frame_submit(flame_b200_ctx* c, int slot, int32_t V, const float* pos,
                                    const float* idepth, const float* weight, int32_t E,
                                    const int32_t* edges, int32_t T, const int32_t* tris, int rows, int cols,
                                    const flame_b200_pd_params* params, int n_iters, float scale,
                                    float* out_map) {
  if (int rc = check_slot(c, slot)) return rc;
  if (int rc = use_device(c)) return rc;
  Slot& s = c->slots[slot];
  if (s.in_flight) {
    set_error("process_frame_submit: the slot still has a frame in flight (call flame_b200_process_frame_wait)");
    return FLAME_B200_ESTATE;
  }
  if (T < 0 || V < 0 || !params || n_iters < 0) {
    set_error("process_frame_submit: bad arguments");
    return FLAME_B200_EINVAL;
  }
  if (!edges && !tris && V > 0) {
    // no topology given: triangulate the vertices here, on the calling thread (utils/delaunay.cc:31-77's place in
    // Flame::update) - the device keeps working on the frames submitted before. The arrays stay with the slot
    // (flame_b200_process_frame_topology).
    const double tt = now_s();
    s.h_tris.resize((size_t)6 * V + 3);
    s.h_edges.resize((size_t)6 * V + 2);
    int32_t nt = 0, ne = 0;
    if (int rc = flame_b200_delaunay(V, pos, 2 * V + 1, &nt, s.h_tris.data(), 3 * V + 1, &ne, s.h_edges.data()))
      return rc;
    s.h_nt = nt;
    s.h_ne = ne;
    T = nt;
    E = ne;
    tris = s.h_tris.data();
    edges = s.h_edges.data();
    c->t_tri += now_s() - tt;
  } else {
    s.h_nt = s.h_ne = -1;
  }
  if (!s.stream) FB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
  if (!s.done) FB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
  if (!s.uploaded) FB_CUDA(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
  // a wave is a run of consecutive slots with identical solver parameters
  if (!c->wave.empty() && (slot != c->wave.back() + 1 || !same_params(*params, c->wave_params) ||
                           n_iters != c->wave_iters || scale != c->wave_scale))
    if (int rc = wave_flush(c)) return rc;
  // work issued on the context's main stream for this slot earlier (uploads, projections) comes first
  if (int rc = join_range(c, slot, slot + 1)) return rc;
  FB_CUDA(cudaEventRecord(s.done, c->stream));
  FB_CUDA(cudaStreamWaitEvent(s.stream, s.done, 0));
  const double t0 = now_s();
  {
    // upload + device-side layout on the slot's own stream: overlaps the waves running on the main stream
    StreamScope scope(c, s.stream);
    // one pinned staging blob per slot: [graph image | triangles | x, w1, w2 coming back]
    s.stage_extra = mesh_stage_bytes(T) + (((size_t)3 * V * 4 + 255) & ~(size_t)255);
    int rc = flame_b200_graph_init(c, slot, V, pos, idepth, weight, E, edges);
    s.stage_extra = 0;
    if (rc) return rc;
    if ((rc = mesh_bind_slot(c, slot, T, tris, nullptr, nullptr, rows, cols, true))) return rc;
    FB_CUDA(cudaEventRecord(s.uploaded, s.stream));
  }
  s.out_V = V;
  s.out_off = s.stage_used + mesh_stage_bytes(T);
  s.out_map = out_map;
  s.in_flight = true;
  s.launched = false;
  c->n_frames++;
  c->h2d_bytes += (uint64_t)V * 16 + (uint64_t)V * 4 + (uint64_t)E * 8 + (uint64_t)T * 12;
  c->t_stage += now_s() - t0;  // includes t_prep (accounted separately inside graph_upload_impl)
  // how many frames like this one the device solves concurrently = the wave size
  int cap = 8;
  if (s.plan.cluster > 0) {
    if (s.plan.cluster != c->cap_cluster || s.plan.variant != c->cap_variant || s.plan.smem_bytes > c->cap_smem) {
      int n = 0;
      if (int rc = pd_resident_capacity(s.plan, &n)) return rc;
      c->cap_cluster = s.plan.cluster;
      c->cap_variant = s.plan.variant;
      c->cap_smem = s.plan.smem_bytes;
      c->cap_frames = std::max(n, 1);
    }
    cap = c->cap_frames;
  }
  c->wave.push_back(slot);
  c->wave_params = *params;
  c->wave_iters = n_iters;
  c->wave_scale = scale;
  if ((int)c->wave.size() >= std::min(cap, std::max(1, c->n_slots / 2))) return wave_flush(c);
  return 0;
}

int flame_b200_process_frame_wait(flame_b200_ctx* c, int slot, float* out_x, float* out_w1, float* out_w2) {
  if (int rc = check_slot(c, slot)) return rc;
  if (int rc = use_device(c)) return rc;
  Slot& s = c->slots[slot];
  if (!s.in_flight) {
    set_error("process_frame_wait: no frame in flight in this slot");
    return FLAME_B200_ESTATE;
  }
  if (!s.launched)  // its wave is still collecting frames
    if (int rc = wave_flush(c)) return rc;
  const double t0 = now_s();
  s.in_flight = false;
  FB_CUDA(cudaEventSynchronize(s.done));
  c->t_wait += now_s() - t0;
  const size_t V = (size_t)s.out_V;
  const float* hf = reinterpret_cast<const float*>(s.h_blob + s.out_off);
  if (V > 0) {
    if (out_x) std::memcpy(out_x, hf, V * 4);
    if (out_w1) std::memcpy(out_w1, hf + V, V * 4);
    if (out_w2) std::memcpy(out_w2, hf + 2 * V, V * 4);
  }
  return 0;
}

int flame_b200_process_frame(flame_b200_ctx* c, int slot, int32_t V, const float* pos,
                             const float* idepth, const float* weight, int32_t E,
                             const int32_t* edges, int32_t T, const int32_t* tris, int rows, int cols,
                             const flame_b200_pd_params* params, int n_iters, float scale,
                             float* out_map, float* out_x, float* out_w1, float* out_w2) {
  if (int rc = flame_b200_process_frame_submit(c, slot, V, pos, idepth, weight, E, edges, T, tris, rows, cols,
                                               params, n_iters, scale, out_map))
    return rc;
  return flame_b200_process_frame_wait(c, slot, out_x, out_w1, out_w2);
}

int flame_b200_process_frame_topology(flame_b200_ctx* c, int slot, int32_t tri_capacity, int32_t* n_tris,
                                      int32_t* tris, int32_t edge_capacity, int32_t* n_edges, int32_t* edges) {
  if (int rc = check_slot(c, slot)) return rc;
  const Slot& s = c->slots[slot];
  if (!n_tris || !n_edges || tri_capacity < 0 || edge_capacity < 0) {
    set_error("process_frame_topology: bad arguments");
    return FLAME_B200_EINVAL;
  }
  if (s.h_nt < 0) {
    set_error("process_frame_topology: the slot's last frame was submitted with the caller's own triangulation");
    return FLAME_B200_ESTATE;
  }
  *n_tris = s.h_nt;
  *n_edges = s.h_ne;
  if ((tris && s.h_nt > tri_capacity) || (edges && s.h_ne > edge_capacity)) {
    set_error("process_frame_topology: output capacity too small");
    return FLAME_B200_ENOMEM;
  }
  if (tris) std::memcpy(tris, s.h_tris.data(), (size_t)3 * s.h_nt * sizeof(int32_t));
  if (edges) std::memcpy(edges, s.h_edges.data(), (size_t)2 * s.h_ne * sizeof(int32_t));
  return 0;
}

int flame_b200_ctx_stats(flame_b200_ctx* c, flame_b200_stats* out, int reset) {
  if (!c || !out) return FLAME_B200_EINVAL;
  out->frames = c->n_frames;
  out->h2d_bytes = c->h2d_bytes;
  out->d2h_bytes = c->d2h_bytes;
  out->prep_seconds = c->t_prep;
  out->stage_seconds = c->t_stage;
  out->enqueue_seconds = c->t_enqueue;
  out->wait_seconds = c->t_wait;
  out->triangulation_seconds = c->t_tri;
  if (reset) {
    c->n_frames = c->h2d_bytes = c->d2h_bytes = 0;
    c->t_prep = c->t_stage = c->t_enqueue = c->t_wait = c->t_tri = 0;
  }
  return 0;
}

// ---- epipolar search ------------------------------------------------------------------------
int flame_b200_set_camera(flame_b200_ctx* c, int width, int height, const float K[9],
                          const float Kinv[9]) {
  if (!c || !K || !Kinv || width <= 0 || height <= 0) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return stereo_set_camera(&c->stereo, width, height, K, Kinv);
}

int flame_b200_frame_upload(flame_b200_ctx* c, int fslot, const uint8_t* img_pad,
                            const float* gradx_pad, const float* grady_pad, int pad,
                            const float pose[7]) {
  if (!c || !img_pad || !gradx_pad || !grady_pad || !pose) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return stereo_frame_upload(&c->stereo, c->stream, fslot, img_pad, gradx_pad, grady_pad, pad, pose);
}

int flame_b200_frame_upload_raw(flame_b200_ctx* c, int fslot, const uint8_t* img, int pad,
                                const float pose[7]) {
  if (!c || !img || !pose) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return stereo_frame_upload_raw(&c->stereo, c->stream, fslot, img, pad, pose);
}

int flame_b200_frame_download(flame_b200_ctx* c, int fslot, uint8_t* img_pad, float* gradx_pad,
                              float* grady_pad) {
  if (!c) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return stereo_frame_download(c->stereo, c->stream, fslot, img_pad, gradx_pad, grady_pad);
}

int flame_b200_features_upload(flame_b200_ctx* c, const flame_b200_feature* feats, int32_t n_feats) {
  if (!c || (n_feats > 0 && !feats) || n_feats < 0) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return stereo_features_upload(c->stereo, c->stream, feats, n_feats);
}

int flame_b200_features_track(flame_b200_ctx* c, const flame_b200_track_params* params, int fnew, int curr_pf) {
  if (!c || !params) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return stereo_features_track(c->stereo, c->stream, params, fnew, curr_pf);
}

int flame_b200_features_download(flame_b200_ctx* c, flame_b200_feature* feats, flame_b200_track_diag* diag,
                                 int32_t* n_updates, int commit) {
  if (!c) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return stereo_features_download(c->stereo, c->stream, feats, diag, n_updates, commit);
}

int flame_b200_update_feature_idepths(flame_b200_ctx* c, const flame_b200_track_params* params,
                                      int fnew, int curr_pf, flame_b200_feature* feats,
                                      int32_t n_feats, flame_b200_track_diag* diag,
                                      int32_t* n_updates) {
  if (!c || !params || (n_feats > 0 && !feats) || n_feats < 0) return FLAME_B200_EINVAL;
  if (int rc = use_device(c)) return rc;
  return stereo_update(c->stereo, c->stream, params, fnew, curr_pf, feats, n_feats, diag, n_updates);
}

}  // extern "C"
