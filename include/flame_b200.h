/*
 * flame_b200.h — C ABI of libflame_b200.so: the B200-native (sm_100a) replacement for FLaME's
 * per-frame variational depth regularisation hot path.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * robustrobotics/flame checkout, src/flame/...). Plain pointers and sizes only; no C++ or torch
 * types cross this boundary. Conventions follow the reference's:
 *   - the caller owns every host buffer; the library copies what it needs before returning
 *     (uploads are staged through pinned memory), so buffers may be reused immediately;
 *   - functions return 0 on success, a negative FLAME_B200_E* code otherwise, never throw;
 *     flame_b200_last_error() returns a thread-local description of the last failure;
 *   - one context = one CUDA stream + `n_slots` independent frame slots. A context is
 *     thread-compatible (serialise calls on it, as Flame serialises step() under graph_mtx_,
 *     flame.cc:100-111); distinct contexts may be driven from distinct threads concurrently;
 *   - work is enqueued asynchronously on the context's stream; calls that return data to
 *     host memory synchronise that stream before returning unless documented otherwise.
 *   - there is NO CPU fallback: with no usable sm_100 device every call fails with
 *     FLAME_B200_ENODEV.
 *
 * Vertex order = caller's order. Edge order = caller's order; edges[2e] is the SOURCE and
 * edges[2e+1] the TARGET of edge e (flame.cc:2085-2096 inserts Triangle's edgelist in order;
 * orientation matters, nltgv2_l1_graph_regularizer.cc:99-101,129-133). Results are bit-identical
 * to the reference's scalar fp32 arithmetic for the PD iteration, the edge weights and the dense
 * projection (see DESIGN.md "Parity").
 */
#ifndef FLAME_B200_H_
#define FLAME_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLAME_B200_VERSION 100

enum {
  FLAME_B200_OK = 0,
  FLAME_B200_EINVAL = -1,   /* bad argument (the reference would FLAME_ASSERT) */
  FLAME_B200_ENODEV = -2,   /* no sm_100 device / CUDA runtime unusable */
  FLAME_B200_ECUDA = -3,    /* a CUDA call failed; see flame_b200_last_error() */
  FLAME_B200_ENOMEM = -4,
  FLAME_B200_ESTATE = -5    /* call sequence error (e.g. pd_run before graph_upload) */
};

typedef struct flame_b200_ctx flame_b200_ctx;

const char* flame_b200_last_error(void);
int flame_b200_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py "gpu_launches"). */
uint64_t flame_b200_kernel_launches(void);

/* ---- context ------------------------------------------------------------------------- */
/* device: CUDA ordinal. n_slots >= 1 independent frames resident at once. */
int flame_b200_ctx_create(int device, int n_slots, flame_b200_ctx** out);
void flame_b200_ctx_destroy(flame_b200_ctx* ctx);
int flame_b200_ctx_sync(flame_b200_ctx* ctx);

/* Optional pipelining of the dense projection. With on != 0, flame_b200_mesh_project issues its kernels
 * on a second stream of the context, so that the solver launches of the NEXT frames (graph_reset,
 * pd_run on flame_b200_ctx_stream) overlap them: a frame cluster occupies 16 SMs and at most 7 clusters
 * are resident on a B200, the projection fills the remaining SMs. Every library call that reads or
 * replaces projection inputs/outputs (mesh_download, mesh_bind, interpolate_mesh, graph uploads,
 * ctx_sync) orders itself after the pending projection; callers that put their OWN work or events on
 * flame_b200_ctx_stream call flame_b200_ctx_join first (device-side wait, does not block the host). */
int flame_b200_ctx_set_async_projection(flame_b200_ctx* ctx, int on);
int flame_b200_ctx_join(flame_b200_ctx* ctx);
/* cudaStream_t of the context, as an opaque pointer (for CUDA-event timing by the caller). */
void* flame_b200_ctx_stream(flame_b200_ctx* ctx);

/* How graphs uploaded AFTER this call are laid out for the resident solver. A frame's primal-dual state
 * lives in the shared memory of one thread-block cluster for the whole solve:
 *   FLAME_B200_PREF_THROUGHPUT (default)  the smallest cluster that holds the graph (9 CTAs for 10 000
 *       Delaunay vertices), so that as many frames as possible are solved concurrently (flame_b200_pd_wave);
 *   FLAME_B200_PREF_LATENCY               the frame is spread over more SMs (up to 16): a single frame
 *       finishes sooner, fewer frames are resident at once.
 * Results are bit-identical either way. */
enum { FLAME_B200_PREF_THROUGHPUT = 0, FLAME_B200_PREF_LATENCY = 1 };
int flame_b200_ctx_set_pd_preference(flame_b200_ctx* ctx, int pref);

/* Pinned (page-locked) host memory for buffers handed to the asynchronous download calls. */
int flame_b200_host_alloc(size_t bytes, void** out);
int flame_b200_host_free(void* p);

/* ---- PD solver: optimizers/nltgv2_l1_graph_regularizer.{h,cc} ------------------------- */

/* nltgv2_l1_graph_regularizer::Params (h:121-129), same fields, same defaults. */
typedef struct {
  float data_factor; /* 0.1   */
  float step_x;      /* 0.001 */
  float step_q;      /* 125   */
  float theta;       /* 0.25  */
  float x_min;       /* 0     */
  float x_max;       /* 10    */
} flame_b200_pd_params;
void flame_b200_pd_default_params(flame_b200_pd_params* p);

/* Flat SoA image of nltgv2_l1_graph_regularizer::Graph (VertexData h:74-90, EdgeData h:95-102).
 * For upload every pointer except those marked nullable must be set; a NULL nullable array
 * means "all zeros" (w*, q*) / "copy of x" (x_bar, x_prev) / "all ones" (data_weight, beta) /
 * "1/edge length computed on the device" (alpha: the flame.cc:2084-2105 rule).
 * For download any pointer may be NULL (skipped). */
typedef struct {
  int32_t V, E;
  const float* pos;        /* 2V, x,y interleaved (VertexData::pos)           */
  const int32_t* edges;    /* 2E, source,target                                */
  const float* data_term;  /* V                                                */
  const float* data_weight;/* V, nullable -> 1                                 */
  const float* x;          /* V                                                */
  const float* w1;         /* V, nullable -> 0 */
  const float* w2;         /* V, nullable -> 0 */
  const float* x_bar;      /* V, nullable -> x */
  const float* w1_bar;     /* V, nullable -> 0 */
  const float* w2_bar;     /* V, nullable -> 0 */
  const float* x_prev;     /* V, nullable -> x */
  const float* w1_prev;    /* V, nullable -> 0 */
  const float* w2_prev;    /* V, nullable -> 0 */
  const float* alpha;      /* E, nullable -> 1/||pos_i - pos_j|| on device     */
  const float* beta;       /* E, nullable -> 1                                 */
  const float* q1;         /* E, nullable -> 0 */
  const float* q2;         /* E, nullable -> 0 */
  const float* q3;         /* E, nullable -> 0 */
} flame_b200_graph_in;

typedef struct {
  float *x, *w1, *w2, *x_bar, *w1_bar, *w2_bar, *x_prev, *w1_prev, *w2_prev; /* V each */
  float *alpha, *beta, *q1, *q2, *q3;                                         /* E each */
} flame_b200_graph_out;

/* Replaces the property writes of Flame::syncGraph (flame.cc:1998-2014, 2035-2048, 2084-2105,
 * 2160-2162): loads one graph into `slot`. Asynchronous; host buffers are consumed on return. */
int flame_b200_graph_upload(flame_b200_ctx* ctx, int slot, const flame_b200_graph_in* g);

/* Convenience for a fresh graph (flame.cc:2039-2048 + 2084-2105): x = x_bar = x_prev = idepth,
 * data_term = idepth, w = q = 0, data_weight = weight (nullable -> 1), alpha = 1/len, beta = 1. */
int flame_b200_graph_init(flame_b200_ctx* ctx, int slot, int32_t V, const float* pos,
                          const float* idepth, const float* weight, int32_t E,
                          const int32_t* edges);

/* Only refresh the data term / weights of an already uploaded graph (flame.cc:1998-2014). */
int flame_b200_graph_set_data(flame_b200_ctx* ctx, int slot, const float* data_term,
                              const float* data_weight);

/* Re-applies the fresh-vertex rule of Flame::syncGraph (flame.cc:2039-2048) on the device to slots
 * [slot0, slot0 + n_slots): x = x_bar = x_prev = data_term, w* = 0, q* = 0. Asynchronous. */
int flame_b200_graph_reset(flame_b200_ctx* ctx, int slot0, int n_slots);

/* Replaces the property reads at flame.cc:372-380. Synchronises the stream. */
int flame_b200_graph_download(flame_b200_ctx* ctx, int slot, const flame_b200_graph_out* out);

/* ---- keeping the solver state on the device from frame to frame ----------------
 * Flame::update (flame.cc:296-318) moves the graph into every new frame with projectGraph and then
 * edits it in syncGraph; with the two calls below the primal/dual state never leaves the GPU: only
 * the projected positions come back (the caller decides removals and re-triangulates on the host)
 * and only the new topology plus two index maps go up. */

/* Replaces the per-vertex body of Flame::projectGraph (flame.cc:1892-1899): for every vertex
 * (u_new, idepth_new) = EpipolarGeometry::project(pos, x * graph_scale) (epipolar_geometry.h:153-181)
 * with the geometry loaded from T_ref_to_new = (q_wxyz, t) (flame.cc:297-298); the slot keeps
 * pos = u_new and x = idepth_new / graph_scale (x_bar / x_prev untouched, as in the reference), and
 * pos_out[2V] / idepth_out[V] (either nullable) receive u_new / idepth_new so the caller can apply the
 * removal rules (:1901-1917). Needs flame_b200_set_camera. The edge records of the slot are stale
 * afterwards: flame_b200_pd_run refuses the slot until flame_b200_graph_resync (or a new upload).
 * Synchronises the stream. */
int flame_b200_graph_project(flame_b200_ctx* ctx, int slot, const float q_wxyz[4], const float t[3],
                             float graph_scale, float* pos_out, float* idepth_out);

/* The same with the REMOVAL RULES of Flame::projectGraph evaluated on the device (flame.cc:1872-1917):
 * remove_out[v] = 1 when the reference would put vertex v into vtx_to_remove - projected outside
 * valid_region (a cv::Rect_<float>: no rounding here, unlike trackFeature's cv::Rect), behind the camera
 * (idepth_new < 0), or, with do_grad_check_after_projection, onto a pixel of the NEW frame (frame slot
 * `fnew`, flame_b200_frame_upload) whose bilinearly interpolated gradient magnitude is below min_grad_mag.
 * The caller still erases those vertices from its feature maps and re-triangulates. Synchronises. */
typedef struct { /* the fields of flame::Params Flame::projectGraph reads */
  float rescale_factor_max;               /* params.h:59  (border = rescale_factor_max * win_size / 2 + 1) */
  int32_t win_size;                       /* fparams.win_size */
  int32_t do_letterbox;                   /* params.h:49 */
  int32_t do_grad_check_after_projection; /* params.h:42 */
  float min_grad_mag;                     /* params.h:39 */
} flame_b200_project_rules;
int flame_b200_graph_project_marked(flame_b200_ctx* ctx, int slot, const float q_wxyz[4], const float t[3],
                                    float graph_scale, const flame_b200_project_rules* rules, int fnew,
                                    float* pos_out, float* idepth_out, uint8_t* remove_out);

typedef struct {
  int32_t V, E;
  const float* pos;          /* 2V  new positions (feat.xy, flame.cc:2005 / :2041)                   */
  const float* data_term;    /* V   feat.idepth_mu / graph_scale (:2006 / :2042)                     */
  const float* data_weight;  /* V or NULL -> 1 (:2007 / :2043)                                       */
  const int32_t* vtx_from;   /* V   index of the vertex in the slot's CURRENT graph, or -1 if new    */
  const float* x_init;       /* V or NULL: x = x_bar = x_prev of NEW vertices (:2127-2162); NULL ->
                                data_term (:2046-2048). Ignored for carried vertices.               */
  const int32_t* edges;      /* 2E  (source, target) in boost::edges order; a carried edge must keep
                                the orientation it had in the current graph (boost::edge() finds the
                                existing descriptor, :2092-2100)                                    */
  const int32_t* edge_from;  /* E   index of the edge in the current graph, or -1 if new (q = 0)     */
  int32_t check_sticky_obstacles; /* params.check_sticky_obstacles: x = data_term where
                                x - data_term > 0.25 on carried vertices (:2016-2019)               */
} flame_b200_graph_sync;

/* Replaces the graph edits of Flame::syncGraph (flame.cc:1984-2162) for a slot that already holds
 * the previous graph: carried vertices keep x, w, x_bar, w_bar, x_prev, w_prev, carried edges keep q,
 * everything new gets the reference's defaults, alpha = 1/||u_i - u_j|| and beta = 1 are recomputed
 * for every edge (:2089-2104). Equivalent, bit for bit, to flame_b200_graph_download + the same edits
 * on the host + flame_b200_graph_upload. Asynchronous. */
int flame_b200_graph_resync(flame_b200_ctx* ctx, int slot, const flame_b200_graph_sync* sync);

enum {
  FLAME_B200_PD_AUTO = 0,     /* resident if the graph fits one cluster's shared memory */
  FLAME_B200_PD_STREAM = 1,   /* state in HBM/L2, two kernels per iteration, any size    */
  FLAME_B200_PD_RESIDENT = 2  /* state in (distributed) shared memory for all iterations */
};

/* Replaces n_iters consecutive calls of nltgv2_l1_graph_regularizer::step(params, &graph)
 * (h:134, cc:33-49) on slots [slot0, slot0 + n_slots). Asynchronous. */
int flame_b200_pd_run(flame_b200_ctx* ctx, int slot0, int n_slots,
                      const flame_b200_pd_params* params, int n_iters, int mode);

/* Host-only diagnostic (no device, no context): the vertex order and launch plan that a graph upload computes
 * for these vertices / edges - the only per-graph host computation of the library (everything else of the
 * layout is built by kernels). busy_sms: optional table [17], entry C = SMs kept busy by clusters of C CTAs
 * (what flame_b200_ctx_create measures with cudaOccupancyMaxActiveClusters; NULL = take the first size that
 * fits). vperm [V]: caller vertex at every internal position (Z-curve chunks, each sorted by degree and owned-edge
 * count); koff [n_chunks * (max_degree + 1)] and chunk_base [n_chunks + 1]: incidence k of the i-th vertex of
 * chunk c lives at slot chunk_base[c] + koff[c][k] + i. Any of the three may be NULL. */
typedef struct {
  int32_t chunk_width, n_chunks, max_degree;
  int32_t cluster;          /* CTAs per frame cluster of the resident solver; 0 = the streaming solver */
  int32_t variant, threads; /* kernel build and block size */
  int32_t max_incidences, max_owned_edges;
  int64_t shared_bytes;     /* dynamic shared memory per CTA */
} flame_b200_graph_layout;
int flame_b200_graph_plan(int32_t V, const float* pos, int32_t E, const int32_t* edges, int preference,
                          const int32_t* busy_sms, flame_b200_graph_layout* out, int32_t* vperm,
                          int32_t koff_capacity, int32_t* koff, int32_t* chunk_base);

/* Number of frames like the one in `slot` that the resident solver runs CONCURRENTLY: one frame is one
 * thread-block cluster, and the GPU holds a limited number of clusters at a time (7 of 16 CTAs on a
 * B200: cudaOccupancyMaxActiveClusters). pd_run over more slots proceeds in waves of this size; batches
 * that are a multiple of it waste no wave. 0 = the slot's graph takes the streaming path. */
int flame_b200_pd_wave(flame_b200_ctx* ctx, int slot, int32_t* frames);

/* Replaces smoothnessCost / dataCost (h:139,144; cc:51-85). exact_order != 0 reproduces the
 * reference's sequential fp32 accumulation order bit-for-bit (slower); 0 uses a tree reduction
 * in fp64. Synchronises the stream. */
int flame_b200_pd_cost(flame_b200_ctx* ctx, int slot, const flame_b200_pd_params* params,
                       int exact_order, float* smoothness_cost, float* data_cost);

/* ---- dense projection: utils::interpolateMesh (utils/image_utils.h:646-651, .cc:373-396) over
 *      utils::DrawShadedTriangleBarycentric (utils/rasterization.cc:164-245) -------------- */

/* tris: 3T vertex indices in Triangle's (clockwise) order, as utils::Delaunay::triangles().
 * values: V floats, NULL -> the slot's current x scaled by `scale` (vtx_idepths_ = x *
 * graph_scale_, flame.cc:376). vtx_valid / tri_valid: nullable (all true).
 * out_host: rows x cols floats. fill_nan != 0 pre-fills the image with quiet NaN (as
 * Flame::update does, flame.cc:412); otherwise out_host is read as the pre-initialised image.
 * Synchronises the stream. */
int flame_b200_interpolate_mesh(flame_b200_ctx* ctx, int slot, int32_t T, const int32_t* tris,
                                int32_t V, const float* pos, const float* values, float scale,
                                const uint8_t* vtx_valid, const uint8_t* tri_valid, int rows,
                                int cols, int fill_nan, float* out_host);

/* Device-resident variant for pipelines and for timing without PCIe: the mesh (tris, validity)
 * is bound to the slot once, the image stays in HBM; `values` come from the slot's x. */
int flame_b200_mesh_bind(flame_b200_ctx* ctx, int slot, int32_t T, const int32_t* tris,
                         const uint8_t* vtx_valid, const uint8_t* tri_valid, int rows, int cols);
int flame_b200_mesh_project(flame_b200_ctx* ctx, int slot0, int n_slots, float scale); /* async */
/* Asynchronous device->host copy of the slot's dense map into PINNED or pageable host memory;
 * complete after flame_b200_ctx_sync(). */
int flame_b200_mesh_download(flame_b200_ctx* ctx, int slot, float* out_host);

/* Device pointers of a slot's outputs, for consumers that stay on the GPU - the batched-frame case of
 * one frame per GPU, the dense maps and vertex idepths gathered to one consumer with NCCL over
 * NVLink (flame_b200/shard.py gather_frames, bench.py --config E) without a trip through host memory.
 * dense_map: rows x cols f32 of the last flame_b200_mesh_project; values: V f32 = x * scale of that call, in
 * the caller's vertex order (vtx_idepths_, flame.cc:376). Valid until the slot's next mesh_bind / upload;
 * reads must be ordered after the producing work (flame_b200_ctx_stream / flame_b200_ctx_join). */
int flame_b200_slot_device_ptrs(flame_b200_ctx* ctx, int slot, void** dense_map, void** values, int32_t* rows,
                                int32_t* cols, int32_t* V);

/* The projection fused with the exchange step of the batched-frame case: after
 * flame_b200_mesh_redirect(ctx, slot, map_dst, values_dst) every flame_b200_mesh_project of the slot stores
 * the dense map DIRECTLY into map_dst (rows x cols f32) and copies the vertex idepths to values_dst (V f32) -
 * device memory that may belong to ANOTHER GPU of the box (peer stores over NVLink: the consumer's gather
 * buffer is filled tile by tile while the projection runs; no separate collective moves the 8.3 MB map).
 * NULL restores the slot's own buffer; a new mesh_bind ends the redirection. While redirected, the slot's own
 * map (mesh_download, mesh_coverage) is not updated.
 * flame_b200_ipc_export / _open share a device allocation between the per-GPU processes of one box (CUDA
 * IPC; peer access is enabled on open); flame_b200_device_alloc gives memory that can be exported. The
 * consumer orders its reads after the producers' streams (e.g. a tiny NCCL all-reduce on those streams). */
int flame_b200_mesh_redirect(flame_b200_ctx* ctx, int slot, void* dense_map_dst, void* values_dst);
int flame_b200_device_alloc(int device, size_t bytes, void** out);
int flame_b200_device_free(void* p);
int flame_b200_ipc_export(void* dev_ptr, uint8_t handle[64]);
int flame_b200_ipc_open(int device, const uint8_t handle[64], void** out);
int flame_b200_ipc_close(void* p);

/* ---- display filters and coverage ----------------------------------------
 * Flame::update filters triangles "for display" (flame.cc:389-407) and getFilteredInverseDepthMap
 * (flame.h:217-228) projects the mesh again honouring them. */
typedef struct { /* params.h:70-85 */
  int32_t do_oblique_triangle_filter;
  float oblique_normal_thresh;      /* max angle [rad] between the triangle normal and the viewing ray */
  float oblique_idepth_diff_factor; /* (max - min) / max of the corner idepths                         */
  float oblique_idepth_diff_abs;    /* max - min                                                       */
  int32_t do_edge_length_filter;
  float edge_length_thresh;         /* fraction of the image width                                     */
  int32_t do_idepth_triangle_filter;
  float min_triangle_idepth;
} flame_b200_tri_filter_params;
void flame_b200_tri_filter_default_params(flame_b200_tri_filter_params* p);

/* Replaces obliqueTriangleFilter + edgeLengthFilter + idepthTriangleFilter (flame.cc:2207-2356) as combined
 * at flame.cc:389-407, on the slot's bound mesh with vtx_idepths = x * scale of the slot's graph (flame.cc:376).
 * tri_valid_out (T bytes, nullable) receives tri_validity_. apply != 0 installs the result on the device as the
 * mesh's triangle validity (in addition to what mesh_bind was given) for the following projections - i.e.
 * mesh_project then produces getFilteredInverseDepthMap's image; the next mesh_bind removes it.
 * The angle test uses the GPU's acosf: a triangle whose angle is within a few ulp of the threshold can
 * differ from libm's decision; everything else is exact. Synchronises if tri_valid_out != NULL. */
int flame_b200_mesh_filter(flame_b200_ctx* ctx, int slot, const flame_b200_tri_filter_params* params,
                           const float Kinv[9], float scale, uint8_t* tri_valid_out, int apply);

/* Diagnostic: the angle [rad] between the inward normal and the viewing ray of every triangle, as the
 * oblique filter of flame_b200_mesh_filter computes it (flame.cc:2255). Everything up to the acosf argument
 * is bit-identical to the reference's arithmetic; acosf itself is the GPU's (<= 2 ulp from libm's), which is
 * why a filter decision can differ for a triangle whose angle lies within ~1e-6 rad of the threshold.
 * angles_out: T floats. Synchronises. */
int flame_b200_mesh_filter_angles(flame_b200_ctx* ctx, int slot, const float Kinv[9], float scale,
                                  float* angles_out);

/* Replaces Flame::getVertexNormals (flame.cc:2554-2640, called at flame.cc:385): per vertex the running,
 * re-normalised average of the outward normals of its triangles in triangle order, with
 * vtx_idepths = x * scale of the slot's graph; triangles with a non-positive corner idepth are skipped.
 * normals_out: 3V floats (x, y, z per vertex, caller order). Synchronises. */
int flame_b200_mesh_normals(flame_b200_ctx* ctx, int slot, const float Kinv[9], float scale, float* normals_out);

/* The coverage statistic of flame.cc:434-443: number of non-NaN pixels of the slot's dense map.
 * Synchronises. */
int flame_b200_mesh_coverage(flame_b200_ctx* ctx, int slot, int64_t* covered);

/* ---- the per-frame tail of Flame::update in one call (flame.cc:312-415): fresh graph from the
 *      features (syncGraph's init rule + edge weights), n_iters x step(), extraction of the smoothed
 *      vertex idepths (x * scale) and plane parameters, interpolateMesh into a NaN-initialised map.
 *      Host buffers in, host buffers out (any output pointer may be NULL); out_map should be
 *      page-locked (flame_b200_host_alloc) for full PCIe speed. Synchronises. */
int flame_b200_process_frame(flame_b200_ctx* ctx, int slot, int32_t V, const float* pos,
                             const float* idepth, const float* weight, int32_t E,
                             const int32_t* edges, int32_t T, const int32_t* tris, int rows, int cols,
                             const flame_b200_pd_params* params, int n_iters, float scale,
                             float* out_map, float* out_x, float* out_w1, float* out_w2);

/* The same call split in two, for a pipeline of frames driven by ONE host thread: _submit copies the
 * inputs into the slot's pinned staging area (the caller's input buffers are free on return) and only
 * ENQUEUES the upload, the solve, the projection and the downloads on the slot's own CUDA stream; frames
 * submitted to different slots of a context overlap (upload of frame k+1, solve of k, download of k-1, and
 * as many concurrent solves as clusters fit the device). out_map (nullable; page-locked memory for a truly
 * asynchronous copy) is written by the device->host copy and must stay valid until _wait returns.
 * _wait blocks until the slot's frame is complete and copies x, w1, w2 (V floats each, any may be NULL) out.
 * A slot holds one frame in flight: submit - wait - submit ...; until _wait, no other call may use the slot.
 * The frame's graph and dense map stay in the slot afterwards (flame_b200_graph_download, mesh_download,
 * mesh_filter ... work as after flame_b200_process_frame). */
int flame_b200_process_frame_submit(flame_b200_ctx* ctx, int slot, int32_t V, const float* pos,
                                    const float* idepth, const float* weight, int32_t E,
                                    const int32_t* edges, int32_t T, const int32_t* tris, int rows, int cols,
                                    const flame_b200_pd_params* params, int n_iters, float scale,
                                    float* out_map);
int flame_b200_process_frame_wait(flame_b200_ctx* ctx, int slot, float* out_x, float* out_w1, float* out_w2);
/* edges == NULL and tris == NULL in flame_b200_process_frame_submit: the call triangulates the vertices itself
 * (flame_b200_delaunay below, on the calling thread, while the device works on the frames submitted before) -
 * the place utils::Delaunay::triangulate has in Flame::update (flame.cc:312-330). E and T are ignored. The
 * arrays it used stay with the slot until its next submit; this returns them (counts always; tris / edges may
 * be NULL; FLAME_B200_ENOMEM if a capacity is too small, 2V + 1 triangles and 3V + 1 edges always suffice). */
int flame_b200_process_frame_topology(flame_b200_ctx* ctx, int slot, int32_t tri_capacity, int32_t* n_tris,
                                      int32_t* tris, int32_t edge_capacity, int32_t* n_edges, int32_t* edges);

/* Host-side accounting of the frame pipeline since context creation / the last reset: where the calling
 * thread's time went (vertex ordering + launch plan, staging copies, enqueueing, waiting) and the bytes
 * queued over PCIe. */
typedef struct {
  uint64_t frames, h2d_bytes, d2h_bytes;
  double prep_seconds;    /* vertex order + launch plan (the only per-graph host computation)     */
  double stage_seconds;   /* process_frame_submit up to the solver launch: prep + staging memcpys + upload calls */
  double enqueue_seconds; /* solver / projection / download launches                               */
  double wait_seconds;    /* blocked in process_frame_wait                                         */
  double triangulation_seconds; /* flame_b200_delaunay inside process_frame_submit (edges == tris == NULL) */
} flame_b200_stats;
int flame_b200_ctx_stats(flame_b200_ctx* ctx, flame_b200_stats* out, int reset);

/* ---- triangulation for the frame pipeline ------------------------------------------
 * Replaces utils::Delaunay::triangulate (utils/delaunay.cc:31-77 over external/triangle) where the caller does
 * not need Triangle's exact index arrays: a host-side (pure CPU, thread-safe, no context) divide-and-conquer
 * Delaunay triangulation with exact integer predicates on positions rounded to 1/256 px. A host thread runs it
 * for frame k+1 while the device works on frame k (bench.py e2e "incl_triangulation").
 * Output in Triangle's conventions: tris = 3 vertex indices per triangle, positively oriented in image
 * coordinates (Triangle's order); edges = every undirected edge once (source, target). The arrays are A
 * Delaunay triangulation of the points, NOT Triangle's arrays: cocircular quadruples may be split along the
 * other diagonal and the edge order differs - callers that need index-identical output keep Triangle and pass
 * its arrays, which every entry point of this library accepts unchanged.
 * pos: 2V floats, |coordinate| < 8192. Coincident points (after rounding) are represented by the first of
 * them; the others get no triangle. tris / edges may be NULL to query the counts. Returns FLAME_B200_ENOMEM
 * (with the counts set) when a capacity is too small: 2V triangles and 3V edges always suffice. */
int flame_b200_delaunay(int32_t V, const float* pos, int32_t tri_capacity, int32_t* n_tris, int32_t* tris,
                        int32_t edge_capacity, int32_t* n_edges, int32_t* edges);

/* ---- epipolar search + measurement + fusion: Flame::updateFeatureIDepths / trackFeature
 *      (flame.cc:1280-1752) over stereo::inverse_depth_filter::{predict,getSearchRegion,search,
 *      update} (stereo/inverse_depth_filter.h:73-107), stereo::line_stereo::match
 *      (stereo/line_stereo.h:75-385), stereo::InverseDepthMeasModel::idepth
 *      (stereo/inverse_depth_meas_model.cc:49-156), stereo::EpipolarGeometry<float>
 *      (stereo/epipolar_geometry.h:44-428) --------------------------------------------------- */

typedef struct { /* stereo::line_stereo::Params, line_stereo.h:47-59 */
  float max_cost;
  int32_t do_subpixel;
  float sample_dist;
  float second_best_factor;
} flame_b200_ls_params;

typedef struct { /* stereo::inverse_depth_filter::Params, inverse_depth_filter.h:48-68 */
  int32_t win_size;
  float search_sigma;
  float min_grad_mag;
  float idepth_min, idepth_max;
  float epilength_min, epilength_max;
  float process_var_factor;
  float process_fail_var_factor;
  flame_b200_ls_params sparams;
} flame_b200_filter_params;

typedef struct { /* stereo::InverseDepthMeasModel::Params, inverse_depth_meas_model.h:44-52 */
  int32_t win_size;
  float pixel_var;
  float epipolar_line_var;
} flame_b200_meas_params;

typedef struct { /* the fields of flame::Params (params.h:36-143) the per-feature driver reads */
  flame_b200_filter_params fparams;
  flame_b200_meas_params zparams;
  float min_baseline;
  float rescale_factor_min, rescale_factor_max;
  float idepth_var_max;
  int32_t max_dropouts;
  float outlier_sigma_thresh;
  int32_t do_meas_fusion;
  int32_t do_letterbox;
} flame_b200_track_params;
void flame_b200_track_default_params(flame_b200_track_params* p);

typedef struct { /* flame::FeatureWithIDepth, flame.h:88-99 */
  uint32_t id;
  uint32_t frame_id;      /* frame slot of the parent pose-frame */
  float x, y;
  float idepth_mu, idepth_var;
  int32_t valid;
  uint32_t num_updates;
  uint32_t num_dropouts;
  int32_t search_status;  /* inverse_depth_filter::Status, inverse_depth_filter.h:38-43 */
} flame_b200_feature;

typedef struct { /* per-feature diagnostics (what the reference keeps in locals) */
  float flow_x, flow_y;
  float mu_meas, var_meas;
  int32_t stage; /* 0 skipped(baseline) 1 predict-fail 2 re-parented 3 re-parent-fail 4 region-fail
                    5 out-of-bounds 6 search-fail 7 sense-fail 8 fuse-fail 9 updated */
} flame_b200_track_diag;

/* Camera intrinsics for the search (Flame ctor, flame.h:123-125). Row-major 3x3. */
int flame_b200_set_camera(flame_b200_ctx* ctx, int width, int height, const float K[9],
                          const float Kinv[9]);

/* Replaces utils::Frame's padded level-0 buffers (utils/frame.cc:54-63; img_pad REFLECT_101,
 * gradients zero-padded, border `pad`) for frame slot `fslot` (0 <= fslot < 64).
 * pose = camera-to-world (qw,qx,qy,qz,tx,ty,tz) as utils::Frame::pose. Asynchronous. */
int flame_b200_frame_upload(flame_b200_ctx* ctx, int fslot, const uint8_t* img_pad,
                            const float* gradx_pad, const float* grady_pad, int pad,
                            const float pose[7]);

/* SURVEY §8(f) rank 3 — frame preparation on the device: the same frame slot filled from the RAW
 * (unpadded, height x width) u8 image; the padded image (cv::copyMakeBorder REFLECT_101) and the
 * zero-padded central-difference gradients (utils::getCentralGradient<uint8_t,float>,
 * utils/image_utils.cc:36-108) that utils::Frame::create builds on the host (utils/frame.cc:54-63,
 * level 0) are computed by a kernel: 9x less host->device traffic per frame. Asynchronous. */
int flame_b200_frame_upload_raw(flame_b200_ctx* ctx, int fslot, const uint8_t* img, int pad,
                                const float pose[7]);
/* Reads a frame slot's padded buffers back (any pointer may be NULL). Synchronises the stream. */
int flame_b200_frame_download(flame_b200_ctx* ctx, int fslot, uint8_t* img_pad, float* gradx_pad,
                              float* grady_pad);

/* Replaces the loop body of Flame::updateFeatureIDepths (flame.cc:1308-1496) for n features:
 * feats is updated in place (host memory), diag (nullable) receives per-feature diagnostics.
 * fnew = frame slot of the new frame, curr_pf = frame slot of the current pose-frame.
 * Returns the number of successful updates in *n_updates. Synchronises the stream.
 * The SSD walk of a feature is held in shared memory by one warp: parameter sets that allow more than 124
 * positions per feature (max(epilength_max, epilength_min) / sample_dist + 4 > 128; the defaults give 36)
 * are rejected with FLAME_B200_EINVAL. */
int flame_b200_update_feature_idepths(flame_b200_ctx* ctx, const flame_b200_track_params* params,
                                      int fnew, int curr_pf, flame_b200_feature* feats,
                                      int32_t n_feats, flame_b200_track_diag* diag,
                                      int32_t* n_updates);

/* SURVEY §8(f) rank 3 - replaces the grid detector of Flame::detectFeatures (flame.cc:1192-1277): on a grid
 * of detection_win_size^2 cells, the pixel of the reference frame (frame slot fref) with the largest squared
 * gradient along the epipolar direction of the (previous frame fprev <- reference) geometry, among pixels
 * whose gradient magnitude and epipolar gradient both exceed min_grad_mag, in cells that hold none of the
 * n_curr current features. features_xy receives up to max_features (x, y) pairs in row-major cell order
 * (the order of the reference's push_back), *n_features their total number. Synchronises. */
typedef struct {
  float rescale_factor_max;   /* params.h:59 */
  int32_t win_size;           /* fparams.win_size */
  int32_t do_letterbox;       /* params.h:49 */
  float min_grad_mag;         /* params.h:39 */
  int32_t detection_win_size; /* params.h:48 */
} flame_b200_detect_params;
int flame_b200_detect_features(flame_b200_ctx* ctx, const flame_b200_detect_params* params, int fref, int fprev,
                               const float* curr_xy, int32_t n_curr, float* features_xy, int32_t max_features,
                               int32_t* n_features);

/* The same with the feature set RESIDENT on the device (features stay in HBM from frame to frame, as the
 * graph does with flame_b200_graph_resync):
 *   _upload   replaces the resident set with n host records (validated against the uploaded frame slots);
 *   _track    runs the loop body of Flame::updateFeatureIDepths on the resident set. Asynchronous, and
 *             out-of-place: the results, diagnostics and update count are written next to the set, so a
 *             repeated call recomputes the same results from the same inputs;
 *   _download copies the results (or, before any _track, the set itself) to host memory; any pointer may be
 *             NULL. commit != 0 makes the results the resident set: the next _track continues from them,
 *             as the reference's in-place update does (flame.cc:1308-1496). Synchronises. */
int flame_b200_features_upload(flame_b200_ctx* ctx, const flame_b200_feature* feats, int32_t n_feats);
int flame_b200_features_track(flame_b200_ctx* ctx, const flame_b200_track_params* params, int fnew, int curr_pf);
int flame_b200_features_download(flame_b200_ctx* ctx, flame_b200_feature* feats, flame_b200_track_diag* diag,
                                 int32_t* n_updates, int commit);

#ifdef __cplusplus
}
#endif
#endif /* FLAME_B200_H_ */
