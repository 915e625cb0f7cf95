/*
 * flame_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, single-threaded restatement of the reference FLaME hot path
 * (robustrobotics/flame @ fea7ceea). Every function cites the reference
 * file:line it follows (paths relative to the reference checkout).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the CPU arm. The product
 * (flame_b200/csrc, libflame_b200.so) never links or calls it.
 *
 * Parity pinning (see DESIGN.md "Oracle"):
 *   - rasteriser, bilinearInterp, clipLineLiangBarsky, EpipolarGeometry: pinned by the
 *     reference's own known-answer tests (tests/test_oracle_kat.py) and, in this
 *     container, by the reference sources compiled verbatim (oracle/_ref).
 *   - PD solver, line stereo, filter, measurement model: the reference has no test
 *     for them; pinned against the reference sources compiled verbatim against
 *     oracle/shim (oracle/_ref/libflame_ref.so) on seeded inputs.
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off (no -march: the reference's Release build
 * emits no FMA).
 */
#ifndef FLAME_ORACLE_H_
#define FLAME_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- PD solver: optimizers/nltgv2_l1_graph_regularizer.{h,cc} ---- */

/* Params, nltgv2_l1_graph_regularizer.h:121-129 (same field order). */
typedef struct {
  float data_factor;
  float step_x;
  float step_q;
  float theta;
  float x_min;
  float x_max;
} oracle_pd_params;

/* Flat SoA view of Graph (VertexData h:74-90, EdgeData h:95-102). Vertex order = caller's
 * order; edge order = caller's order (== Triangle edgelist order for a fresh graph,
 * SURVEY N3); edges[2e] = source, edges[2e+1] = target. */
typedef struct {
  int32_t V, E;
  const float* pos;     /* 2V: x,y interleaved */
  float *x, *w1, *w2;
  float *x_bar, *w1_bar, *w2_bar;
  float *x_prev, *w1_prev, *w2_prev;
  const float *data_term, *data_weight;
  const int32_t* edges; /* 2E */
  const float *alpha, *beta;
  float *q1, *q2, *q3;
} oracle_graph;

void oracle_pd_default_params(oracle_pd_params* p);
void oracle_pd_step(const oracle_pd_params* p, oracle_graph* g);             /* cc:33-49 */
void oracle_pd_run(const oracle_pd_params* p, oracle_graph* g, int n_iters); /* step x n */
float oracle_pd_smoothness_cost(const oracle_pd_params* p, const oracle_graph* g); /* cc:51-71 */
float oracle_pd_data_cost(const oracle_pd_params* p, const oracle_graph* g);       /* cc:73-85 */
/* flame.cc:2084-2105: alpha = 1/||u_i-u_j||, beta = 1. */
void oracle_edge_weights(int32_t E, const int32_t* edges, const float* pos,
                         float* alpha, float* beta);

/* ---- dense projection: utils/rasterization.cc:164-245, utils/image_utils.cc:373-396 ---- */

/* img: rows x cols float, continuous, pre-initialised by the caller (NaN in Flame::update,
 * flame.cc:412). Writes that would fall outside the buffer are skipped (the reference would
 * write out of bounds there). */
void oracle_draw_shaded_triangle_barycentric(int p1x, int p1y, int p2x, int p2y, int p3x,
                                             int p3y, float v1, float v2, float v3,
                                             float* img, int rows, int cols);
void oracle_interpolate_mesh(int32_t T, const int32_t* tris, int32_t V, const float* pos,
                             const float* values, const uint8_t* vtx_valid,
                             const uint8_t* tri_valid, float* img, int rows, int cols);

/* ---- display filters (flame.cc:389-407) and the coverage statistic (flame.cc:434-443) ----
 * Flame::obliqueTriangleFilter / edgeLengthFilter / idepthTriangleFilter are static members inside flame.cc,
 * which as a whole cannot be compiled here (OpenCV/Boost/Sophus). PINNED on the functions themselves:
 * oracle/ref_capi/slice_flame.py slices them verbatim out of flame.cc, oracle/ref_capi/ref_filters.cc compiles
 * the slices against oracle/shim, and tests/test_oracle_kat.py asserts this restatement equals them. What
 * stays restated is Eigen 3.2's arithmetic inside the shim (3x3 * 3x1 as ((a0 b0 + a1 b1) + a2 b2),
 * normalize() = v / sqrt(squaredNorm), matrix / scalar as a true division), as for the epipolar geometry. */
typedef struct { /* params.h:70-85 */
  int32_t do_oblique_triangle_filter;
  float oblique_normal_thresh, oblique_idepth_diff_factor, oblique_idepth_diff_abs;
  int32_t do_edge_length_filter;
  float edge_length_thresh;
  int32_t do_idepth_triangle_filter;
  float min_triangle_idepth;
} oracle_tri_filter_params;
void oracle_tri_filter_default_params(oracle_tri_filter_params* p);
/* validity[T] (1 = keep); angle[T] (nullable) receives fabs(acos(ray . normal)) of the oblique test. */
void oracle_triangle_filters(const oracle_tri_filter_params* p, const float Kinv[9], int width, int32_t T,
                             const int32_t* tris, const float* pos, const float* idepths,
                             uint8_t* validity, float* angle);
int64_t oracle_coverage(int rows, int cols, const float* img);
/* Flame::getVertexNormals (flame.cc:2554-2640), pinned the same way. normals: 3V floats. */
void oracle_vertex_normals(const float Kinv[9], int32_t V, const float* pos, const float* idepths, int32_t T,
                           const int32_t* tris, float* normals);

/* ---- image helpers: utils/image_utils.h:199-272, utils/image_utils.cc:269-371 ---- */
float oracle_bilinear_u8(int rows, int cols, int step_bytes, const uint8_t* data, float x,
                         float y);
float oracle_bilinear_f32(int rows, int cols, int step_bytes, const float* data, float x,
                          float y);
int oracle_clip_line_liang_barsky(float xmin, float xmax, float ymin, float ymax, float x0,
                                  float y0, float x1, float y1, float* x0c, float* y0c,
                                  float* x1c, float* y1c);

/* ---- frame preparation: utils/image_utils.cc:36-108 (getCentralGradient<uint8_t,float>), utils/frame.cc:54-63 ---- */
void oracle_central_gradient_u8(int width, int height, const uint8_t* img, float* gradx, float* grady);
/* Level-0 buffers of utils::Frame::create: img_pad = copyMakeBorder(img, REFLECT_101), gradients of
 * the unpadded image zero-padded (BORDER_CONSTANT), all (height+2*border) x (width+2*border). */
void oracle_frame_prep(int width, int height, int border, const uint8_t* img, uint8_t* img_pad,
                       float* gradx_pad, float* grady_pad);

/* ---- stereo/epipolar_geometry.h ---- */
typedef struct {
  float K[9], Kinv[9];           /* row-major 3x3 */
  float q[4];                    /* q_ref_to_cmp as (w,x,y,z) */
  float t[3];                    /* t_ref_to_cmp */
  float t_cmp_to_ref[3];
  float KRKinv[9];
  float KRKinv3[3];
  float Kt[3];
  float epipole[2];
} oracle_epigeo;

void oracle_epigeo_init(oracle_epigeo* g, const float K[9], const float Kinv[9]);
void oracle_epigeo_load(oracle_epigeo* g, const float q_wxyz[4], const float t[3]); /* :86-103 */
void oracle_epigeo_project(const oracle_epigeo* g, float ux, float uy, float idepth,
                           float* cx, float* cy);                                    /* :127-143 */
void oracle_epigeo_project_idepth(const oracle_epigeo* g, float ux, float uy, float idepth,
                                  float* cx, float* cy, float* new_idepth);          /* :153-180 */
void oracle_epigeo_max_depth_projection(const oracle_epigeo* g, float ux, float uy,
                                        float* ix, float* iy);                       /* :191-201 */
void oracle_epigeo_min_depth_projection(const oracle_epigeo* g, float ux, float uy,
                                        float* mx, float* my);                       /* :237-263 */
void oracle_epigeo_epiline(const oracle_epigeo* g, float ux, float uy, float* ix, float* iy,
                           float* ex, float* ey);                                    /* :282-300 */
void oracle_epigeo_reference_epiline(const oracle_epigeo* g, float ux, float uy, float* ex,
                                     float* ey);                                     /* :311-331 */
float oracle_epigeo_disparity(const oracle_epigeo* g, float ux, float uy, float cx, float cy,
                              float* ix, float* iy, float* ex, float* ey);           /* :342-347 */
float oracle_epigeo_disparity_to_depth(const oracle_epigeo* g, float ux, float uy, float ix,
                                       float iy, float ex, float ey, float disp);    /* :365-380 */
float oracle_epigeo_disparity_to_idepth(const oracle_epigeo* g, float ux, float uy, float ix,
                                        float iy, float ex, float ey, float disp);   /* :392-407 */
/* static project(K,q,t,p), :115-120 */
void oracle_epigeo_project_world(const float K[9], const float q_wxyz[4], const float t[3],
                                 const float p[3], float* ux, float* uy);

/* SE3 helpers standing in for Sophus::SE3f compose/inverse (flame.cc:1314-1317). Pose =
 * (q w,x,y,z ; t x,y,z). out = inverse(a) * b. */
void oracle_se3_inv_mul(const float a[7], const float b[7], float out[7]);

/* ---- stereo/inverse_depth_filter.{h,cc}, stereo/line_stereo.h ---- */
enum { ORACLE_SEARCH_SUCCESS = 0, ORACLE_FAIL_REF_PATCH_GRADIENT = 1,
       ORACLE_FAIL_AMBIGUOUS_MATCH = 2, ORACLE_FAIL_MAX_COST = 3 }; /* inverse_depth_filter.h:38-43 */
enum { ORACLE_LS_SUCCESS = 0, ORACLE_LS_FAIL_AMBIGUOUS_MATCH = 1,
       ORACLE_LS_FAIL_MAX_COST = 2 };                               /* line_stereo.h:38-42 */

typedef struct { /* line_stereo.h:47-59 */
  float max_cost;
  int do_subpixel;
  float sample_dist;
  float second_best_factor;
} oracle_ls_params;

typedef struct { /* inverse_depth_filter.h:48-68 */
  int win_size;
  float search_sigma;
  float min_grad_mag;
  float idepth_min, idepth_max;
  float epilength_min, epilength_max;
  float process_var_factor;
  float process_fail_var_factor;
  oracle_ls_params sparams;
} oracle_filter_params;

typedef struct { /* inverse_depth_meas_model.h:44-52 */
  int win_size;
  float pixel_var;
  float epipolar_line_var;
} oracle_meas_params;

void oracle_filter_default_params(oracle_filter_params* p);
void oracle_meas_default_params(oracle_meas_params* p);

int oracle_line_stereo_match(float rescale_factor, const float ref_patch[5], const uint8_t* img,
                             int rows, int cols, int step, float sx, float sy, float ex,
                             float ey, float* ucx, float* ucy, float* residual,
                             const oracle_ls_params* p);                  /* line_stereo.h:75-385 */
int oracle_filter_predict(const oracle_epigeo* g, float process_var_factor, float ux, float uy,
                          float mu, float var, float* cx, float* cy, float* mu_pred,
                          float* var_pred);                               /* filter.cc:35-63 */
int oracle_filter_get_search_region(const oracle_filter_params* p, const oracle_epigeo* g,
                                    int width, int height, float ux, float uy, float mu_prior,
                                    float var_prior, float* sx, float* sy, float* ex, float* ey,
                                    float* epx, float* epy);              /* filter.cc:65-182 */
int oracle_filter_search(const oracle_filter_params* p, const oracle_epigeo* g,
                         float rescale_factor, const uint8_t* img_ref, const uint8_t* img_cmp,
                         int rows, int cols, int step, float urx, float ury, float sx, float sy,
                         float ex, float ey, float* ucx, float* ucy);     /* filter.cc:184-266 */
int oracle_filter_update(float mu_pred, float var_pred, float mu_meas, float var_meas,
                         float* mu_post, float* var_post, float outlier_sigma_thresh);
                                                                          /* filter.cc:268-305 */
/* inverse_depth_meas_model.cc:49-156. gradx/grady: PADDED cmp gradients (rows x cols floats). */
int oracle_meas_idepth(const oracle_meas_params* p, const oracle_epigeo* g, const float* gradx_cmp,
                       const float* grady_cmp, int rows, int cols, float urx, float ury, float ucx,
                       float ucy, float* mu, float* var);

/* ---- flame.cc:1280-1752 per-feature driver (updateFeatureIDepths + trackFeature) ---- */
typedef struct { /* FeatureWithIDepth, flame.h:88-99 */
  uint32_t id;
  uint32_t frame_id;     /* index into the frame table below */
  float x, y;
  float idepth_mu, idepth_var;
  int32_t valid;
  uint32_t num_updates;
  uint32_t num_dropouts;
  int32_t search_status;
} oracle_feature;

typedef struct { /* one entry per pose-frame the features may be parented to */
  float pose[7];          /* q(w,x,y,z), t(x,y,z): camera-to-world */
  const uint8_t* img_pad; /* (H+2*pad) x (W+2*pad) u8 */
  const float* gradx_pad; /* same size, f32 */
  const float* grady_pad;
} oracle_frame;

typedef struct { /* the subset of flame::Params (params.h:36-143) the driver reads */
  oracle_filter_params fparams;
  oracle_meas_params zparams;
  float min_baseline;
  float rescale_factor_min, rescale_factor_max;
  float idepth_var_max;
  int max_dropouts;
  float outlier_sigma_thresh;
  int do_meas_fusion;
  int do_letterbox;
} oracle_track_params;

typedef struct { /* per-feature diagnostics, for element-wise parity */
  float flow_x, flow_y;   /* matched pixel (unpadded), valid if tracked */
  float mu_meas, var_meas;
  int32_t stage;          /* 0 skipped(baseline) 1 predict-fail 2 moved/re-parented 3 move-fail
                             4 region-fail 5 oob 6 search-fail 7 sense-fail 8 fuse-fail 9 updated */
} oracle_track_diag;

void oracle_track_default_params(oracle_track_params* p);
/* Returns the number of successful updates. frames[curr_pf] is the current pose-frame (re-parent
 * target), frames[fnew] the new frame. width/height are the UNPADDED image size; pad the border. */
int oracle_update_feature_idepths(const oracle_track_params* p, const float K[9],
                                  const float Kinv[9], int width, int height, int pad,
                                  const oracle_frame* frames, int n_frames, int fnew, int curr_pf,
                                  oracle_feature* feats, int n_feats, oracle_track_diag* diag);

/* ---- Flame::projectGraph (flame.cc:1862-1937): per-vertex projection + removal marks ----
 * pos (2V) and x (V) are updated in place as the reference does (vtx.pos = u_new, vtx.x = idepth_new /
 * graph_scale); idepth_out[V] (nullable) = idepth_new; remove[V] = 1 where the vertex lands in vtx_to_remove.
 * gradx / grady: the new frame's UNPADDED level-0 gradients (rows x cols), only read with the gradient check.
 * Pinned on the function itself sliced out of flame.cc (oracle/ref_capi/ref_project.cc). */
typedef struct {
  float rescale_factor_max;
  int32_t win_size, do_letterbox, do_grad_check_after_projection;
  float min_grad_mag;
} oracle_project_rules;
void oracle_project_graph(const float K[9], const float Kinv[9], const float q_wxyz[4], const float t[3],
                          int width, int height, const oracle_project_rules* rules, float graph_scale, int32_t V,
                          float* pos, float* x, const float* gradx, const float* grady, float* idepth_out,
                          uint8_t* remove);

/* ---- Flame::detectFeatures, the single-pass grid detector (flame.cc:1192-1277) ----
 * gradx / grady: the reference frame's UNPADDED gradients; pose_ref / pose_prev: camera-to-world (qw qx qy qz
 * tx ty tz). out_xy: up to max_out (x, y) pairs in row-major cell order; returns the number of detections.
 * Pinned on the function itself sliced out of flame.cc (oracle/ref_capi/ref_detect.cc). */
typedef struct {
  float rescale_factor_max;
  int32_t win_size, do_letterbox;
  float min_grad_mag;
  int32_t detection_win_size;
} oracle_detect_params;
int32_t oracle_detect_features(const float K[9], const float Kinv[9], const float pose_ref[7],
                               const float pose_prev[7], int width, int height, const oracle_detect_params* p,
                               const float* gradx, const float* grady, const float* curr_xy, int32_t n_curr,
                               float* out_xy, int32_t max_out);

#ifdef __cplusplus
}
#endif
#endif  /* FLAME_ORACLE_H_ */
