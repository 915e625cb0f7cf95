/*
 * flame_oracle.c — CPU ORACLE (test infrastructure, NOT product code). See flame_oracle.h.
 *
 * Plain-C restatement of the reference's hot path. Arithmetic is written in the same
 * association order as the reference's C++ expressions; compile with -ffp-contract=off.
 * Citations are reference paths (src/flame/...) with line numbers.
 */
#include "flame_oracle.h"

#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

/* utils/image_utils.h:91-94  fast_abs: (r > 0) ? r : -r */
static inline float fast_absf(float r) { return (r > 0) ? r : -r; }

/* ====================================================================================
 * PD solver — optimizers/nltgv2_l1_graph_regularizer.{h,cc}
 * ==================================================================================== */

void oracle_pd_default_params(oracle_pd_params* p) {
  /* nltgv2_l1_graph_regularizer.h:121-129 */
  p->data_factor = 0.1f;
  p->step_x = 0.001f;
  p->step_q = 125.0f;
  p->theta = 0.25f;
  p->x_min = 0.0f;
  p->x_max = 10.0f;
}

/* h:171-176  proxNLTGV2Conj(step, q): q / (|q| > 1 ? |q| : 1) */
static inline float prox_nltgv2_conj(float q) {
  float absq = fast_absf(q);
  return q / (absq > 1 ? absq : 1);
}

/* h:179-197  proxL1 */
static inline float prox_l1(float x_min, float x_max, float step_x, float data_weight, float x,
                            float data) {
  float diff = x - data;
  float thresh = step_x * data_weight;
  float new_x = 0.0f;
  if (diff > thresh) {
    new_x = x - thresh;
  } else if (diff < -thresh) {
    new_x = x + thresh;
  } else {
    new_x = data;
  }
  new_x = (new_x < x_min) ? x_min : new_x;
  new_x = (new_x > x_max) ? x_max : new_x;
  return new_x;
}

/* cc:89-114 */
static void dual_step(const oracle_pd_params* p, oracle_graph* g) {
  for (int32_t e = 0; e < g->E; ++e) {
    int32_t ii = g->edges[2 * e], jj = g->edges[2 * e + 1];
    float alpha = g->alpha[e], beta = g->beta[e];
    float pix = g->pos[2 * ii], piy = g->pos[2 * ii + 1];
    float pjx = g->pos[2 * jj], pjy = g->pos[2 * jj + 1];

    float K1x = alpha * (g->x_bar[ii] - g->x_bar[jj]);      /* cc:99 */
    K1x -= alpha * (pix - pjx) * g->w1_bar[ii];             /* cc:100 */
    K1x -= alpha * (piy - pjy) * g->w2_bar[ii];             /* cc:101 */
    g->q1[e] = prox_nltgv2_conj(g->q1[e] + p->step_q * K1x); /* cc:102 */

    float K2x = beta * (g->w1_bar[ii] - g->w1_bar[jj]);     /* cc:105 */
    g->q2[e] = prox_nltgv2_conj(g->q2[e] + p->step_q * K2x);

    float K3x = beta * (g->w2_bar[ii] - g->w2_bar[jj]);     /* cc:109 */
    g->q3[e] = prox_nltgv2_conj(g->q3[e] + p->step_q * K3x);
  }
}

/* cc:116-154 */
static void primal_step(const oracle_pd_params* p, oracle_graph* g) {
  for (int32_t e = 0; e < g->E; ++e) {
    int32_t ii = g->edges[2 * e], jj = g->edges[2 * e + 1];
    float alpha = g->alpha[e], beta = g->beta[e];
    float q1 = g->q1[e], q2 = g->q2[e], q3 = g->q3[e];
    float pix = g->pos[2 * ii], piy = g->pos[2 * ii + 1];
    float pjx = g->pos[2 * jj], pjy = g->pos[2 * jj + 1];

    g->x[ii] -= q1 * p->step_x * alpha;                     /* cc:126 */
    g->x[jj] += q1 * p->step_x * alpha;                     /* cc:127 */
    g->w1[ii] += q1 * p->step_x * alpha * (pix - pjx);      /* cc:129-130 */
    g->w2[ii] += q1 * p->step_x * alpha * (piy - pjy);      /* cc:132-133 */
    g->w1[ii] -= q2 * p->step_x * beta;                     /* cc:136 */
    g->w1[jj] += q2 * p->step_x * beta;                     /* cc:137 */
    g->w2[ii] -= q3 * p->step_x * beta;                     /* cc:140 */
    g->w2[jj] += q3 * p->step_x * beta;                     /* cc:141 */
  }
  for (int32_t v = 0; v < g->V; ++v) {                      /* cc:145-151 */
    g->x[v] = prox_l1(p->x_min, p->x_max, p->step_x, p->data_factor * g->data_weight[v],
                      g->x[v], g->data_term[v]);
  }
}

/* cc:156-174 */
static void extra_gradient_step(const oracle_pd_params* p, oracle_graph* g) {
  for (int32_t v = 0; v < g->V; ++v) {
    float new_x_bar = g->x[v] + p->theta * (g->x[v] - g->x_prev[v]);
    new_x_bar = (new_x_bar < p->x_min) ? p->x_min : new_x_bar;
    new_x_bar = (new_x_bar > p->x_max) ? p->x_max : new_x_bar;
    g->x_bar[v] = new_x_bar;
    g->w1_bar[v] = g->w1[v] + p->theta * (g->w1[v] - g->w1_prev[v]);
    g->w2_bar[v] = g->w2[v] + p->theta * (g->w2[v] - g->w2_prev[v]);
  }
}

/* cc:33-49 */
void oracle_pd_step(const oracle_pd_params* p, oracle_graph* g) {
  for (int32_t v = 0; v < g->V; ++v) {
    g->x_prev[v] = g->x[v];
    g->w1_prev[v] = g->w1[v];
    g->w2_prev[v] = g->w2[v];
  }
  dual_step(p, g);
  primal_step(p, g);
  extra_gradient_step(p, g);
}

void oracle_pd_run(const oracle_pd_params* p, oracle_graph* g, int n_iters) {
  for (int it = 0; it < n_iters; ++it) oracle_pd_step(p, g);
}

/* cc:51-71 */
float oracle_pd_smoothness_cost(const oracle_pd_params* p, const oracle_graph* g) {
  float cost = 0.0f;
  for (int32_t e = 0; e < g->E; ++e) {
    int32_t ii = g->edges[2 * e], jj = g->edges[2 * e + 1];
    float dx = g->pos[2 * ii] - g->pos[2 * jj];
    float dy = g->pos[2 * ii + 1] - g->pos[2 * jj + 1];
    float a = g->x[ii] - g->x[jj] - g->w1[ii] * dx - g->w2[ii] * dy;
    a = (a >= 0) ? a : -a;
    cost += g->alpha[e] * a;
    float b1 = g->w1[ii] - g->w1[jj];
    b1 = (b1 >= 0) ? b1 : -b1;
    float b2 = g->w2[ii] - g->w2[jj];
    b2 = (b2 >= 0) ? b2 : -b2;
    cost += g->beta[e] * b1 + g->beta[e] * b2;
  }
  return p->data_factor * cost;
}

/* cc:73-85 */
float oracle_pd_data_cost(const oracle_pd_params* p, const oracle_graph* g) {
  (void)p;
  float cost = 0.0f;
  for (int32_t v = 0; v < g->V; ++v) {
    float diff = (g->x[v] - g->data_term[v]) * g->data_weight[v];
    diff = (diff > 0) ? diff : -diff;
    cost += diff;
  }
  return cost;
}

/* flame.cc:2084-2105 */
void oracle_edge_weights(int32_t E, const int32_t* edges, const float* pos, float* alpha,
                         float* beta) {
  for (int32_t e = 0; e < E; ++e) {
    int32_t ii = edges[2 * e], jj = edges[2 * e + 1];
    float dx = pos[2 * ii] - pos[2 * jj];
    float dy = pos[2 * ii + 1] - pos[2 * jj + 1];
    float edge_length = (float)sqrt((double)(dx * dx + dy * dy)); /* flame.cc:2092 */
    alpha[e] = 1.0f / edge_length;                                /* flame.cc:2102 */
    beta[e] = 1.0f;                                               /* flame.cc:2103 */
  }
}

/* ====================================================================================
 * Dense projection — utils/rasterization.{h,cc}, utils/image_utils.cc:373-396
 * ==================================================================================== */

/* rasterization.h:108-118 */
static inline int min3i(int x, int y, int z) { return x < y ? (x < z ? x : z) : (y < z ? y : z); }
static inline int max3i(int x, int y, int z) { return x > y ? (x > z ? x : z) : (y > z ? y : z); }

/* rasterization.h:120-154  Edge::init, one lane at a time (the reference computes 4 lanes of
 * A*x + B*y + C with SSE mul/add; lane k has x = origin.x + k). */
typedef struct {
  float A, B, C;
  float one_step_x; /* A * 4 */
  float one_step_y; /* B * 1 */
} edge_fn;

static void edge_init(edge_fn* e, int v0x, int v0y, int v1x, int v1y, int ox, int oy, float row[4]) {
  e->A = (float)(v1y - v0y);
  e->B = (float)(v0x - v1x);
  e->C = (float)(v1x * v0y - v0x * v1y);
  e->one_step_x = e->A * 4;
  e->one_step_y = e->B * 1;
  for (int k = 0; k < 4; ++k) {
    float x = (float)(ox + k), y = (float)oy;
    row[k] = (e->A * x + e->B * y) + e->C;
  }
}

/* rasterization.cc:164-245 */
void oracle_draw_shaded_triangle_barycentric(int p1x, int p1y, int p2x, int p2y, int p3x,
                                             int p3y, float v1, float v2, float v3,
                                             float* img, int rows, int cols) {
  int xmin = min3i(p1x, p2x, p3x), ymin = min3i(p1y, p2y, p3y);
  int xmax = max3i(p1x, p2x, p3x), ymax = max3i(p1y, p2y, p3y);
  edge_fn e12, e23, e31;
  float w1_row[4], w2_row[4], w3_row[4];
  edge_init(&e23, p2x, p2y, p3x, p3y, xmin, ymin, w1_row);
  edge_init(&e31, p3x, p3y, p1x, p1y, xmin, ymin, w2_row);
  edge_init(&e12, p1x, p1y, p2x, p2y, xmin, ymin, w3_row);
  const long n = (long)rows * cols;

  for (int py = ymin; py <= ymax; py += 1) {
    float w1[4], w2[4], w3[4];
    memcpy(w1, w1_row, sizeof(w1));
    memcpy(w2, w2_row, sizeof(w2));
    memcpy(w3, w3_row, sizeof(w3));
    for (int px = xmin; px <= xmax; px += 4) {
      for (int k = 0; k < 4; ++k) {
        int mask = (w1[k] >= 0.0f) && (w2[k] >= 0.0f) && (w3[k] >= 0.0f);
        float norm = w1[k] + (w2[k] + w3[k]);
        float val = (v1 * w1[k] + (v2 * w2[k] + v3 * w3[k])) / norm;
        long addr = (long)py * cols + px + k; /* linear, as the reference's pointer math */
        if (mask && addr >= 0 && addr < n) img[addr] = val;
      }
      for (int k = 0; k < 4; ++k) {
        w1[k] = w1[k] + e23.one_step_x;
        w2[k] = w2[k] + e31.one_step_x;
        w3[k] = w3[k] + e12.one_step_x;
      }
    }
    for (int k = 0; k < 4; ++k) {
      w1_row[k] = w1_row[k] + e23.one_step_y;
      w2_row[k] = w2_row[k] + e31.one_step_y;
      w3_row[k] = w3_row[k] + e12.one_step_y;
    }
  }
}

/* cv::Point2f -> cv::Point conversion used at image_utils.cc:385-391: OpenCV saturate_cast<int>(float)
 * == cvRound == round-half-to-even in the default rounding mode. */
static inline int cv_round(float v) { return (int)lrintf(v); }

/* image_utils.cc:373-396 */
void oracle_interpolate_mesh(int32_t T, const int32_t* tris, int32_t V, const float* pos,
                             const float* values, const uint8_t* vtx_valid,
                             const uint8_t* tri_valid, float* img, int rows, int cols) {
  (void)V;
  for (int32_t ii = 0; ii < T; ++ii) {
    int32_t a = tris[3 * ii], b = tris[3 * ii + 1], c = tris[3 * ii + 2];
    if ((tri_valid == NULL || tri_valid[ii]) &&
        (vtx_valid == NULL || (vtx_valid[a] && vtx_valid[b] && vtx_valid[c]))) {
      oracle_draw_shaded_triangle_barycentric(
          cv_round(pos[2 * c]), cv_round(pos[2 * c + 1]), cv_round(pos[2 * b]),
          cv_round(pos[2 * b + 1]), cv_round(pos[2 * a]), cv_round(pos[2 * a + 1]), values[c],
          values[b], values[a], img, rows, cols);
    }
  }
}

/* ====================================================================================
 * Display filters — flame.cc:389-407, 2207-2356 (pinned on the verbatim slices of oracle/ref_capi/ref_filters.cc)
 * ==================================================================================== */
void oracle_tri_filter_default_params(oracle_tri_filter_params* p) { /* params.h:70-85 */
  p->do_oblique_triangle_filter = 1;
  p->oblique_normal_thresh = 1.39626f;
  p->oblique_idepth_diff_factor = 0.35f;
  p->oblique_idepth_diff_abs = 0.1f;
  p->do_edge_length_filter = 1;
  p->edge_length_thresh = 0.333f;
  p->do_idepth_triangle_filter = 1;
  p->min_triangle_idepth = 0.01f;
}

/* p = Kinv * (x, y, 1) / idepth (flame.cc:2226-2239) */
static void backproject(const float Kinv[9], float x, float y, float id, float out[3]) {
  for (int i = 0; i < 3; ++i) {
    float r = Kinv[3 * i] * x;
    r += Kinv[3 * i + 1] * y;
    r += Kinv[3 * i + 2] * 1.0f;
    out[i] = r / id;
  }
}

static void normalize3(float v[3]) { /* Eigen 3.2: *this /= norm() */
  float n2 = v[0] * v[0];
  n2 += v[1] * v[1];
  n2 += v[2] * v[2];
  const float n = sqrtf(n2);
  v[0] /= n; v[1] /= n; v[2] /= n;
}

void oracle_triangle_filters(const oracle_tri_filter_params* p, const float Kinv[9], int width, int32_t T,
                             const int32_t* tris, const float* pos, const float* idepths,
                             uint8_t* validity, float* angle_out) {
  for (int32_t ii = 0; ii < T; ++ii) validity[ii] = 1; /* flame.cc:389 */
  float dist_thresh2 = p->edge_length_thresh * width;  /* flame.cc:2298-2299 */
  dist_thresh2 *= dist_thresh2;
  for (int32_t ii = 0; ii < T; ++ii) {
    const int32_t a = tris[3 * ii], b = tris[3 * ii + 1], c = tris[3 * ii + 2];
    const float id0 = idepths[a], id1 = idepths[b], id2 = idepths[c];
    if (angle_out) angle_out[ii] = 0.0f;
    if (p->do_oblique_triangle_filter) { /* flame.cc:2220-2268 */
      float p0[3], p1[3], p2[3];
      backproject(Kinv, pos[2 * a], pos[2 * a + 1], id0, p0);
      backproject(Kinv, pos[2 * b], pos[2 * b + 1], id1, p1);
      backproject(Kinv, pos[2 * c], pos[2 * c + 1], id2, p2);
      const float d1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
      const float d2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
      float n[3] = {d1[1] * d2[2] - d1[2] * d2[1], d1[2] * d2[0] - d1[0] * d2[2],
                    d1[0] * d2[1] - d1[1] * d2[0]};
      normalize3(n);
      float ray[3] = {((p0[0] + p1[0]) + p2[0]) / 3, ((p0[1] + p1[1]) + p2[1]) / 3,
                      ((p0[2] + p1[2]) + p2[2]) / 3};
      normalize3(ray);
      float dot = ray[0] * n[0];
      dot += ray[1] * n[1];
      dot += ray[2] * n[2];
      const float angle = fabsf(acosf(dot));
      if (angle_out) angle_out[ii] = angle;
      if (angle > p->oblique_normal_thresh) validity[ii] = 0;
      float min_id = (id0 < id1) ? id0 : id1;
      min_id = (min_id < id2) ? min_id : id2;
      float max_id = (id0 > id1) ? id0 : id1;
      max_id = (max_id > id2) ? max_id : id2;
      if ((max_id - min_id) / max_id > p->oblique_idepth_diff_factor) validity[ii] = 0;
      if (max_id - min_id > p->oblique_idepth_diff_abs) validity[ii] = 0;
    }
    if (p->do_edge_length_filter) { /* flame.cc:2301-2317 */
      const float d01x = pos[2 * a] - pos[2 * b], d01y = pos[2 * a + 1] - pos[2 * b + 1];
      const float d02x = pos[2 * a] - pos[2 * c], d02y = pos[2 * a + 1] - pos[2 * c + 1];
      const float d12x = pos[2 * b] - pos[2 * c], d12y = pos[2 * b + 1] - pos[2 * c + 1];
      const float dist01 = d01x * d01x + d01y * d01y;
      const float dist02 = d02x * d02x + d02y * d02y;
      const float dist12 = d12x * d12x + d12y * d12y;
      if ((dist01 > dist_thresh2) || (dist02 > dist_thresh2) || (dist12 > dist_thresh2)) validity[ii] = 0;
    }
    if (p->do_idepth_triangle_filter) { /* flame.cc:2341-2349 */
      const float mean_idepth = (id0 + id1 + id2) / 3;
      if (mean_idepth < p->min_triangle_idepth) validity[ii] = 0;
    }
  }
}

/* flame.cc:2554-2640: running average of the (outward) normals of the triangles around each vertex, in
 * triangle order, re-normalised after every update */
void oracle_vertex_normals(const float Kinv[9], int32_t V, const float* pos, const float* idepths, int32_t T,
                           const int32_t* tris, float* normals) {
  int* counts = (int*)calloc((size_t)(V > 0 ? V : 1), sizeof(int));
  for (int32_t ii = 0; ii < 3 * V; ++ii) normals[ii] = 0.0f;
  for (int32_t ii = 0; ii < T; ++ii) {
    const int32_t v[3] = {tris[3 * ii], tris[3 * ii + 1], tris[3 * ii + 2]};
    const float id0 = idepths[v[0]], id1 = idepths[v[1]], id2 = idepths[v[2]];
    if ((id0 <= 0) || (id1 <= 0) || (id2 <= 0)) continue;
    float p0[3], p1[3], p2[3];
    backproject(Kinv, pos[2 * v[0]], pos[2 * v[0] + 1], id0, p0);
    backproject(Kinv, pos[2 * v[1]], pos[2 * v[1] + 1], id1, p1);
    backproject(Kinv, pos[2 * v[2]], pos[2 * v[2] + 1], id2, p2);
    const float d1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
    const float d2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
    float n[3] = {d2[1] * d1[2] - d2[2] * d1[1], d2[2] * d1[0] - d2[0] * d1[2],
                  d2[0] * d1[1] - d2[1] * d1[0]}; /* delta2.cross(delta1): outward */
    normalize3(n);
    for (int k = 0; k < 3; ++k) {
      float* nv = normals + 3 * (size_t)v[k];
      const float c = (float)counts[v[k]], c1 = (float)(counts[v[k]] + 1);
      nv[0] = (c * nv[0] + n[0]) / c1;
      nv[1] = (c * nv[1] + n[1]) / c1;
      nv[2] = (c * nv[2] + n[2]) / c1;
      normalize3(nv);
      counts[v[k]]++;
    }
  }
  free(counts);
}

/* flame.cc:434-443 */
int64_t oracle_coverage(int rows, int cols, const float* img) {
  int64_t coverage = 0;
  for (int ii = 0; ii < rows; ++ii)
    for (int jj = 0; jj < cols; ++jj)
      if (!isnan(img[(size_t)ii * cols + jj])) coverage++;
  return coverage;
}

/* ====================================================================================
 * Image helpers — utils/image_utils.h:199-272, utils/image_utils.cc:269-371
 * ==================================================================================== */

/* image_utils.h:199-214 */
static inline void bilinear_weights(float x, float y, float* w00, float* w01, float* w10,
                                    float* w11) {
  int x_floor = (int)x;
  int y_floor = (int)y;
  float dx = x - x_floor;
  float dy = y - y_floor;
  *w11 = dx * dy;
  *w01 = dx - *w11;
  *w10 = dy - *w11;
  *w00 = 1.0f - dx - dy + *w11;
}

/* image_utils.h:230-255, ChannelType = uint8_t */
float oracle_bilinear_u8(int rows, int cols, int step, const uint8_t* data, float x, float y) {
  (void)rows; (void)cols;
  int x_floor = (int)x, y_floor = (int)y;
  float w00, w01, w10, w11;
  bilinear_weights(x, y, &w00, &w01, &w10, &w11);
  const uint8_t* d = data + (size_t)y_floor * step + x_floor;
  return w00 * d[0] + w01 * d[1] + w10 * d[step] + w11 * d[1 + step];
}

/* image_utils.h:230-255, ChannelType = float */
float oracle_bilinear_f32(int rows, int cols, int step, const float* data, float x, float y) {
  (void)rows; (void)cols;
  int x_floor = (int)x, y_floor = (int)y;
  float w00, w01, w10, w11;
  bilinear_weights(x, y, &w00, &w01, &w10, &w11);
  const uint8_t* d = (const uint8_t*)data + (size_t)y_floor * step + (size_t)x_floor * 4;
  return w00 * (*(const float*)d) + w01 * (*(const float*)(d + 4)) +
         w10 * (*(const float*)(d + step)) + w11 * (*(const float*)(d + 4 + step));
}

/* image_utils.cc:269-371 */
int oracle_clip_line_liang_barsky(float xmin, float xmax, float ymin, float ymax, float x0,
                                  float y0, float x1, float y1, float* x0c, float* y0c,
                                  float* x1c, float* y1c) {
  float t0 = 0.0f, t1 = 1.0f;
  float xdelta = x1 - x0, ydelta = y1 - y0;
  for (int edge = 0; edge < 4; edge++) {
    float p = 1.0f, q = 0.0f, r = 0.0f;
    if (edge == 0) {
      p = -xdelta; q = -(xmin - x0);
    } else if (edge == 1) {
      p = xdelta; q = (xmax - x0);
    } else if (edge == 2) {
      p = -ydelta; q = -(ymin - y0);
    } else {
      p = ydelta; q = (ymax - y0);
    }
    r = q / p;                                  /* cc:309 (before the p == 0 test) */
    if (p == 0 && q < 0) return 0;              /* cc:310-311 */
    if (p < 0) {
      if (r > t1) return 0;
      else if (r > t0) t0 = r;
    } else if (p > 0) {
      if (r < t0) return 0;
      else if (r < t1) t1 = r;
    }
  }
  float a0 = x0 + t0 * xdelta, b0 = y0 + t0 * ydelta;
  float a1 = x0 + t1 * xdelta, b1 = y0 + t1 * ydelta;
  if (a0 < xmin) a0 = xmin;
  if (a0 > xmax) a0 = xmax;
  if (b0 < ymin) b0 = ymin;
  if (b0 > ymax) b0 = ymax;
  if (a1 < xmin) a1 = xmin;
  if (a1 > xmax) a1 = xmax;
  if (b1 < ymin) b1 = ymin;
  if (b1 > ymax) b1 = ymax;
  *x0c = a0; *y0c = b0; *x1c = a1; *y1c = b1;
  return 1;
}

/* ====================================================================================
 * Frame preparation — utils/image_utils.cc:36-108, utils/frame.cc:54-63
 * ==================================================================================== */

/* image_utils.cc:36-108 / image_utils.h:425-477: 0.5*(right-left) inside, forward / backward
 * differences on the first / last column and row. */
void oracle_central_gradient_u8(int width, int height, const uint8_t* img, float* gradx, float* grady) {
  for (int ii = 0; ii < height; ++ii) {
    for (int jj = 1; jj < width - 1; ++jj) {
      int idx = ii * width + jj;
      float left = img[idx - 1], right = img[idx + 1];
      gradx[idx] = (float)(0.5 * (right - left));
    }
    gradx[ii * width] = (float)img[ii * width + 1] - (float)img[ii * width];
    gradx[ii * width + width - 1] = (float)img[ii * width + width - 1] - (float)img[ii * width + width - 2];
  }
  for (int ii = 1; ii < height - 1; ++ii)
    for (int jj = 0; jj < width; ++jj) {
      int idx = ii * width + jj;
      float up = img[idx - width], down = img[idx + width];
      grady[idx] = (float)(0.5 * (down - up));
    }
  for (int jj = 0; jj < width; ++jj) {
    grady[jj] = (float)img[jj + width] - (float)img[jj];
    grady[(height - 1) * width + jj] = (float)img[(height - 1) * width + jj] - (float)img[(height - 2) * width + jj];
  }
}

/* OpenCV borderInterpolate(p, len, BORDER_REFLECT_101): gfedcb|abcdefgh|gfedcba */
static int reflect101(int p, int len) {
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}

/* utils/frame.cc:54-63 (level 0; Flame passes num_levels = 1, border = fparams.win_size, flame.cc:149-150) */
void oracle_frame_prep(int width, int height, int border, const uint8_t* img, uint8_t* img_pad,
                       float* gradx_pad, float* grady_pad) {
  const int pw = width + 2 * border, ph = height + 2 * border;
  float* gx = (float*)malloc(sizeof(float) * (size_t)width * height);
  float* gy = (float*)malloc(sizeof(float) * (size_t)width * height);
  oracle_central_gradient_u8(width, height, img, gx, gy);
  for (int y = 0; y < ph; ++y)
    for (int x = 0; x < pw; ++x) {
      const int sx = x - border, sy = y - border;
      img_pad[(size_t)y * pw + x] = img[(size_t)reflect101(sy, height) * width + reflect101(sx, width)];
      const int inside = (sx >= 0 && sx < width && sy >= 0 && sy < height);
      gradx_pad[(size_t)y * pw + x] = inside ? gx[(size_t)sy * width + sx] : 0.0f;
      grady_pad[(size_t)y * pw + x] = inside ? gy[(size_t)sy * width + sx] : 0.0f;
    }
  free(gx);
  free(gy);
}

/* ====================================================================================
 * Eigen / Sophus arithmetic the reference relies on (third-party, not vendored:
 * Eigen 3.2.10, Sophus b474f05 — README.md:30-35, scripts/eigen.sh, scripts/sophus.sh).
 * Restated from their published algorithms; float summation order inside these helpers is
 * NOT verifiable offline, so geometry parity is tolerance-based.
 * ==================================================================================== */

/* Eigen Quaternion::toRotationMatrix (Geometry/Quaternion.h). q = (w,x,y,z). Row-major out. */
static void quat_to_rot(const float q[4], float R[9]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  const float tx = 2.0f * x, ty = 2.0f * y, tz = 2.0f * z;
  const float twx = tx * w, twy = ty * w, twz = tz * w;
  const float txx = tx * x, txy = ty * x, txz = tz * x;
  const float tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1.0f - (tyy + tzz); R[1] = txy - twz;          R[2] = txz + twy;
  R[3] = txy + twz;          R[4] = 1.0f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;          R[7] = tyz + twx;          R[8] = 1.0f - (txx + tyy);
}

/* Eigen QuaternionBase::_transformVector: uv = q.vec x v; uv += uv; v + w*uv + q.vec x uv */
static void quat_rotate(const float q[4], const float v[3], float out[3]) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  float uvx = y * v[2] - z * v[1];
  float uvy = z * v[0] - x * v[2];
  float uvz = x * v[1] - y * v[0];
  uvx += uvx; uvy += uvy; uvz += uvz;
  out[0] = (v[0] + w * uvx) + (y * uvz - z * uvy);
  out[1] = (v[1] + w * uvy) + (z * uvx - x * uvz);
  out[2] = (v[2] + w * uvz) + (x * uvy - y * uvx);
}

/* Eigen QuaternionBase::inverse: conjugate / squaredNorm */
static void quat_inverse(const float q[4], float out[4]) {
  float n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0];
  if (n2 > 0) {
    out[0] = q[0] / n2; out[1] = -q[1] / n2; out[2] = -q[2] / n2; out[3] = -q[3] / n2;
  } else {
    out[0] = out[1] = out[2] = out[3] = 0.0f;
  }
}

/* Eigen quaternion product a*b (scalar path) */
static void quat_mul(const float a[4], const float b[4], float out[4]) {
  out[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  out[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  out[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  out[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

/* 3x3 row-major product, coefficient = l0*r0 + l1*r1 + l2*r2 accumulated left to right
 * (Eigen 3.2 product_coeff_impl<DefaultTraversal>). */
static void mat3_mul(const float A[9], const float B[9], float C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      float r = A[3 * i] * B[j];
      r += A[3 * i + 1] * B[3 + j];
      r += A[3 * i + 2] * B[6 + j];
      C[3 * i + j] = r;
    }
}

static void mat3_vec(const float A[9], const float v[3], float out[3]) {
  for (int i = 0; i < 3; ++i) {
    float r = A[3 * i] * v[0];
    r += A[3 * i + 1] * v[1];
    r += A[3 * i + 2] * v[2];
    out[i] = r;
  }
}

/* Sophus SE3Group: inverse() = (R^-1, R^-1 * (-t)); a*b = (Ra*Rb [renormalised], ta + Ra*tb).
 * out = a^-1 * b.  (flame.cc:1314-1317, inverse_depth_meas_model.h:69) */
void oracle_se3_inv_mul(const float a[7], const float b[7], float out[7]) {
  float qi[4] = {a[0], -a[1], -a[2], -a[3]}; /* SO3 inverse = conjugate */
  float nt[3] = {a[4] * -1.0f, a[5] * -1.0f, a[6] * -1.0f};
  float ti[3];
  quat_rotate(qi, nt, ti);
  float rb[3];
  quat_rotate(qi, b + 4, rb);
  float q[4];
  quat_mul(qi, b, q);
  float n2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3] + q[0] * q[0];
  if (n2 != 1.0f) { /* Sophus SO3 operator*=: first-order renormalisation */
    float s = 2.0f / (1.0f + n2);
    q[0] *= s; q[1] *= s; q[2] *= s; q[3] *= s;
  }
  out[0] = q[0]; out[1] = q[1]; out[2] = q[2]; out[3] = q[3];
  out[4] = ti[0] + rb[0]; out[5] = ti[1] + rb[1]; out[6] = ti[2] + rb[2];
}

/* ====================================================================================
 * stereo/epipolar_geometry.h
 * ==================================================================================== */

void oracle_epigeo_init(oracle_epigeo* g, const float K[9], const float Kinv[9]) {
  memset(g, 0, sizeof(*g));
  memcpy(g->K, K, sizeof(g->K));
  memcpy(g->Kinv, Kinv, sizeof(g->Kinv));
  g->q[0] = 1.0f;
}

/* :86-103 */
void oracle_epigeo_load(oracle_epigeo* g, const float q[4], const float t[3]) {
  memcpy(g->q, q, sizeof(g->q));
  memcpy(g->t, t, sizeof(g->t));
  float qi[4], r[3];
  quat_inverse(q, qi);
  quat_rotate(qi, t, r);
  g->t_cmp_to_ref[0] = -r[0]; g->t_cmp_to_ref[1] = -r[1]; g->t_cmp_to_ref[2] = -r[2];
  float R[9], KR[9];
  quat_to_rot(q, R);
  mat3_mul(g->K, R, KR);
  mat3_mul(KR, g->Kinv, g->KRKinv);
  g->KRKinv3[0] = g->KRKinv[6]; g->KRKinv3[1] = g->KRKinv[7]; g->KRKinv3[2] = g->KRKinv[8];
  mat3_vec(g->K, t, g->Kt);
  if (t[2] > 0) {
    g->epipole[0] = (g->K[0] * t[0] + g->K[2] * t[2]) / t[2];
    g->epipole[1] = (g->K[4] * t[1] + g->K[5] * t[2]) / t[2];
  }
}

/* :115-120 */
void oracle_epigeo_project_world(const float K[9], const float q[4], const float t[3],
                                 const float p[3], float* ux, float* uy) {
  float qi[4], d[3] = {p[0] - t[0], p[1] - t[1], p[2] - t[2]}, pc[3];
  quat_inverse(q, qi);
  quat_rotate(qi, d, pc);
  *ux = (K[0] * pc[0] + K[2] * pc[2]) / pc[2];
  *uy = (K[4] * pc[1] + K[5] * pc[2]) / pc[2];
}

/* :191-201 */
void oracle_epigeo_max_depth_projection(const oracle_epigeo* g, float ux, float uy, float* ix,
                                        float* iy) {
  float h[3] = {ux, uy, 1.0f}, c[3];
  mat3_vec(g->KRKinv, h, c);
  float inv = 1.0f / c[2];
  *ix = c[0] * inv;
  *iy = c[1] * inv;
}

/* :127-143 */
void oracle_epigeo_project(const oracle_epigeo* g, float ux, float uy, float idepth, float* cx,
                           float* cy) {
  if (idepth == 0.0f) {
    oracle_epigeo_max_depth_projection(g, ux, uy, cx, cy);
    return;
  }
  float depth = 1.0f / idepth;
  float h[3] = {ux * depth, uy * depth, depth}, c[3];
  mat3_vec(g->KRKinv, h, c);
  c[0] = c[0] + g->Kt[0]; c[1] = c[1] + g->Kt[1]; c[2] = c[2] + g->Kt[2];
  float inv = 1.0f / c[2];
  *cx = c[0] * inv;
  *cy = c[1] * inv;
}

/* :153-180 */
void oracle_epigeo_project_idepth(const oracle_epigeo* g, float ux, float uy, float idepth,
                                  float* cx, float* cy, float* new_idepth) {
  if (idepth == 0.0f) {
    oracle_epigeo_max_depth_projection(g, ux, uy, cx, cy);
    *new_idepth = 0.0f;
    return;
  }
  float depth = 1.0f / idepth;
  float p_ref[3] = {g->Kinv[0] * ux + g->Kinv[2], g->Kinv[4] * uy + g->Kinv[5], 1.0f};
  p_ref[0] *= depth; p_ref[1] *= depth; p_ref[2] *= depth;
  float r[3];
  quat_rotate(g->q, p_ref, r);
  float pc[3] = {r[0] + g->t[0], r[1] + g->t[1], r[2] + g->t[2]};
  float u0 = g->K[0] * pc[0] + g->K[2] * pc[2];
  float u1 = g->K[4] * pc[1] + g->K[5] * pc[2];
  *new_idepth = 1.0f / pc[2];
  *cx = u0 * (*new_idepth);
  *cy = u1 * (*new_idepth);
}

/* :237-263 */
void oracle_epigeo_min_depth_projection(const oracle_epigeo* g, float ux, float uy, float* mx,
                                        float* my) {
  if (g->t[2] > 0) {
    *mx = g->epipole[0];
    *my = g->epipole[1];
  } else if (g->t[2] == 0) {
    float ex = g->K[0] * g->t[0], ey = g->K[4] * g->t[1];
    float ix, iy;
    oracle_epigeo_max_depth_projection(g, ux, uy, &ix, &iy);
    /* u_inf + 1e6 * epi: cv::operator*(double, Point2f) rounds the double product to float */
    *mx = ix + (float)(1e6 * (double)ex);
    *my = iy + (float)(1e6 * (double)ey);
  } else {
    float qp[3] = {g->Kinv[0] * ux + g->Kinv[2], g->Kinv[4] * uy + g->Kinv[5], 1.0f}, r[3];
    quat_rotate(g->q, qp, r);
    float min_depth = (1.0f - g->t[2]) / r[2];
    float pc[3] = {min_depth * r[0] + g->t[0], min_depth * r[1] + g->t[1],
                   min_depth * r[2] + g->t[2]};
    *mx = (g->K[0] * pc[0] + g->K[2] * pc[2]) / pc[2];
    *my = (g->K[4] * pc[1] + g->K[5] * pc[2]) / pc[2];
  }
}

/* :282-300 */
void oracle_epigeo_epiline(const oracle_epigeo* g, float ux, float uy, float* ix, float* iy,
                           float* ex, float* ey) {
  float zx, zy;
  oracle_epigeo_min_depth_projection(g, ux, uy, &zx, &zy);
  oracle_epigeo_max_depth_projection(g, ux, uy, ix, iy);
  float epx = zx - *ix, epy = zy - *iy;
  float norm2 = epx * epx + epy * epy;
  if ((double)norm2 > 1e-10) {
    /* `1.0f / sqrt(norm2)`: unqualified sqrt on a float resolves to ::sqrt(double), so the quotient is
     * formed in double and rounded once on assignment (verified against oracle/_ref) */
    float inv_norm = (float)(1.0 / sqrt((double)norm2));
    epx *= inv_norm;
    epy *= inv_norm;
  } else {
    epx = 0.0f;
    epy = 0.0f;
  }
  *ex = epx;
  *ey = epy;
}

/* :311-331 */
void oracle_epigeo_reference_epiline(const oracle_epigeo* g, float ux, float uy, float* ex,
                                     float* ey) {
  float rx = -g->K[0] * g->t_cmp_to_ref[0] + g->t_cmp_to_ref[2] * (ux - g->K[2]);
  float ry = -g->K[4] * g->t_cmp_to_ref[1] + g->t_cmp_to_ref[2] * (uy - g->K[5]);
  float n2 = rx * rx + ry * ry;
  float inv = (float)(1.0 / sqrt((double)n2)); /* :325, double quotient as above */
  *ex = rx * inv;
  *ey = ry * inv;
}

/* :342-347 */
float oracle_epigeo_disparity(const oracle_epigeo* g, float ux, float uy, float cx, float cy,
                              float* ix, float* iy, float* ex, float* ey) {
  oracle_epigeo_epiline(g, ux, uy, ix, iy, ex, ey);
  return *ex * (cx - *ix) + *ey * (cy - *iy);
}

/* :365-380 */
float oracle_epigeo_disparity_to_depth(const oracle_epigeo* g, float ux, float uy, float ix,
                                       float iy, float ex, float ey, float disp) {
  float w = g->KRKinv3[0] * ux + g->KRKinv3[1] * uy + g->KRKinv3[2];
  float wd = w * disp;
  float Ax = ex * wd, Ay = ey * wd; /* cv: Point_(b.x*a, b.y*a) */
  float bx = g->Kt[0] - g->Kt[2] * (ix + disp * ex);
  float by = g->Kt[1] - g->Kt[2] * (iy + disp * ey);
  float ATA = Ax * Ax + Ay * Ay;
  float ATb = Ax * bx + Ay * by;
  return ATb / ATA;
}

/* :392-407 */
float oracle_epigeo_disparity_to_idepth(const oracle_epigeo* g, float ux, float uy, float ix,
                                        float iy, float ex, float ey, float disp) {
  float w = g->KRKinv3[0] * ux + g->KRKinv3[1] * uy + g->KRKinv3[2];
  float Ax = g->Kt[0] - g->Kt[2] * (ix + disp * ex);
  float Ay = g->Kt[1] - g->Kt[2] * (iy + disp * ey);
  float wd = w * disp;
  float bx = ex * wd, by = ey * wd;
  float ATA = Ax * Ax + Ay * Ay;
  float ATb = Ax * bx + Ay * by;
  return ATb / ATA;
}

/* ====================================================================================
 * stereo/line_stereo.h:75-385
 * ==================================================================================== */

void oracle_filter_default_params(oracle_filter_params* p) {
  /* inverse_depth_filter.h:48-68 */
  p->win_size = 5;
  p->search_sigma = 2.0f;
  p->min_grad_mag = 5.0f;
  p->idepth_min = 1e-3f;
  p->idepth_max = 2.0f;
  p->epilength_min = 3.0f;
  p->epilength_max = 32.0f;
  p->process_var_factor = 1.01f;
  p->process_fail_var_factor = 1.1f;
  /* line_stereo.h:47-59 */
  p->sparams.max_cost = 1300.0f;
  p->sparams.do_subpixel = 1;
  p->sparams.sample_dist = 1.0f;
  p->sparams.second_best_factor = 1.5f;
}

void oracle_meas_default_params(oracle_meas_params* p) {
  /* inverse_depth_meas_model.h:44-52 */
  p->win_size = 5;
  p->pixel_var = 16.0f;
  p->epipolar_line_var = 1.0f;
}

int oracle_line_stereo_match(float rescale_factor, const float ref_patch[5], const uint8_t* img,
                             int rows, int cols, int step, float sx, float sy, float ex,
                             float ey, float* ucx, float* ucy, float* residual,
                             const oracle_ls_params* params) {
  float realVal_m2 = ref_patch[0], realVal_m1 = ref_patch[1], realVal = ref_patch[2];
  float realVal_p1 = ref_patch[3], realVal_p2 = ref_patch[4];

  float incx = ex - sx, incy = ey - sy;
  float eplLength = (float)sqrt((double)(incx * incx + incy * incy));
  incx *= params->sample_dist / eplLength;
  incy *= params->sample_dist / eplLength;

  float cpx = sx, cpy = sy;
  float val_cp_m2 = oracle_bilinear_u8(rows, cols, step, img, cpx - 2.0f * incx, cpy - 2.0f * incy);
  float val_cp_m1 = oracle_bilinear_u8(rows, cols, step, img, cpx - incx, cpy - incy);
  float val_cp = oracle_bilinear_u8(rows, cols, step, img, cpx, cpy);
  float val_cp_p1 = oracle_bilinear_u8(rows, cols, step, img, cpx + incx, cpy + incy);
  float val_cp_p2;

  int loopCounter = 0;
  float best_match_x = -1, best_match_y = -1;
  float best_match_err = FLT_MAX;
  float second_best_match_err = FLT_MAX;
  float best_match_errPre = NAN, best_match_errPost = NAN;
  float best_match_DiffErrPre = NAN, best_match_DiffErrPost = NAN;
  int bestWasLastLoop = 0;
  float eeLast = -1;
  float e1A = NAN, e1B = NAN, e2A = NAN, e2B = NAN, e3A = NAN, e3B = NAN, e4A = NAN, e4B = NAN,
        e5A = NAN, e5B = NAN;
  int loopCBest = -1, loopCSecond = -1;

  while ((((incx < 0) == (cpx > ex)) && ((incy < 0) == (cpy > ey))) || loopCounter == 0) {
    val_cp_p2 = oracle_bilinear_u8(rows, cols, step, img, cpx + 2 * incx, cpy + 2 * incy);
    float ee = 0.0f;
    if (loopCounter % 2 == 0) {
      e1A = val_cp_p2 - realVal_p2; ee += e1A * e1A;
      e2A = val_cp_p1 - realVal_p1; ee += e2A * e2A;
      e3A = val_cp - realVal;       ee += e3A * e3A;
      e4A = val_cp_m1 - realVal_m1; ee += e4A * e4A;
      e5A = val_cp_m2 - realVal_m2; ee += e5A * e5A;
    } else {
      e1B = val_cp_p2 - realVal_p2; ee += e1B * e1B;
      e2B = val_cp_p1 - realVal_p1; ee += e2B * e2B;
      e3B = val_cp - realVal;       ee += e3B * e3B;
      e4B = val_cp_m1 - realVal_m1; ee += e4B * e4B;
      e5B = val_cp_m2 - realVal_m2; ee += e5B * e5B;
    }

    if (ee < best_match_err) {
      second_best_match_err = best_match_err;
      loopCSecond = loopCBest;
      best_match_err = ee;
      loopCBest = loopCounter;
      best_match_errPre = eeLast;
      best_match_DiffErrPre = e1A * e1B + e2A * e2B + e3A * e3B + e4A * e4B + e5A * e5B;
      best_match_errPost = -1;
      best_match_DiffErrPost = -1;
      best_match_x = cpx;
      best_match_y = cpy;
      bestWasLastLoop = 1;
    } else {
      if (bestWasLastLoop) {
        best_match_errPost = ee;
        best_match_DiffErrPost = e1A * e1B + e2A * e2B + e3A * e3B + e4A * e4B + e5A * e5B;
        bestWasLastLoop = 0;
      }
      if (ee < second_best_match_err) {
        second_best_match_err = ee;
        loopCSecond = loopCounter;
      }
    }

    eeLast = ee;
    val_cp_m2 = val_cp_m1;
    val_cp_m1 = val_cp;
    val_cp = val_cp_p1;
    val_cp_p1 = val_cp_p2;
    cpx += incx;
    cpy += incy;
    loopCounter++;
  }

  *residual = best_match_err;
  if (best_match_err > 4.0f * params->max_cost) return ORACLE_LS_FAIL_MAX_COST; /* :261 */

  /* :275-276: fast_abs(int) > 1.0f */
  int dl = loopCBest - loopCSecond;
  dl = (dl > 0) ? dl : -dl;
  if (((float)dl > 1.0f) && (params->second_best_factor * best_match_err > second_best_match_err))
    return ORACLE_LS_FAIL_AMBIGUOUS_MATCH;

  if (params->do_subpixel) { /* :288-350 */
    float gradPre_pre = -(best_match_errPre - best_match_DiffErrPre);
    float gradPre_this = +(best_match_err - best_match_DiffErrPre);
    float gradPost_this = -(best_match_err - best_match_DiffErrPost);
    float gradPost_post = +(best_match_errPost - best_match_DiffErrPost);
    int interpPost = 0, interpPre = 0;
    if (best_match_errPre < 0 || best_match_errPost < 0) {
    } else if ((gradPost_this < 0) ^ (gradPre_this < 0)) {
    } else if ((gradPre_pre < 0) ^ (gradPre_this < 0)) {
      if ((gradPost_post < 0) ^ (gradPost_this < 0)) {
      } else {
        interpPre = 1;
      }
    } else if ((gradPost_post < 0) ^ (gradPost_this < 0)) {
      interpPost = 1;
    }
    if (interpPre) {
      float d = gradPre_this / (gradPre_this - gradPre_pre);
      best_match_x -= d * incx;
      best_match_y -= d * incy;
      best_match_err = best_match_err - 2 * d * gradPre_this - (gradPre_pre - gradPre_this) * d * d;
    } else if (interpPost) {
      float d = gradPost_this / (gradPost_this - gradPost_post);
      best_match_x += d * incx;
      best_match_y += d * incy;
      best_match_err =
          best_match_err + 2 * d * gradPost_this + (gradPost_post - gradPost_this) * d * d;
    }
  }
  *residual = best_match_err;

  float sampleDist = params->sample_dist * rescale_factor; /* :355 */
  float gradAlongLine = 0;
  float tmp = realVal_p2 - realVal_p1;
  gradAlongLine += tmp * tmp;
  tmp = realVal_p1 - realVal;
  gradAlongLine += tmp * tmp;
  tmp = realVal - realVal_m1;
  gradAlongLine += tmp * tmp;
  tmp = realVal_m1 - realVal_m2;
  gradAlongLine += tmp * tmp;
  gradAlongLine /= sampleDist * sampleDist;

  if (best_match_err > params->max_cost + sqrtf(gradAlongLine) * 20) /* :370 */
    return ORACLE_LS_FAIL_MAX_COST;

  *ucx = best_match_x;
  *ucy = best_match_y;
  return ORACLE_LS_SUCCESS;
}

/* ====================================================================================
 * stereo/inverse_depth_filter.cc
 * ==================================================================================== */

/* :35-63 */
int oracle_filter_predict(const oracle_epigeo* g, float process_var_factor, float ux, float uy,
                          float mu, float var, float* cx, float* cy, float* mu_pred,
                          float* var_pred) {
  oracle_epigeo_project_idepth(g, ux, uy, mu, cx, cy, mu_pred);
  if (*mu_pred < 0.0f) {
    *mu_pred = 0.0f;
    *var_pred = (float)1e10;
    return 0;
  }
  float var_factor4 = *mu_pred / mu;
  var_factor4 *= var_factor4;
  var_factor4 *= var_factor4;
  if ((double)mu < 1e-6) var_factor4 = 1;
  *var_pred = process_var_factor * var_factor4 * var;
  return 1;
}

/* :65-182 */
int oracle_filter_get_search_region(const oracle_filter_params* p, const oracle_epigeo* g,
                                    int width, int height, float ux, float uy, float mu_prior,
                                    float var_prior, float* sx, float* sy, float* ex, float* ey,
                                    float* epx, float* epy) {
  float id_min = p->idepth_min, id_max = p->idepth_max;
  if (!isnan(mu_prior) && !isnan(var_prior)) {
    float sigma = (float)sqrt((double)var_prior);
    id_min = mu_prior - p->search_sigma * sigma;
    id_max = mu_prior + p->search_sigma * sigma;
  }
  id_min = (id_min < p->idepth_min) ? p->idepth_min : id_min;
  id_max = (id_max > p->idepth_max) ? p->idepth_max : id_max;
  if (id_max < id_min) return 0;

  oracle_epigeo_project(g, ux, uy, id_min, sx, sy);
  oracle_epigeo_project(g, ux, uy, id_max, ex, ey);

  float dx = *ex - *sx, dy = *ey - *sy;
  float epilength = (float)sqrt((double)(dx * dx + dy * dy));
  if (epilength <= 0) return 0;
  *epx = dx / epilength;
  *epy = dy / epilength;

  /* cv::Rect valid_region(1, 1, width-2, height-2): tl = (1,1), br = (width-1, height-1) */
  float rx0 = 1.0f, ry0 = 1.0f, rx1 = (float)(width - 1), ry1 = (float)(height - 1);
  float a, b, c, d;
  if (!oracle_clip_line_liang_barsky(rx0, rx1, ry0, ry1, *sx, *sy, *ex, *ey, &a, &b, &c, &d))
    return 0;
  *sx = a; *sy = b; *ex = c; *ey = d;

  dx = *ex - *sx; dy = *ey - *sy;
  epilength = (float)sqrt((double)(dx * dx + dy * dy));
  if (epilength <= 0) return 0;

  if (epilength < p->epilength_min) {
    float pad = (p->epilength_min - epilength) / 2.0f;
    *sx -= pad * (*epx); *sy -= pad * (*epy); /* cv: Point_(b.x*a, b.y*a) then -= */
    *ex += pad * (*epx); *ey += pad * (*epy);
  }
  if (epilength > p->epilength_max) {
    epilength = p->epilength_max;
    *ex = *sx + epilength * (*epx);
    *ey = *sy + epilength * (*epy);
  }

  if (!oracle_clip_line_liang_barsky(rx0, rx1, ry0, ry1, *sx, *sy, *ex, *ey, &a, &b, &c, &d))
    return 0;
  *sx = a; *sy = b; *ex = c; *ey = d;
  return 1;
}

/* :184-266 */
int oracle_filter_search(const oracle_filter_params* p, const oracle_epigeo* g,
                         float rescale_factor, const uint8_t* img_ref, const uint8_t* img_cmp,
                         int rows, int cols, int step, float urx, float ury, float sx, float sy,
                         float ex, float ey, float* ucx, float* ucy) {
  float erx, ery;
  oracle_epigeo_reference_epiline(g, urx, ury, &erx, &ery);

  float ref_patch[5];
  ref_patch[0] = oracle_bilinear_u8(rows, cols, step, img_ref, urx - 2 * erx * rescale_factor,
                                    ury - 2 * ery * rescale_factor);
  ref_patch[1] = oracle_bilinear_u8(rows, cols, step, img_ref, urx - erx * rescale_factor,
                                    ury - ery * rescale_factor);
  ref_patch[2] = oracle_bilinear_u8(rows, cols, step, img_ref, urx, ury);
  ref_patch[3] = oracle_bilinear_u8(rows, cols, step, img_ref, urx + erx * rescale_factor,
                                    ury + ery * rescale_factor);
  ref_patch[4] = oracle_bilinear_u8(rows, cols, step, img_ref, urx + 2 * erx * rescale_factor,
                                    ury + 2 * ery * rescale_factor);

  float ref_patch_grad = 0.0f;
  for (int ii = 1; ii < 5; ++ii) {
    float abs_grad = fast_absf(ref_patch[ii] - ref_patch[ii - 1]);
    if (abs_grad > ref_patch_grad) ref_patch_grad = abs_grad;
  }
  if (ref_patch_grad < p->min_grad_mag) return ORACLE_FAIL_REF_PATCH_GRADIENT;

  float residual = FLT_MAX;
  int result = oracle_line_stereo_match(rescale_factor, ref_patch, img_cmp, rows, cols, step, sx,
                                        sy, ex, ey, ucx, ucy, &residual, &p->sparams);
  if (result == ORACLE_LS_FAIL_AMBIGUOUS_MATCH) return ORACLE_FAIL_AMBIGUOUS_MATCH;
  if (result == ORACLE_LS_FAIL_MAX_COST) return ORACLE_FAIL_MAX_COST;
  return ORACLE_SEARCH_SUCCESS;
}

/* :268-305 */
int oracle_filter_update(float mu_pred, float var_pred, float mu_meas, float var_meas,
                         float* mu_post, float* var_post, float outlier_sigma_thresh) {
  if (!isnan(mu_pred) && (mu_pred > 0.0f)) {
    float w = var_pred + var_meas;
    *mu_post = (var_meas * mu_pred + var_pred * mu_meas) / w;
    *var_post = (var_pred * var_meas) / w;
  } else {
    *mu_post = mu_meas;
    *var_post = var_meas;
  }
  float res = (mu_meas - mu_pred);
  float dist = res * res / var_pred;
  if (dist > outlier_sigma_thresh * outlier_sigma_thresh) return 0;
  *mu_post = (*mu_post <= 0) ? 0.0f : *mu_post;
  return 1;
}

/* ====================================================================================
 * stereo/inverse_depth_meas_model.cc:49-156
 * ==================================================================================== */
int oracle_meas_idepth(const oracle_meas_params* p, const oracle_epigeo* g, const float* gradx_cmp,
                       const float* grady_cmp, int rows, int cols, float urx, float ury, float ucx,
                       float ucy, float* mu, float* var) {
  float ix, iy, ex, ey;
  float disp = oracle_epigeo_disparity(g, urx, ury, ucx, ucy, &ix, &iy, &ex, &ey);
  if ((double)disp < 1e-3) {
    *mu = 0.0f; *var = (float)1e10;
    return 0;
  }
  *mu = oracle_epigeo_disparity_to_idepth(g, urx, ury, ix, iy, ex, ey, disp);
  if (*mu < 0.0f) {
    *mu = 0.0f; *var = (float)1e10;
    return 0;
  }
  float off = (float)(p->win_size / 2 + 1); /* cc:87 — NOT the padding */
  float gx = oracle_bilinear_f32(rows, cols, cols * 4, gradx_cmp, ucx + off, ucy + off);
  float gy = oracle_bilinear_f32(rows, cols, cols * 4, grady_cmp, ucx + off, ucy + off);
  float gnorm = (float)sqrt((double)(gx * gx + gy * gy));
  if ((double)gnorm < 1e-3) {
    *mu = 0.0f; *var = (float)1e10;
    return 0;
  }
  float ngx = gx / gnorm, ngy = gy / gnorm;
  float epi_dot_ngrad = ngx * ex + ngy * ey;
  float geo_var = p->epipolar_line_var / (epi_dot_ngrad * epi_dot_ngrad);
  if ((double)fast_absf(epi_dot_ngrad) < 1e-3) {
    *mu = 0.0f; *var = (float)1e10;
    return 0;
  }
  float epi_dot_grad = gx * ex + gy * ey;
  float photo_var = 2 * p->pixel_var / (epi_dot_grad * epi_dot_grad);
  float disp_min = disp - disp / 10;
  float disp_max = disp + disp / 10;
  float idepth_min = oracle_epigeo_disparity_to_idepth(g, urx, ury, ix, iy, ex, ey, disp_min);
  float idepth_max = oracle_epigeo_disparity_to_idepth(g, urx, ury, ix, iy, ex, ey, disp_max);
  float alpha = (idepth_max - idepth_min) / (disp_max - disp_min);
  *var = alpha * alpha * (geo_var + photo_var);
  return 1;
}

/* ====================================================================================
 * flame.cc:1280-1752 — updateFeatureIDepths + trackFeature (debug drawing removed)
 * ==================================================================================== */

void oracle_track_default_params(oracle_track_params* p) {
  oracle_filter_default_params(&p->fparams);
  oracle_meas_default_params(&p->zparams);
  /* params.h:58-67,90 */
  p->min_baseline = 0.01f;
  p->rescale_factor_min = 0.7f;
  p->rescale_factor_max = 1.4f;
  p->idepth_var_max = 0.5f * 0.5f;
  p->max_dropouts = 5;
  p->outlier_sigma_thresh = 3.0f;
  p->do_meas_fusion = 1;
  p->do_letterbox = 0;
}

/* cv::Rect_<int>::contains takes a Point_<int> (OpenCV 3.2 core/types.hpp): the Point2f handed to it at
 * flame.cc:1624,1680 is first converted with saturate_cast<int> = cvRound (round half to even), and the
 * ROUNDED point is tested: x <= pt.x < x+width && y <= pt.y < y+height. */
static int rect_contains(int rx, int ry, int rw, int rh, float px, float py) {
  const int ix = (int)lrintf(px), iy = (int)lrintf(py);
  return rx <= ix && ix < rx + rw && ry <= iy && iy < ry + rh;
}

/* flame.cc:1536-1752. Returns 1 on success (flow valid). *stage gets the exit point. */
static int track_feature(const oracle_track_params* P, const float K[9], const float Kinv[9],
                         int width, int height, int pad, const oracle_frame* frames, int fnew,
                         int curr_pf, const oracle_epigeo* epigeo, oracle_feature* feat,
                         float* flow_x, float* flow_y, int* stage) {
  float ucx, ucy, idepth_cmp, var_cmp;
  int pred_success = oracle_filter_predict(epigeo, P->fparams.process_var_factor, feat->x, feat->y,
                                           feat->idepth_mu, feat->idepth_var, &ucx, &ucy,
                                           &idepth_cmp, &var_cmp); /* :1553 */
  if (!pred_success) { *stage = 1; return 0; }

  int row_offset = 0;
  if (P->do_letterbox) row_offset = height / 3;
  int border = (int)(P->rescale_factor_max * P->fparams.win_size / 2 + 1); /* :1576 */
  int vr_x = border, vr_y = border + row_offset;
  int vr_w = width - 2 * border, vr_h = height - 2 * border - 2 * row_offset;

  float rescale_factor = 1.0f;
  if ((feat->idepth_mu > 0.0f) && (idepth_cmp > 0.0f)) rescale_factor = idepth_cmp / feat->idepth_mu;

  if ((rescale_factor <= P->rescale_factor_min) || (rescale_factor >= P->rescale_factor_max)) {
    /* :1604-1659 re-parent the feature to the current pose-frame */
    oracle_epigeo epipf;
    oracle_epigeo_init(&epipf, K, Kinv);
    float T[7];
    oracle_se3_inv_mul(frames[curr_pf].pose, frames[feat->frame_id].pose, T);
    oracle_epigeo_load(&epipf, T, T + 4);
    float upx, upy, idepth_pf, var_pf;
    int move_success = oracle_filter_predict(&epipf, P->fparams.process_var_factor, feat->x,
                                             feat->y, feat->idepth_mu, feat->idepth_var, &upx,
                                             &upy, &idepth_pf, &var_pf);
    if (!move_success || !rect_contains(vr_x, vr_y, vr_w, vr_h, upx, upy)) {
      feat->valid = 0;
      *stage = 3;
      return 0;
    }
    feat->frame_id = (uint32_t)curr_pf;
    feat->x = upx;
    feat->y = upy;
    float old_idepth = feat->idepth_mu;
    feat->idepth_mu = idepth_pf;
    float var_factor4 = idepth_pf / old_idepth;
    var_factor4 *= var_factor4;
    var_factor4 *= var_factor4;
    if ((double)idepth_pf < 1e-6) var_factor4 = 1;
    feat->idepth_var *= var_factor4;
    *stage = 2;
    return 0;
  }

  float sx, sy, ex, ey, epx, epy;
  int region_success = oracle_filter_get_search_region(&P->fparams, epigeo, width, height, feat->x,
                                                       feat->y, feat->idepth_mu, feat->idepth_var,
                                                       &sx, &sy, &ex, &ey, &epx, &epy); /* :1663 */
  if (!region_success) { *stage = 4; return 0; }

  float off = (float)pad; /* :1677-1678 */
  if (!rect_contains(vr_x, vr_y, vr_w, vr_h, feat->x, feat->y)) { *stage = 5; return 0; }

  int prows = height + 2 * pad, pcols = width + 2 * pad;
  int status = oracle_filter_search(&P->fparams, epigeo, rescale_factor,
                                    frames[feat->frame_id].img_pad, frames[fnew].img_pad, prows,
                                    pcols, pcols, feat->x + off, feat->y + off, sx + off, sy + off,
                                    ex + off, ey + off, &ucx, &ucy); /* :1690 */
  feat->search_status = status;
  if (status != ORACLE_SEARCH_SUCCESS) { *stage = 6; return 0; }

  *flow_x = ucx - off;
  *flow_y = ucy - off;
  return 1;
}

/* flame.cc:1344-1376 (repeated at :1403-1431, :1448-1476): failure bookkeeping */
static void feature_fail(const oracle_track_params* P, oracle_feature* f) {
  f->idepth_var *= P->fparams.process_fail_var_factor;
  if (f->idepth_var > P->idepth_var_max) f->valid = 0;
  f->num_dropouts++;
  if (f->num_dropouts > (uint32_t)P->max_dropouts) f->valid = 0;
}

int oracle_update_feature_idepths(const oracle_track_params* P, const float K[9],
                                  const float Kinv[9], int width, int height, int pad,
                                  const oracle_frame* frames, int n_frames, int fnew, int curr_pf,
                                  oracle_feature* feats, int n_feats, oracle_track_diag* diag) {
  (void)n_frames;
  int num_total_updates = 0;
  for (int ii = 0; ii < n_feats; ++ii) { /* :1308 */
    oracle_feature* f = &feats[ii];
    oracle_track_diag dg = {NAN, NAN, NAN, NAN, 0};
    oracle_epigeo epigeo;
    oracle_epigeo_init(&epigeo, K, Kinv);
    float T[7];
    oracle_se3_inv_mul(frames[fnew].pose, frames[f->frame_id].pose, T); /* :1314 */
    oracle_epigeo_load(&epigeo, T, T + 4);

    float baseline = (float)sqrt((double)(T[4] * T[4] + T[5] * T[5] + T[6] * T[6])); /* :1320 */
    if (baseline < P->min_baseline) {
      if (diag) diag[ii] = dg;
      continue;
    }

    float flow_x = NAN, flow_y = NAN;
    int stage = 0;
    uint32_t ref_frame = f->frame_id;
    float ref_x = f->x, ref_y = f->y;
    int track_success = track_feature(P, K, Kinv, width, height, pad, frames, fnew, curr_pf,
                                      &epigeo, f, &flow_x, &flow_y, &stage); /* :1329 */
    if (!track_success) {
      feature_fail(P, f);
      dg.stage = stage;
      if (diag) diag[ii] = dg;
      continue;
    }
    dg.flow_x = flow_x;
    dg.flow_y = flow_y;

    /* :1380-1390: meas model uses pfs.at(fii.frame_id) and fii.xy (unchanged on success) */
    (void)ref_frame; (void)ref_x; (void)ref_y;
    oracle_epigeo mgeo;
    oracle_epigeo_init(&mgeo, K, Kinv);
    oracle_se3_inv_mul(frames[fnew].pose, frames[f->frame_id].pose, T); /* meas_model.h:69 */
    oracle_epigeo_load(&mgeo, T, T + 4);
    float mu_meas, var_meas;
    int prows = height + 2 * pad, pcols = width + 2 * pad;
    int sense_success = oracle_meas_idepth(&P->zparams, &mgeo, frames[fnew].gradx_pad,
                                           frames[fnew].grady_pad, prows, pcols, f->x, f->y, flow_x,
                                           flow_y, &mu_meas, &var_meas);
    dg.mu_meas = mu_meas;
    dg.var_meas = var_meas;
    if (!sense_success) {
      feature_fail(P, f);
      dg.stage = 7;
      if (diag) diag[ii] = dg;
      continue;
    }

    float mu_post, var_post;
    int fuse_success = oracle_filter_update(f->idepth_mu, f->idepth_var, mu_meas, var_meas, &mu_post,
                                            &var_post, P->outlier_sigma_thresh); /* :1436 */
    if (!fuse_success) {
      feature_fail(P, f);
      dg.stage = 8;
      if (diag) diag[ii] = dg;
      continue;
    }
    if (P->do_meas_fusion) {
      f->idepth_mu = mu_post;
      f->idepth_var = var_post;
    } else {
      f->idepth_mu = mu_meas;
      f->idepth_var = var_meas;
    }
    f->valid = 1;
    f->num_updates++;
    f->num_dropouts = 0;
    num_total_updates++;
    dg.stage = 9;
    if (diag) diag[ii] = dg;
  }
  return num_total_updates;
}


/* ====================================================================================
 * Flame::projectGraph — flame.cc:1862-1937 (vertex loop and removal marks)
 * ==================================================================================== */
void oracle_project_graph(const float K[9], const float Kinv[9], const float q_wxyz[4], const float t[3],
                          int width, int height, const oracle_project_rules* rules, float graph_scale, int32_t V,
                          float* pos, float* x, const float* gradx, const float* grady, float* idepth_out,
                          uint8_t* remove) {
  oracle_epigeo g;
  oracle_epigeo_init(&g, K, Kinv);
  oracle_epigeo_load(&g, q_wxyz, t);
  int row_offset = 0;
  if (rules->do_letterbox) row_offset = height / 3;                                  /* :1875-1878 */
  const int border = (int)(rules->rescale_factor_max * rules->win_size / 2 + 1);     /* :1881 */
  /* cv::Rect_<float> valid_region(border, border + row_offset, width - 2*border, height - 2*border - 2*row_offset) */
  const float rx = (float)border, ry = (float)(border + row_offset);
  const float rw = (float)(width - 2 * border), rh = (float)(height - 2 * border - 2 * row_offset);
  for (int32_t v = 0; v < V; ++v) {
    float ux, uy, idepth_new;
    oracle_epigeo_project_idepth(&g, pos[2 * v], pos[2 * v + 1], x[v] * graph_scale, &ux, &uy, &idepth_new);
    pos[2 * v] = ux;
    pos[2 * v + 1] = uy;
    x[v] = idepth_new / graph_scale;
    if (idepth_out) idepth_out[v] = idepth_new;
    const int inside = rx <= ux && ux < rx + rw && ry <= uy && uy < ry + rh;
    int rem = !inside || (idepth_new < 0.0f);                                        /* :1901 */
    if (!rem && rules->do_grad_check_after_projection) {                             /* :1907-1916 */
      const float gx = oracle_bilinear_f32(height, width, width * 4, gradx, ux, uy);
      const float gy = oracle_bilinear_f32(height, width, width * 4, grady, ux, uy);
      if (gx * gx + gy * gy < rules->min_grad_mag * rules->min_grad_mag) rem = 1;
    }
    remove[v] = (uint8_t)rem;
  }
}

/* ====================================================================================
 * Flame::detectFeatures, single-pass grid detector — flame.cc:1192-1277
 * ==================================================================================== */
int32_t oracle_detect_features(const float K[9], const float Kinv[9], const float pose_ref[7],
                               const float pose_prev[7], int width, int height, const oracle_detect_params* p,
                               const float* gradx, const float* grady, const float* curr_xy, int32_t n_curr,
                               float* out_xy, int32_t max_out) {
  int row_offset = 0;
  if (p->do_letterbox) row_offset = height / 3;                                      /* :840-844 */
  const int border = (int)(p->rescale_factor_max * p->win_size / 2 + 1);             /* :848 */
  const int win_size = p->detection_win_size;
  float fh = (float)height / win_size, fw = (float)width / win_size;                 /* fast_ceil, :1195-1196 */
  int hclvl = (int)fh, wclvl = (int)fw;
  if (hclvl < fh) ++hclvl;
  if (wclvl < fw) ++wclvl;
  const size_t cells = (size_t)hclvl * wclvl;
  uint8_t* cmask = (uint8_t*)malloc(cells ? cells : 1);
  memset(cmask, 255, cells);
  for (int32_t i = 0; i < n_curr; ++i) {                                             /* :1199-1204 */
    const uint32_t x_clvl = (uint32_t)(curr_xy[2 * i] / win_size);
    const uint32_t y_clvl = (uint32_t)(curr_xy[2 * i + 1] / win_size);
    if (x_clvl < (uint32_t)wclvl && y_clvl < (uint32_t)hclvl) cmask[(size_t)y_clvl * wclvl + x_clvl] = 0;
  }
  float T[7];
  oracle_se3_inv_mul(pose_prev, pose_ref, T);                                        /* :1207 */
  oracle_epigeo g;
  oracle_epigeo_init(&g, K, Kinv);
  oracle_epigeo_load(&g, T, T + 4);
  float* best = (float*)calloc(cells ? cells : 1, sizeof(float));
  float* best_px = (float*)calloc(2 * (cells ? cells : 1), sizeof(float));
  const float grad_thresh2 = p->min_grad_mag * p->min_grad_mag;
  for (uint32_t ii = (uint32_t)(border + row_offset); ii < (uint32_t)(height - border - row_offset); ++ii) {
    for (uint32_t jj = (uint32_t)border; jj < (uint32_t)(width - border); ++jj) {
      const float gx = gradx[(size_t)ii * width + jj], gy = grady[(size_t)ii * width + jj];
      const float gmag2 = gx * gx + gy * gy;
      if (gmag2 < grad_thresh2) continue;
      float ex, ey;
      oracle_epigeo_reference_epiline(&g, (float)ii, (float)jj, &ex, &ey);           /* sic: Point2f(ii, jj), :1231 */
      const float epigrad = gx * ex + gy * ey;
      const float epigrad2 = epigrad * epigrad;
      if (epigrad2 < grad_thresh2) continue;
      const int ii_clvl = (int)((float)ii / win_size), jj_clvl = (int)((float)jj / win_size);
      const size_t idx = (size_t)ii_clvl * wclvl + jj_clvl;
      if (epigrad2 >= best[idx]) {
        best[idx] = epigrad2;
        best_px[2 * idx] = (float)jj;
        best_px[2 * idx + 1] = (float)ii;
      }
    }
  }
  int32_t n = 0;
  for (size_t idx = 0; idx < cells; ++idx)                                           /* :1262-1270 */
    if (cmask[idx] > 0 && best[idx] > 0) {
      if (n < max_out) {
        out_xy[2 * n] = best_px[2 * idx];
        out_xy[2 * n + 1] = best_px[2 * idx + 1];
      }
      ++n;
    }
  free(cmask);
  free(best);
  free(best_px);
  return n;
}
