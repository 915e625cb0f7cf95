// C entry points over the reference's display filters and vertex normals (test infrastructure;
// oracle/_ref only). The four member functions are sliced VERBATIM out of the reference's flame.cc by
// oracle/ref_capi/slice_flame.py (flame.cc itself needs the full OpenCV/Boost/Sophus APIs) and compiled
// against oracle/shim; this file declares just enough of class Flame around them and adapts flat arrays.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "flame/params.h"
#include "flame/types.h"
#include "flame/utils/assert.h"
#include "flame/utils/stats_tracker.h"

namespace flame {

class Flame {  // the static members of flame::Flame (flame.h:405-475) that the slices define
 public:
  static void obliqueTriangleFilter(const Params& params, const Matrix3f& Kinv,
                                    const std::vector<cv::Point2f>& vertices, const std::vector<float>& idepths,
                                    const std::vector<Triangle>& triangles, std::vector<bool>* validity,
                                    utils::StatsTracker* stats);
  static void edgeLengthFilter(const Params& params, int width, int height,
                               const std::vector<cv::Point2f>& vertices, const std::vector<Triangle>& triangles,
                               std::vector<bool>* validity, utils::StatsTracker* stats);
  static void idepthTriangleFilter(const Params& params, const std::vector<float>& idepths,
                                   const std::vector<Triangle>& triangles, std::vector<bool>* validity,
                                   utils::StatsTracker* stats);
  static void getVertexNormals(const Params& params, const Matrix3f& Kinv, const std::vector<cv::Point2f>& vtx,
                               const std::vector<float>& idepths, const std::vector<Triangle>& triangles,
                               std::vector<Vector3f>* normals, utils::StatsTracker* stats);
};

#include "flame_slices_filters.inc"

}  // namespace flame

struct ref_tri_filter_params {  // same layout as oracle_tri_filter_params / flame_b200_tri_filter_params
  int32_t do_oblique_triangle_filter;
  float oblique_normal_thresh, oblique_idepth_diff_factor, oblique_idepth_diff_abs;
  int32_t do_edge_length_filter;
  float edge_length_thresh;
  int32_t do_idepth_triangle_filter;
  float min_triangle_idepth;
};

static flame::Matrix3f to_mat(const float K[9]) {
  flame::Matrix3f m;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) m(i, j) = K[3 * i + j];
  return m;
}

static void unflatten(int32_t V, const float* pos, int32_t T, const int32_t* tris,
                      std::vector<cv::Point2f>* vtx, std::vector<flame::Triangle>* triangles) {
  vtx->resize(V);
  for (int v = 0; v < V; ++v) (*vtx)[v] = cv::Point2f(pos[2 * v], pos[2 * v + 1]);
  triangles->resize(T);
  for (int t = 0; t < T; ++t) (*triangles)[t] = flame::Triangle(tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]);
}

// The filters as Flame::update combines them (flame.cc:389-407).
extern "C" void ref_triangle_filters(const ref_tri_filter_params* p, const float Kinv[9], int width, int height,
                                     int32_t V, int32_t T, const int32_t* tris, const float* pos,
                                     const float* idepths, uint8_t* validity) {
  flame::Params params;
  params.debug_quiet = true;
  params.oblique_normal_thresh = p->oblique_normal_thresh;
  params.oblique_idepth_diff_factor = p->oblique_idepth_diff_factor;
  params.oblique_idepth_diff_abs = p->oblique_idepth_diff_abs;
  params.edge_length_thresh = p->edge_length_thresh;
  params.min_triangle_idepth = p->min_triangle_idepth;
  std::vector<cv::Point2f> vtx;
  std::vector<flame::Triangle> triangles;
  unflatten(V, pos, T, tris, &vtx, &triangles);
  std::vector<float> id(idepths, idepths + V);
  flame::utils::StatsTracker stats;
  std::vector<bool> valid;
  valid.assign(triangles.size(), true);
  if (p->do_oblique_triangle_filter)
    flame::Flame::obliqueTriangleFilter(params, to_mat(Kinv), vtx, id, triangles, &valid, &stats);
  if (p->do_edge_length_filter)
    flame::Flame::edgeLengthFilter(params, width, height, vtx, triangles, &valid, &stats);
  if (p->do_idepth_triangle_filter) flame::Flame::idepthTriangleFilter(params, id, triangles, &valid, &stats);
  for (int t = 0; t < T; ++t) validity[t] = valid[t] ? 1 : 0;
}

extern "C" void ref_vertex_normals(const float Kinv[9], int32_t V, const float* pos, const float* idepths,
                                   int32_t T, const int32_t* tris, float* normals) {
  flame::Params params;
  params.debug_quiet = true;
  std::vector<cv::Point2f> vtx;
  std::vector<flame::Triangle> triangles;
  unflatten(V, pos, T, tris, &vtx, &triangles);
  std::vector<float> id(idepths, idepths + V);
  flame::utils::StatsTracker stats;
  std::vector<flame::Vector3f> n;
  flame::Flame::getVertexNormals(params, to_mat(Kinv), vtx, id, triangles, &n, &stats);
  for (int v = 0; v < V; ++v)
    for (int k = 0; k < 3; ++k) normals[3 * v + k] = n[v](k);
}
