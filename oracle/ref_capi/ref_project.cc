// C entry point over the reference's Flame::projectGraph (test infrastructure; oracle/_ref only). The member
// function is sliced VERBATIM out of flame.cc by oracle/ref_capi/slice_flame.py and compiled against
// oracle/shim (Boost.Graph / OpenCV / Eigen / Sophus stand-ins); this file declares just enough of class
// Flame around it and adapts flat arrays: it builds a graph of V isolated vertices, runs projectGraph, and
// reports which vertices the function removed.
#include <cstdint>
#include <cstdio>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "flame/optimizers/nltgv2_l1_graph_regularizer.h"
#include "flame/params.h"
#include "flame/stereo/epipolar_geometry.h"
#include "flame/types.h"
#include "flame/utils/assert.h"
#include "flame/utils/frame.h"
#include "flame/utils/image_utils.h"
#include "flame/utils/stats_tracker.h"

namespace flame {

using Graph = optimizers::nltgv2_l1_graph_regularizer::Graph;                // flame.h:54-64
using VertexHandle = optimizers::nltgv2_l1_graph_regularizer::VertexHandle;
using FeatureToVtx = std::unordered_map<uint32_t, VertexHandle>;
using VtxToFeature = std::unordered_map<VertexHandle, uint32_t>;

class Flame {
 public:
  static void projectGraph(const Params& params, const stereo::EpipolarGeometry<float>& epigeo,
                           const utils::Frame& fnew, Graph* graph, float graph_scale, FeatureToVtx* feat_to_vtx,
                           VtxToFeature* vtx_to_feat, utils::StatsTracker* stats);
};

#include "flame_slices_project.inc"

}  // namespace flame

struct ref_project_rules {
  float rescale_factor_max;
  int32_t win_size, do_letterbox, do_grad_check_after_projection;
  float min_grad_mag;
};

extern "C" void ref_project_graph(const float K[9], const float Kinv[9], const float q[4], const float t[3],
                                  int width, int height, const ref_project_rules* r, float graph_scale, int32_t V,
                                  float* pos, float* x, const float* gradx, const float* grady, uint8_t* remove) {
  using namespace flame;
  Params params;
  params.debug_quiet = true;
  params.rescale_factor_max = r->rescale_factor_max;
  params.fparams.win_size = r->win_size;
  params.do_letterbox = r->do_letterbox != 0;
  params.do_grad_check_after_projection = r->do_grad_check_after_projection != 0;
  params.min_grad_mag = r->min_grad_mag;
  Matrix3f Km, Kinvm;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Km(i, j) = K[3 * i + j];
      Kinvm(i, j) = Kinv[3 * i + j];
    }
  stereo::EpipolarGeometry<float> epigeo(Km, Kinvm);
  epigeo.loadGeometry(Eigen::Quaternionf(q[0], q[1], q[2], q[3]), Vector3f(t[0], t[1], t[2]));
  utils::Frame fnew(1);
  fnew.img[0] = Image1b(height, width);  // only its size is read (flame.cc:1872-1873)
  fnew.gradx[0] = Image1f(height, width, const_cast<float*>(gradx));
  fnew.grady[0] = Image1f(height, width, const_cast<float*>(grady));
  Graph graph;
  FeatureToVtx feat_to_vtx;
  VtxToFeature vtx_to_feat;
  std::vector<VertexHandle> handles(V);
  for (int v = 0; v < V; ++v) {
    handles[v] = boost::add_vertex(optimizers::nltgv2_l1_graph_regularizer::VertexData(), graph);
    graph[handles[v]].pos = cv::Point2f(pos[2 * v], pos[2 * v + 1]);
    graph[handles[v]].x = x[v];
    feat_to_vtx[v] = handles[v];
    vtx_to_feat[handles[v]] = v;
  }
  utils::StatsTracker stats;
  Flame::projectGraph(params, epigeo, fnew, &graph, graph_scale, &feat_to_vtx, &vtx_to_feat, &stats);
  for (int v = 0; v < V; ++v) {
    remove[v] = feat_to_vtx.count(v) ? 0 : 1;
    if (!remove[v]) {  // removed vertices are gone: their projected state cannot be read back
      pos[2 * v] = graph[handles[v]].pos.x;
      pos[2 * v + 1] = graph[handles[v]].pos.y;
      x[v] = graph[handles[v]].x;
    }
  }
}
