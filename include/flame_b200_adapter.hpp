// flame_b200_adapter.hpp — header-only C++11 adapter that offers the reference's own call
// signatures on top of the C ABI (flame_b200.h), so that flame::Flame can switch this path by
// changing one namespace alias and three call sites.
//
// It is written against the CONCEPTS the reference code uses, not against Boost/OpenCV headers:
//   Graph   : boost::adjacency_list<..., VertexData, EdgeData> as declared in
//             src/flame/optimizers/nltgv2_l1_graph_regularizer.h:107-112 (needs boost::vertices,
//             boost::edges, boost::source, boost::target, operator[] — found by ADL)
//   Point2f : anything with float members x, y (cv::Point2f)
//   Triangle: anything indexable with [0..2] (cv::Vec3i)
//   Image   : anything with rows, cols and a float* accessor ptr<float>(0) (cv::Mat / cv::Mat1f)
// so it compiles unchanged inside the reference tree and, for the tests here, against oracle/shim.
//
// Error convention: the reference's functions return void / bool and abort through FLAME_ASSERT on
// programmer errors (src/flame/utils/assert.h:111). The adapter throws std::runtime_error carrying
// flame_b200_last_error() instead of exiting the process.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "flame_b200.h"

namespace flame_b200 {

inline void check(int rc, const char* what) {
  if (rc != 0)
    throw std::runtime_error(std::string(what) + ": " + flame_b200_last_error() + " (" +
                             std::to_string(rc) + ")");
}

// One CUDA stream + device buffers; one per flame::Flame instance (serialise calls on it exactly as
// Flame serialises dgraph::step() under graph_mtx_, flame.cc:100-111).
class Session {
 public:
  explicit Session(int device = 0, int n_slots = 1) : ctx_(nullptr) {
    check(flame_b200_ctx_create(device, n_slots, &ctx_), "flame_b200_ctx_create");
  }
  ~Session() { flame_b200_ctx_destroy(ctx_); }
  Session(const Session&) = delete;
  Session& operator=(const Session&) = delete;
  flame_b200_ctx* get() const { return ctx_; }
  void sync() { check(flame_b200_ctx_sync(ctx_), "flame_b200_ctx_sync"); }

 private:
  flame_b200_ctx* ctx_;
};

namespace optimizers {
namespace nltgv2_l1_graph_regularizer {

// Same fields / defaults as the reference's Params (nltgv2_l1_graph_regularizer.h:121-129).
struct Params {
  float data_factor = 0.1f;
  float step_x = 0.001f;
  float step_q = 125.0f;
  float theta = 0.25f;
  float x_min = 0.0f;
  float x_max = 10.0f;
  template <class RefParams>
  static Params from(const RefParams& r) {
    Params p;
    p.data_factor = r.data_factor; p.step_x = r.step_x; p.step_q = r.step_q;
    p.theta = r.theta; p.x_min = r.x_min; p.x_max = r.x_max;
    return p;
  }
  flame_b200_pd_params c() const {
    flame_b200_pd_params q = {data_factor, step_x, step_q, theta, x_min, x_max};
    return q;
  }
};

// Device-resident image of a reference Graph. upload() flattens the graph in the canonical order
// of SURVEY N3 (vertices: boost::vertices order; edges: boost::edges order, which for the
// reference's undirected adjacency_list is insertion order = Triangle's edgelist order), step()
// replaces n calls of dgraph::step(params, &graph), download() writes x, w, x_bar, x_prev, q back.
template <class Graph>
class DeviceGraph {
 public:
  DeviceGraph(Session* s, int slot = 0) : s_(s), slot_(slot), V_(0), E_(0) {}

  void upload(const Graph& g) {
    vh_.clear();
    eh_.clear();
    index_.clear();
    auto vr = vertices(g);
    for (auto it = vr.first; it != vr.second; ++it) {
      index_[*it] = static_cast<int32_t>(vh_.size());
      vh_.push_back(*it);
    }
    auto er = edges(g);
    for (auto it = er.first; it != er.second; ++it) eh_.push_back(*it);
    V_ = static_cast<int32_t>(vh_.size());
    E_ = static_cast<int32_t>(eh_.size());
    buf_.assign(static_cast<size_t>(13) * V_ + static_cast<size_t>(5) * E_, 0.0f);
    ebuf_.assign(static_cast<size_t>(2) * E_, 0);
    float* pos = buf_.data();
    float* f[11];
    for (int k = 0; k < 11; ++k) f[k] = pos + 2 * static_cast<size_t>(V_) + static_cast<size_t>(k) * V_;
    for (int32_t v = 0; v < V_; ++v) {
      const auto& d = g[vh_[v]];
      pos[2 * v] = d.pos.x; pos[2 * v + 1] = d.pos.y;
      f[0][v] = d.x; f[1][v] = d.w1; f[2][v] = d.w2;
      f[3][v] = d.x_bar; f[4][v] = d.w1_bar; f[5][v] = d.w2_bar;
      f[6][v] = d.x_prev; f[7][v] = d.w1_prev; f[8][v] = d.w2_prev;
      f[9][v] = d.data_term; f[10][v] = d.data_weight;
    }
    float* e0 = pos + 13 * static_cast<size_t>(V_);
    for (int32_t e = 0; e < E_; ++e) {
      const auto& d = g[eh_[e]];
      ebuf_[2 * e] = index_[source(eh_[e], g)];
      ebuf_[2 * e + 1] = index_[target(eh_[e], g)];
      e0[e] = d.alpha; e0[E_ + e] = d.beta;
      e0[2 * static_cast<size_t>(E_) + e] = d.q1;
      e0[3 * static_cast<size_t>(E_) + e] = d.q2;
      e0[4 * static_cast<size_t>(E_) + e] = d.q3;
    }
    flame_b200_graph_in in;
    in.V = V_; in.E = E_;
    in.pos = pos; in.edges = ebuf_.data();
    in.x = f[0]; in.w1 = f[1]; in.w2 = f[2];
    in.x_bar = f[3]; in.w1_bar = f[4]; in.w2_bar = f[5];
    in.x_prev = f[6]; in.w1_prev = f[7]; in.w2_prev = f[8];
    in.data_term = f[9]; in.data_weight = f[10];
    in.alpha = e0; in.beta = e0 + E_;
    in.q1 = e0 + 2 * static_cast<size_t>(E_); in.q2 = e0 + 3 * static_cast<size_t>(E_);
    in.q3 = e0 + 4 * static_cast<size_t>(E_);
    check(flame_b200_graph_upload(s_->get(), slot_, &in), "flame_b200_graph_upload");
  }

  // Flame::projectGraph's vertex loop (flame.cc:1892-1899) on the device, for the graph last passed to
  // upload()/resync(): pos and x of every vertex are also written back into the Boost graph so that the
  // removal rules (:1901-1917) and syncGraph run on the host as before. (q_wxyz, t) = T_prev_to_new
  // (flame.cc:297-298). Needs flame_b200_set_camera on the session.
  void project(const float q_wxyz[4], const float t[3], float graph_scale, Graph* g) {
    std::vector<float> pos(static_cast<size_t>(2) * V_ + 2), idepth(static_cast<size_t>(V_) + 1);
    check(flame_b200_graph_project(s_->get(), slot_, q_wxyz, t, graph_scale, pos.data(), idepth.data()),
          "flame_b200_graph_project");
    for (int32_t v = 0; v < V_; ++v) {
      auto& d = (*g)[vh_[v]];
      d.pos.x = pos[2 * v];
      d.pos.y = pos[2 * v + 1];
      d.x = idepth[v] / graph_scale;
    }
  }

  // The same including the REMOVAL RULES of Flame::projectGraph (flame.cc:1901-1917) evaluated on the device:
  // returns, in boost::vertices order of the last upload()/resync(), whether the reference would remove the
  // vertex (outside valid_region, behind the camera, or - do_grad_check_after_projection - on a gradient-free
  // pixel of the new frame in frame slot `fnew_slot`). The caller then does what :1920-1928 does for the marked
  // vertices (clear_vertex / remove_vertex / erase from the feature maps). `RefParams` = flame::Params.
  template <class RefParams>
  std::vector<uint8_t> projectMarked(const RefParams& params, const float q_wxyz[4], const float t[3],
                                     float graph_scale, int fnew_slot, Graph* g) {
    flame_b200_project_rules r;
    r.rescale_factor_max = params.rescale_factor_max;
    r.win_size = params.fparams.win_size;
    r.do_letterbox = params.do_letterbox ? 1 : 0;
    r.do_grad_check_after_projection = params.do_grad_check_after_projection ? 1 : 0;
    r.min_grad_mag = params.min_grad_mag;
    std::vector<float> pos(static_cast<size_t>(2) * V_ + 2), idepth(static_cast<size_t>(V_) + 1);
    std::vector<uint8_t> remove(static_cast<size_t>(V_) + 1, 0);
    check(flame_b200_graph_project_marked(s_->get(), slot_, q_wxyz, t, graph_scale, &r, fnew_slot, pos.data(),
                                          idepth.data(), remove.data()),
          "flame_b200_graph_project_marked");
    for (int32_t v = 0; v < V_; ++v) {
      auto& d = (*g)[vh_[v]];
      d.pos.x = pos[2 * v];
      d.pos.y = pos[2 * v + 1];
      d.x = idepth[v] / graph_scale;
    }
    remove.resize(static_cast<size_t>(V_));
    return remove;
  }

  // The vertex handle at position v of the flat order (to act on projectMarked's result).
  typename std::decay<decltype(*vertices(*static_cast<const Graph*>(nullptr)).first)>::type vertex(int32_t v) const {
    return vh_[v];
  }

  // After Flame::syncGraph edited the Boost graph (vertices removed / added, edges re-triangulated,
  // flame.cc:1984-2162): rebuild the device graph from it and carry the solver state of the surviving
  // vertices and edges over ON THE DEVICE (no download / upload of x, w, q). `is_new(vertex)` must say
  // whether the vertex was added by this syncGraph call (:2030-2049) - handle identity alone cannot,
  // a list node address may be reused. A surviving edge is recognised by its (source, target) pair,
  // which Boost keeps stable for an undirected edge that was not removed. New vertices take x = x_bar =
  // x_prev from their VertexData::x (what :2160-2162 stored).
  template <class IsNew>
  void resync(const Graph& g, IsNew is_new, bool check_sticky_obstacles = false) {
    std::unordered_map<typename std::decay<VH>::type, int32_t> old_index;
    old_index.swap(index_);
    std::unordered_map<uint64_t, int32_t> old_edge;
    for (int32_t e = 0; e < E_; ++e)
      old_edge[(static_cast<uint64_t>(static_cast<uint32_t>(ebuf_[2 * e])) << 32) |
               static_cast<uint32_t>(ebuf_[2 * e + 1])] = e;
    vh_.clear();
    eh_.clear();
    std::vector<int32_t> vtx_from, edge_from;
    auto vr = vertices(g);
    for (auto it = vr.first; it != vr.second; ++it) {
      index_[*it] = static_cast<int32_t>(vh_.size());
      vh_.push_back(*it);
      auto f = old_index.find(*it);
      vtx_from.push_back((f == old_index.end() || is_new(*it)) ? -1 : f->second);
    }
    auto er = edges(g);
    for (auto it = er.first; it != er.second; ++it) eh_.push_back(*it);
    V_ = static_cast<int32_t>(vh_.size());
    E_ = static_cast<int32_t>(eh_.size());
    buf_.assign(static_cast<size_t>(13) * V_ + static_cast<size_t>(5) * E_, 0.0f);
    ebuf_.assign(static_cast<size_t>(2) * E_, 0);
    float* pos = buf_.data();
    float* data = pos + 2 * static_cast<size_t>(V_);
    float* wt = data + V_;
    float* xi = wt + V_;
    for (int32_t v = 0; v < V_; ++v) {
      const auto& d = g[vh_[v]];
      pos[2 * v] = d.pos.x; pos[2 * v + 1] = d.pos.y;
      data[v] = d.data_term; wt[v] = d.data_weight; xi[v] = d.x;
    }
    edge_from.assign(E_, -1);
    for (int32_t e = 0; e < E_; ++e) {
      const int32_t a = index_[source(eh_[e], g)], b = index_[target(eh_[e], g)];
      ebuf_[2 * e] = a;
      ebuf_[2 * e + 1] = b;
      if (vtx_from[a] >= 0 && vtx_from[b] >= 0) {
        auto f = old_edge.find((static_cast<uint64_t>(static_cast<uint32_t>(vtx_from[a])) << 32) |
                               static_cast<uint32_t>(vtx_from[b]));
        if (f != old_edge.end()) edge_from[e] = f->second;
      }
    }
    flame_b200_graph_sync y;
    y.V = V_; y.E = E_;
    y.pos = pos; y.data_term = data; y.data_weight = wt;
    y.vtx_from = vtx_from.data(); y.x_init = xi;
    y.edges = ebuf_.data(); y.edge_from = edge_from.data();
    y.check_sticky_obstacles = check_sticky_obstacles ? 1 : 0;
    check(flame_b200_graph_resync(s_->get(), slot_, &y), "flame_b200_graph_resync");
  }

  void step(const Params& params, int n_iters = 1, int mode = FLAME_B200_PD_AUTO) {
    const flame_b200_pd_params p = params.c();
    check(flame_b200_pd_run(s_->get(), slot_, 1, &p, n_iters, mode), "flame_b200_pd_run");
  }

  void download(Graph* g) {
    float* pos = buf_.data();
    float* f[9];
    for (int k = 0; k < 9; ++k) f[k] = pos + 2 * static_cast<size_t>(V_) + static_cast<size_t>(k) * V_;
    float* e0 = pos + 13 * static_cast<size_t>(V_);
    flame_b200_graph_out out;
    out.x = f[0]; out.w1 = f[1]; out.w2 = f[2];
    out.x_bar = f[3]; out.w1_bar = f[4]; out.w2_bar = f[5];
    out.x_prev = f[6]; out.w1_prev = f[7]; out.w2_prev = f[8];
    out.alpha = nullptr; out.beta = nullptr;
    out.q1 = e0 + 2 * static_cast<size_t>(E_); out.q2 = e0 + 3 * static_cast<size_t>(E_);
    out.q3 = e0 + 4 * static_cast<size_t>(E_);
    check(flame_b200_graph_download(s_->get(), slot_, &out), "flame_b200_graph_download");
    for (int32_t v = 0; v < V_; ++v) {
      auto& d = (*g)[vh_[v]];
      d.x = f[0][v]; d.w1 = f[1][v]; d.w2 = f[2][v];
      d.x_bar = f[3][v]; d.w1_bar = f[4][v]; d.w2_bar = f[5][v];
      d.x_prev = f[6][v]; d.w1_prev = f[7][v]; d.w2_prev = f[8][v];
    }
    for (int32_t e = 0; e < E_; ++e) {
      auto& d = (*g)[eh_[e]];
      d.q1 = out.q1[e]; d.q2 = out.q2[e]; d.q3 = out.q3[e];
    }
  }

  float smoothnessCost(const Params& params, bool exact_order = true) {
    float s = 0.f, d = 0.f;
    const flame_b200_pd_params p = params.c();
    check(flame_b200_pd_cost(s_->get(), slot_, &p, exact_order ? 1 : 0, &s, &d), "flame_b200_pd_cost");
    return s;
  }
  float dataCost(const Params& params, bool exact_order = true) {
    float s = 0.f, d = 0.f;
    const flame_b200_pd_params p = params.c();
    check(flame_b200_pd_cost(s_->get(), slot_, &p, exact_order ? 1 : 0, &s, &d), "flame_b200_pd_cost");
    return d;
  }
  int32_t num_vertices() const { return V_; }
  int32_t num_edges() const { return E_; }

 private:
  typedef decltype(*vertices(*static_cast<const Graph*>(nullptr)).first) VH;
  typedef decltype(*edges(*static_cast<const Graph*>(nullptr)).first) EH;
  Session* s_;
  int slot_;
  int32_t V_, E_;
  std::vector<typename std::decay<VH>::type> vh_;
  std::vector<typename std::decay<EH>::type> eh_;
  std::unordered_map<typename std::decay<VH>::type, int32_t> index_;
  std::vector<float> buf_;
  std::vector<int32_t> ebuf_;
};

// Drop-in for `void step(const Params&, Graph*)` (nltgv2_l1_graph_regularizer.h:134): one full
// upload -> n_iters iterations -> download round trip. Prefer a persistent DeviceGraph in the
// optimiser thread (upload after syncGraph, step in the loop, download before the reads at
// flame.cc:372-380); this overload exists so the reference's call site compiles unchanged.
template <class RefParams, class Graph>
void step(Session* s, const RefParams& params, Graph* graph, int n_iters = 1) {
  DeviceGraph<Graph> dg(s);
  dg.upload(*graph);
  dg.step(Params::from(params), n_iters);
  dg.download(graph);
}

}  // namespace nltgv2_l1_graph_regularizer
}  // namespace optimizers

namespace utils {

// Drop-in for utils::interpolateMesh (src/flame/utils/image_utils.h:646-651): `img` must be a
// continuous float image, pre-initialised by the caller (NaN in Flame::update, flame.cc:412).
template <class Triangle, class Point2f, class Image>
void interpolateMesh(Session* s, const std::vector<Triangle>& triangles,
                     const std::vector<Point2f>& vertices, const std::vector<float>& values,
                     const std::vector<bool>& vtx_validity, const std::vector<bool>& tri_validity,
                     Image* img, int slot = 0) {
  const int32_t T = static_cast<int32_t>(triangles.size());
  const int32_t V = static_cast<int32_t>(vertices.size());
  std::vector<int32_t> tris(static_cast<size_t>(3) * T);
  std::vector<float> pos(static_cast<size_t>(2) * V);
  std::vector<uint8_t> vv(V), tv(T);
  for (int32_t t = 0; t < T; ++t) {
    tris[3 * t] = triangles[t][0]; tris[3 * t + 1] = triangles[t][1]; tris[3 * t + 2] = triangles[t][2];
    tv[t] = tri_validity[t] ? 1 : 0;
  }
  for (int32_t v = 0; v < V; ++v) {
    pos[2 * v] = vertices[v].x; pos[2 * v + 1] = vertices[v].y;
    vv[v] = vtx_validity[v] ? 1 : 0;
  }
  check(flame_b200_interpolate_mesh(s->get(), slot, T, tris.data(), V, pos.data(), values.data(), 1.0f,
                                    vv.data(), tv.data(), img->rows, img->cols, /*fill_nan=*/0,
                                    img->template ptr<float>(0)),
        "flame_b200_interpolate_mesh");
}

}  // namespace utils

namespace stereo {

// Drop-in for the loop body of Flame::updateFeatureIDepths (flame.cc:1308-1496): `feats` is any
// contiguous array of records layout-compatible with flame_b200_feature (flame::FeatureWithIDepth
// has the same members in the same order; copy field-wise if your ABI pads `bool valid`).
inline int updateFeatureIDepths(Session* s, const flame_b200_track_params& params, int fnew_slot,
                                int curr_pf_slot, flame_b200_feature* feats, int32_t n,
                                flame_b200_track_diag* diag = nullptr) {
  int32_t n_updates = 0;
  check(flame_b200_update_feature_idepths(s->get(), &params, fnew_slot, curr_pf_slot, feats, n, diag,
                                          &n_updates),
        "flame_b200_update_feature_idepths");
  return n_updates;
}

}  // namespace stereo

}  // namespace flame_b200
