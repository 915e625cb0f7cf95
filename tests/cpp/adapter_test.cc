// GPU test of include/flame_b200_adapter.hpp against the reference's own code: the reference's
// Graph type and dgraph::step (nltgv2_l1_graph_regularizer.cc compiled verbatim, oracle/_ref/obj)
// vs flame_b200::optimizers::nltgv2_l1_graph_regularizer::DeviceGraph / step on the same graph;
// utils::DrawShadedTriangleBarycentric (rasterization.cc verbatim) vs flame_b200::utils::interpolateMesh.
// Built by tests/cpp/build.sh when /root/reference is present; the binary travels to the GPU box.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <unordered_map>
#include <vector>

#include "flame/optimizers/nltgv2_l1_graph_regularizer.h"
#include "flame/utils/rasterization.h"
#include "flame_b200_adapter.hpp"

namespace dgraph = flame::optimizers::nltgv2_l1_graph_regularizer;

static bool same_bits(float a, float b) { return std::memcmp(&a, &b, 4) == 0 || (std::isnan(a) && std::isnan(b)); }

int main() {
  const int n = 40;  // 1600 vertices, ~4600 edges
  dgraph::Graph gref, ggpu;
  std::vector<dgraph::VertexHandle> vr, vg;
  unsigned seed = 12345u;
  auto rnd = [&seed]() { seed = seed * 1664525u + 1013904223u; return (seed >> 8) * (1.0f / 16777216.0f); };
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      dgraph::VertexData d;
      d.pos = cv::Point2f(10.0f + 15.0f * x + 5.0f * rnd(), 8.0f + 11.0f * y + 5.0f * rnd());
      d.data_term = 0.5f + 0.002f * x + ((x > 20 && y > 10) ? 0.3f : 0.0f) + 0.02f * rnd();
      d.x = d.x_bar = d.x_prev = d.data_term;
      d.data_weight = 0.5f + rnd();
      vr.push_back(boost::add_vertex(d, gref));
      vg.push_back(boost::add_vertex(d, ggpu));
    }
  auto add = [&](int a, int b) {
    dgraph::EdgeData e;
    cv::Point2f diff = gref[vr[a]].pos - gref[vr[b]].pos;
    e.alpha = 1.0f / std::sqrt(diff.x * diff.x + diff.y * diff.y);
    e.beta = 1.0f;
    boost::add_edge(vr[a], vr[b], e, gref);
    boost::add_edge(vg[a], vg[b], e, ggpu);
  };
  for (int y = 0; y < n; ++y)
    for (int x = 0; x < n; ++x) {
      const int i = y * n + x;
      if (x + 1 < n) add(i, i + 1);
      if (y + 1 < n) add(i + n, i);           // mixed orientations on purpose
      if (x + 1 < n && y + 1 < n) add(i, i + n + 1);
    }
  dgraph::Params p;
  const int iters = 60;
  for (int it = 0; it < iters; ++it) dgraph::step(p, &gref);

  flame_b200::Session sess(0, 1);
  namespace fb = flame_b200::optimizers::nltgv2_l1_graph_regularizer;
  fb::DeviceGraph<dgraph::Graph> dg(&sess);
  dg.upload(ggpu);
  dg.step(fb::Params::from(p), iters);
  const float sc = dg.smoothnessCost(fb::Params::from(p)), dc = dg.dataCost(fb::Params::from(p));
  dg.download(&ggpu);
  int bad = 0;
  for (size_t i = 0; i < vr.size(); ++i) {
    const dgraph::VertexData &a = gref[vr[i]], &b = ggpu[vg[i]];
    bad += !same_bits(a.x, b.x) + !same_bits(a.w1, b.w1) + !same_bits(a.w2, b.w2) +
           !same_bits(a.x_bar, b.x_bar) + !same_bits(a.w1_bar, b.w1_bar) + !same_bits(a.x_prev, b.x_prev);
  }
  auto er = boost::edges(gref);
  auto eg = boost::edges(ggpu);
  for (auto i = er.first, j = eg.first; i != er.second; ++i, ++j)
    bad += !same_bits(gref[*i].q1, ggpu[*j].q1) + !same_bits(gref[*i].q2, ggpu[*j].q2) +
           !same_bits(gref[*i].q3, ggpu[*j].q3);
  const float sr = dgraph::smoothnessCost(p, gref), dr = dgraph::dataCost(p, gref);
  bad += !same_bits(sr, sc) + !same_bits(dr, dc);
  std::printf("PD adapter: V=%zu E=%zu iters=%d mismatches=%d cost ref(%g,%g) gpu(%g,%g)\n", vr.size(),
              boost::num_edges(gref), iters, bad, sr, dr, sc, dc);

  // ---- a graph sync (Flame::syncGraph's edits, flame.cc:1984-2123) replayed on both graphs: the
  // reference keeps stepping its Boost graph; the adapter carries x/w/q over on the device.
  int bad_sync = 0;
  {
    std::vector<char> fresh_vertex(vr.size(), 0), gone(vr.size(), 0);
    for (size_t i = 7; i < vr.size(); i += 13) {           // remove ~8% of the vertices
      boost::clear_vertex(vr[i], gref); boost::remove_vertex(vr[i], gref);
      boost::clear_vertex(vg[i], ggpu); boost::remove_vertex(vg[i], ggpu);
      gone[i] = 1;
    }
    for (size_t i = 0; i < vr.size(); ++i) {               // survivors move, get new data
      if (gone[i]) continue;
      const float dx = rnd() - 0.5f, dy = rnd() - 0.5f, dd = 0.01f * (rnd() - 0.5f);
      for (int k = 0; k < 2; ++k) {
        dgraph::VertexData& d = k ? ggpu[vg[i]] : gref[vr[i]];
        d.pos.x += dx; d.pos.y += dy; d.data_term += dd;
      }
      // the host copy of the solver state is stale by design: prove the device copy is what is used
      ggpu[vg[i]].x = 123.0f; ggpu[vg[i]].w1 = -7.0f; ggpu[vg[i]].x_bar = 55.0f;
    }
    for (int k = 0; k < 60; ++k) {                         // new vertices (x = x_bar = x_prev = data)
      dgraph::VertexData d;
      d.pos = cv::Point2f(12.0f + 590.0f * rnd(), 9.0f + 430.0f * rnd());
      d.data_term = 0.5f + 0.3f * rnd();
      d.x = d.x_bar = d.x_prev = d.data_term;
      d.data_weight = 1.0f;
      vr.push_back(boost::add_vertex(d, gref));
      vg.push_back(boost::add_vertex(d, ggpu));
      fresh_vertex.push_back(1); gone.push_back(0);
    }
    for (size_t i = 3; i + 41 < vr.size(); i += 5) {       // re-triangulation: drop / add edges
      const size_t j = i + 41;
      if (gone[i] || gone[j]) continue;
      if (boost::edge(vr[i], vr[j], gref).second) continue;
      boost::add_edge(vr[i], vr[j], dgraph::EdgeData(), gref);
      boost::add_edge(vg[i], vg[j], dgraph::EdgeData(), ggpu);
    }
    for (size_t i = 0; i + 1 < vr.size(); i += 9) {
      if (gone[i] || gone[i + 1]) continue;
      auto a = boost::edge(vr[i], vr[i + 1], gref);
      auto b = boost::edge(vg[i], vg[i + 1], ggpu);
      if (a.second) { boost::remove_edge(a.first, gref); boost::remove_edge(b.first, ggpu); }
    }
    for (size_t i = vr.size() - 60; i < vr.size(); ++i)    // connect the new vertices
      for (size_t j : {i / 3, i / 2, i - 1}) {
        if (gone[j] || boost::edge(vr[i], vr[j], gref).second) continue;
        boost::add_edge(vr[i], vr[j], dgraph::EdgeData(), gref);
        boost::add_edge(vg[i], vg[j], dgraph::EdgeData(), ggpu);
      }
    auto eit = boost::edges(gref);
    for (auto i = eit.first; i != eit.second; ++i) {       // flame.cc:2089-2104, host path only
      cv::Point2f diff = gref[boost::source(*i, gref)].pos - gref[boost::target(*i, gref)].pos;
      gref[*i].alpha = 1.0f / std::sqrt(diff.x * diff.x + diff.y * diff.y);
      gref[*i].beta = 1.0f;
    }
    for (auto i = boost::edges(ggpu).first; i != boost::edges(ggpu).second; ++i) ggpu[*i].q1 = 99.0f;
    std::unordered_map<dgraph::VertexHandle, char> is_new;
    for (size_t i = 0; i < vg.size(); ++i)
      if (!gone[i]) is_new[vg[i]] = fresh_vertex[i];
    dg.resync(ggpu, [&](dgraph::VertexHandle v) { return is_new[v] != 0; });
    const int more = 35;
    for (int it = 0; it < more; ++it) dgraph::step(p, &gref);
    dg.step(fb::Params::from(p), more);
    dg.download(&ggpu);
    size_t nv = 0;
    for (size_t i = 0; i < vr.size(); ++i) {
      if (gone[i]) continue;
      ++nv;
      const dgraph::VertexData &a = gref[vr[i]], &b = ggpu[vg[i]];
      bad_sync += !same_bits(a.x, b.x) + !same_bits(a.w1, b.w1) + !same_bits(a.w2, b.w2) +
                  !same_bits(a.x_bar, b.x_bar) + !same_bits(a.w2_bar, b.w2_bar) + !same_bits(a.x_prev, b.x_prev);
    }
    auto e1 = boost::edges(gref);
    auto e2 = boost::edges(ggpu);
    for (auto i = e1.first, j = e2.first; i != e1.second; ++i, ++j)
      bad_sync += !same_bits(gref[*i].q1, ggpu[*j].q1) + !same_bits(gref[*i].q2, ggpu[*j].q2) +
                  !same_bits(gref[*i].q3, ggpu[*j].q3);
    bad_sync += !same_bits(dgraph::smoothnessCost(p, gref), dg.smoothnessCost(fb::Params::from(p)));
    std::printf("graph resync adapter: V=%zu E=%zu mismatches=%d\n", nv, boost::num_edges(gref), bad_sync);
  }

  // the one-shot drop-in overload
  dgraph::Graph g2;
  {
    std::vector<dgraph::VertexHandle> h;
    for (int i = 0; i < 3; ++i) {
      dgraph::VertexData d;
      d.pos = cv::Point2f(10.f * i, 3.f * i * i);
      d.x = d.x_bar = d.x_prev = d.data_term = 0.4f + 0.1f * i;
      h.push_back(boost::add_vertex(d, g2));
    }
    dgraph::EdgeData e;
    e.alpha = 0.1f;
    boost::add_edge(h[0], h[1], e, g2);
    boost::add_edge(h[2], h[1], e, g2);
  }
  fb::step(&sess, p, &g2, 5);

  // dense projection
  const int rows = 120, cols = 160;
  std::vector<cv::Vec3i> tris;
  std::vector<cv::Point2f> pts;
  std::vector<float> vals;
  for (int y = 0; y < 6; ++y)
    for (int x = 0; x < 8; ++x) {
      pts.push_back(cv::Point2f(5.f + 21.f * x + 4.f * rnd(), 4.f + 22.f * y + 4.f * rnd()));
      vals.push_back(0.3f + rnd());
    }
  for (int y = 0; y + 1 < 6; ++y)
    for (int x = 0; x + 1 < 8; ++x) {
      const int i = y * 8 + x;
      tris.push_back(cv::Vec3i(i, i + 1, i + 8));          // clockwise in image coordinates
      tris.push_back(cv::Vec3i(i + 1, i + 9, i + 8));
    }
  std::vector<bool> vv(pts.size(), true), tv(tris.size(), true);
  tv[3] = false;
  const float qnan = std::numeric_limits<float>::quiet_NaN();
  std::vector<float> a(rows * cols + 8, qnan), b(rows * cols + 8, qnan);
  cv::Mat ma(rows, cols, CV_32FC1, a.data()), mb(rows, cols, CV_32FC1, b.data());
  for (size_t t = 0; t < tris.size(); ++t)
    if (tv[t])  // image_utils.cc:380-391
      flame::utils::DrawShadedTriangleBarycentric(pts[tris[t][2]], pts[tris[t][1]], pts[tris[t][0]],
                                                  vals[tris[t][2]], vals[tris[t][1]], vals[tris[t][0]], &ma);
  flame_b200::utils::interpolateMesh(&sess, tris, pts, vals, vv, tv, &mb);
  int bad2 = 0;
  for (int i = 0; i < rows * cols; ++i) bad2 += !same_bits(a[i], b[i]);
  std::printf("interpolateMesh adapter: %dx%d T=%zu mismatches=%d\n", rows, cols, tris.size(), bad2);
  if (bad == 0 && bad2 == 0 && bad_sync == 0) {
    std::printf("ADAPTER_OK\n");
    return 0;
  }
  return 1;
}
