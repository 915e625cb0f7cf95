// C entry points over the reference's NLTGV2-L1 solver compiled VERBATIM
// (src/flame/optimizers/nltgv2_l1_graph_regularizer.cc) against oracle/shim. Test infrastructure
// only (oracle/_ref). The Graph is filled exactly as Flame::syncGraph does: vertices in caller
// order, edges add_edge(source, target) in caller (Triangle edgelist) order (flame.cc:2085-2096).
#include <cstdint>
#include <vector>

#include "flame/optimizers/nltgv2_l1_graph_regularizer.h"

namespace dgraph = flame::optimizers::nltgv2_l1_graph_regularizer;

struct ref_graph_arrays {
  int32_t V, E;
  const float* pos;
  float *x, *w1, *w2, *x_bar, *w1_bar, *w2_bar, *x_prev, *w1_prev, *w2_prev;
  const float *data_term, *data_weight;
  const int32_t* edges;
  const float *alpha, *beta;
  float *q1, *q2, *q3;
};

static void build(const ref_graph_arrays* a, dgraph::Graph* g, std::vector<dgraph::VertexHandle>* vh,
                  std::vector<dgraph::EdgeHandle>* eh) {
  vh->resize(a->V);
  for (int v = 0; v < a->V; ++v) {
    dgraph::VertexData d;
    d.pos = cv::Point2f(a->pos[2 * v], a->pos[2 * v + 1]);
    d.x = a->x[v]; d.w1 = a->w1[v]; d.w2 = a->w2[v];
    d.x_bar = a->x_bar[v]; d.w1_bar = a->w1_bar[v]; d.w2_bar = a->w2_bar[v];
    d.x_prev = a->x_prev[v]; d.w1_prev = a->w1_prev[v]; d.w2_prev = a->w2_prev[v];
    d.data_term = a->data_term[v]; d.data_weight = a->data_weight[v];
    (*vh)[v] = boost::add_vertex(d, *g);
  }
  eh->resize(a->E);
  for (int e = 0; e < a->E; ++e) {
    dgraph::EdgeData d;
    d.alpha = a->alpha[e]; d.beta = a->beta[e];
    d.q1 = a->q1[e]; d.q2 = a->q2[e]; d.q3 = a->q3[e];
    (*eh)[e] = boost::add_edge((*vh)[a->edges[2 * e]], (*vh)[a->edges[2 * e + 1]], d, *g).first;
  }
}

extern "C" void ref_pd_run(const float* params6, ref_graph_arrays* a, int n_iters) {
  dgraph::Params p;
  p.data_factor = params6[0]; p.step_x = params6[1]; p.step_q = params6[2];
  p.theta = params6[3]; p.x_min = params6[4]; p.x_max = params6[5];
  dgraph::Graph g;
  std::vector<dgraph::VertexHandle> vh;
  std::vector<dgraph::EdgeHandle> eh;
  build(a, &g, &vh, &eh);
  for (int it = 0; it < n_iters; ++it) dgraph::step(p, &g);
  for (int v = 0; v < a->V; ++v) {
    const dgraph::VertexData& d = g[vh[v]];
    a->x[v] = d.x; a->w1[v] = d.w1; a->w2[v] = d.w2;
    a->x_bar[v] = d.x_bar; a->w1_bar[v] = d.w1_bar; a->w2_bar[v] = d.w2_bar;
    a->x_prev[v] = d.x_prev; a->w1_prev[v] = d.w1_prev; a->w2_prev[v] = d.w2_prev;
  }
  for (int e = 0; e < a->E; ++e) {
    const dgraph::EdgeData& d = g[eh[e]];
    a->q1[e] = d.q1; a->q2[e] = d.q2; a->q3[e] = d.q3;
  }
}

extern "C" void ref_pd_costs(const float* params6, ref_graph_arrays* a, float* smooth, float* data) {
  dgraph::Params p;
  p.data_factor = params6[0]; p.step_x = params6[1]; p.step_q = params6[2];
  p.theta = params6[3]; p.x_min = params6[4]; p.x_max = params6[5];
  dgraph::Graph g;
  std::vector<dgraph::VertexHandle> vh;
  std::vector<dgraph::EdgeHandle> eh;
  build(a, &g, &vh, &eh);
  *smooth = dgraph::smoothnessCost(p, g);
  *data = dgraph::dataCost(p, g);
}
