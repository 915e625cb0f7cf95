// Minimal stand-in for boost::adjacency_list<hash_setS, hash_setS, undirectedS, VP, EP> with the
// free functions the reference's regulariser uses (vertices, edges, source, target, add_vertex,
// add_edge, tie, operator[]) plus the edit calls Flame::syncGraph makes (edge, remove_edge,
// clear_vertex, remove_vertex), which tests/cpp uses to replay a graph sync. Edge iteration walks
// the global edge list in insertion order, as Boost's undirected adjacency_list does (its m_edges
// is a std::list appended by add_edge). Vertex iteration is insertion order (Boost's hash_setS order is address dependent and does
// not affect per-vertex arithmetic). Test infrastructure only (oracle/_ref).
#pragma once
#include <cstddef>
#include <list>
#include <utility>
#include <vector>

namespace boost {

struct hash_setS {};
struct undirectedS {};

template <typename A, typename B>
struct tie_t {
  A& a; B& b;
  template <typename P> tie_t& operator=(const P& p) { a = p.first; b = p.second; return *this; }
};
template <typename A, typename B> tie_t<A, B> tie(A& a, B& b) { tie_t<A, B> t = {a, b}; return t; }

template <typename OutS, typename VtxS, typename Dir, typename VP, typename EP>
class adjacency_list {
 public:
  struct VNode { VP prop; };
  struct ENode { VNode* s; VNode* t; EP prop; };
  typedef VNode* vertex_descriptor;
  typedef ENode* edge_descriptor;
  struct vertex_iterator {
    typename std::list<VNode>::iterator it;
    vertex_descriptor operator*() const { return &*it; }
    vertex_iterator& operator++() { ++it; return *this; }
    bool operator!=(const vertex_iterator& o) const { return it != o.it; }
    bool operator==(const vertex_iterator& o) const { return it == o.it; }
  };
  struct edge_iterator {
    typename std::list<ENode>::iterator it;
    edge_descriptor operator*() const { return &*it; }
    edge_iterator& operator++() { ++it; return *this; }
    bool operator!=(const edge_iterator& o) const { return it != o.it; }
    bool operator==(const edge_iterator& o) const { return it == o.it; }
  };
  VP& operator[](vertex_descriptor v) { return v->prop; }
  const VP& operator[](vertex_descriptor v) const { return v->prop; }
  EP& operator[](edge_descriptor e) { return e->prop; }
  const EP& operator[](edge_descriptor e) const { return e->prop; }
  mutable std::list<VNode> vlist;
  mutable std::list<ENode> elist;
};

template <typename G> struct graph_traits {
  typedef typename G::vertex_descriptor vertex_descriptor;
  typedef typename G::edge_descriptor edge_descriptor;
};

#define FB_AL adjacency_list<O, V, D, VP, EP>
template <typename O, typename V, typename D, typename VP, typename EP>
std::pair<typename FB_AL::vertex_iterator, typename FB_AL::vertex_iterator> vertices(const FB_AL& g) {
  typename FB_AL::vertex_iterator a = {g.vlist.begin()}, b = {g.vlist.end()};
  return std::make_pair(a, b);
}
template <typename O, typename V, typename D, typename VP, typename EP>
std::pair<typename FB_AL::edge_iterator, typename FB_AL::edge_iterator> edges(const FB_AL& g) {
  typename FB_AL::edge_iterator a = {g.elist.begin()}, b = {g.elist.end()};
  return std::make_pair(a, b);
}
template <typename O, typename V, typename D, typename VP, typename EP>
typename FB_AL::vertex_descriptor source(typename FB_AL::edge_descriptor e, const FB_AL&) { return e->s; }
template <typename O, typename V, typename D, typename VP, typename EP>
typename FB_AL::vertex_descriptor target(typename FB_AL::edge_descriptor e, const FB_AL&) { return e->t; }
template <typename O, typename V, typename D, typename VP, typename EP>
typename FB_AL::vertex_descriptor add_vertex(const VP& p, FB_AL& g) {
  typename FB_AL::VNode n; n.prop = p; g.vlist.push_back(n); return &g.vlist.back();
}
template <typename O, typename V, typename D, typename VP, typename EP>
std::pair<typename FB_AL::edge_descriptor, bool> add_edge(typename FB_AL::vertex_descriptor s,
                                                          typename FB_AL::vertex_descriptor t,
                                                          const EP& p, FB_AL& g) {
  typename FB_AL::ENode n; n.s = s; n.t = t; n.prop = p; g.elist.push_back(n);
  return std::make_pair(&g.elist.back(), true);
}
template <typename O, typename V, typename D, typename VP, typename EP>
std::pair<typename FB_AL::edge_descriptor, bool> edge(typename FB_AL::vertex_descriptor s,
                                                      typename FB_AL::vertex_descriptor t, const FB_AL& g) {
  for (auto it = g.elist.begin(); it != g.elist.end(); ++it)
    if ((it->s == s && it->t == t) || (it->s == t && it->t == s)) return std::make_pair(&*it, true);
  return std::make_pair(static_cast<typename FB_AL::edge_descriptor>(nullptr), false);
}
template <typename O, typename V, typename D, typename VP, typename EP>
void remove_edge(typename FB_AL::edge_descriptor e, FB_AL& g) {
  for (auto it = g.elist.begin(); it != g.elist.end(); ++it)
    if (&*it == e) { g.elist.erase(it); return; }
}
template <typename O, typename V, typename D, typename VP, typename EP>
void clear_vertex(typename FB_AL::vertex_descriptor v, FB_AL& g) {
  for (auto it = g.elist.begin(); it != g.elist.end();)
    if (it->s == v || it->t == v) it = g.elist.erase(it); else ++it;
}
template <typename O, typename V, typename D, typename VP, typename EP>
void remove_vertex(typename FB_AL::vertex_descriptor v, FB_AL& g) {
  for (auto it = g.vlist.begin(); it != g.vlist.end(); ++it)
    if (&*it == v) { g.vlist.erase(it); return; }
}
template <typename O, typename V, typename D, typename VP, typename EP>
std::size_t num_vertices(const FB_AL& g) { return g.vlist.size(); }
template <typename O, typename V, typename D, typename VP, typename EP>
std::size_t num_edges(const FB_AL& g) { return g.elist.size(); }
#undef FB_AL

}  // namespace boost
