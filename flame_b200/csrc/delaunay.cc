// Host-side Delaunay triangulation for the frame pipeline.
//
// The reference triangulates the feature set of every frame with Shewchuk's Triangle
// (src/flame/utils/delaunay.cc:31-77, "zneQB"); at 10 000 vertices that is ~10 ms of host time, two orders of
// magnitude above the GPU's share of a frame. Triangle itself is the caller's (it is not part of this
// library, and callers that need index-identical output keep using it: triangles / edges are plain inputs of
// every entry point). This file gives the pipeline its own triangulator, so that a host thread can
// triangulate frame k+1 while the device works on frame k:
//
//   * divide and conquer (Guibas & Stolfi 1985) on a quad-edge structure with alternating cuts (Dwyer 1987),
//     O(V log V), no super-triangle;
//   * EXACT predicates: positions are rounded to 1/256 px (int32 fixed point, |coordinate| < 8192 px), the
//     orientation test is evaluated in int64 and the in-circle test in __int128, so collinear and cocircular
//     inputs - features sit on the pixel grid - need no tolerances and cannot loop;
//   * output in Triangle's conventions: triangles with positive signed area in image coordinates (what the
//     reference calls clockwise, y pointing down), every undirected edge once.
//
// What this redefines: the INDEX ARRAYS are a Delaunay triangulation of the same points, not Triangle's
// arrays (different but equally valid diagonals in cocircular quadruples, different edge order). Everything
// downstream is bit-exact given the arrays. tests/test_delaunay.py checks the empty-circumcircle property,
// the Euler counts and the winding against the reference's Triangle.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "internal.h"

namespace fb {
namespace {

typedef int32_t Edge;  // 4 * quad + rotation

struct Quad {        // one undirected edge: its four directed / dual edges side by side (one cache line per two)
  Edge next[4];      // onext of rotation 0..3
  int32_t org[4];    // origin vertex (meaningful for rotation 0 and 2)
};

struct Mesh {
  std::vector<Quad> quads;
  std::vector<uint8_t> dead;  // per quad
  const int32_t* px;
  const int32_t* py;

  static Edge rot(Edge e) { return (e & ~3) | ((e + 1) & 3); }
  static Edge sym(Edge e) { return (e & ~3) | ((e + 2) & 3); }
  static Edge invrot(Edge e) { return (e & ~3) | ((e + 3) & 3); }
  Edge& onext(Edge e) { return quads[e >> 2].next[e & 3]; }
  Edge on(Edge e) const { return quads[e >> 2].next[e & 3]; }
  int32_t org(Edge e) const { return quads[e >> 2].org[e & 3]; }
  Edge oprev(Edge e) const { return rot(on(rot(e))); }
  Edge lnext(Edge e) const { return rot(on(invrot(e))); }
  Edge rprev(Edge e) const { return on(sym(e)); }
  int32_t dest(Edge e) const { return org(sym(e)); }

  Edge make_edge(int32_t a, int32_t b) {
    const Edge q = (Edge)(quads.size() << 2);
    quads.push_back(Quad{{q, q + 3, q + 2, q + 1}, {a, -1, b, -1}});
    dead.push_back(0);
    return q;
  }
  void splice(Edge a, Edge b) {
    const Edge alpha = rot(on(a)), beta = rot(on(b));
    std::swap(onext(a), onext(b));
    std::swap(onext(alpha), onext(beta));
  }
  Edge connect(Edge a, Edge b) {
    const Edge e = make_edge(dest(a), org(b));
    splice(e, lnext(a));
    splice(sym(e), b);
    return e;
  }
  void remove(Edge e) {
    splice(e, oprev(e));
    splice(sym(e), oprev(sym(e)));
    dead[e >> 2] = 1;
  }

  // > 0: a, b, c counter-clockwise (positive signed area in the coordinates as given)
  int64_t ccw(int32_t a, int32_t b, int32_t c) const {
    return (int64_t)(px[b] - px[a]) * (py[c] - py[a]) - (int64_t)(py[b] - py[a]) * (px[c] - px[a]);
  }
  // d strictly inside the circle through a, b, c (a, b, c counter-clockwise). |differences| < 2^23, so the
  // squared lengths and the 2x2 minors fit int64 (< 2^47) and each product is ONE 64x64 -> 128-bit multiply.
  bool incircle(int32_t a, int32_t b, int32_t c, int32_t d) const {
    const int64_t dx = px[d], dy = py[d];
    const int64_t ax = px[a] - dx, ay = py[a] - dy;
    const int64_t bx = px[b] - dx, by = py[b] - dy;
    const int64_t cx = px[c] - dx, cy = py[c] - dy;
    const int64_t a2 = ax * ax + ay * ay, b2 = bx * bx + by * by, c2 = cx * cx + cy * cy;
    const int64_t mbc = bx * cy - by * cx, mac = ax * cy - ay * cx, mab = ax * by - ay * bx;
    // double filter: the three products are < 2^94 and exact terms; the double evaluation is off by less than
    // 2^-50 of the sum of their magnitudes - decide by sign unless the result is inside that band
    const double ta = (double)a2 * (double)mbc, tb = (double)b2 * (double)mac, tc = (double)c2 * (double)mab;
    const double det = ta - tb + tc;
    const double band = (std::fabs(ta) + std::fabs(tb) + std::fabs(tc)) * 1e-14;
    if (det > band) return true;
    if (det < -band) return false;
    const __int128 exact = (__int128)a2 * mbc - (__int128)b2 * mac + (__int128)c2 * mab;
    return exact > 0;
  }
  bool right_of(int32_t p, Edge e) const { return ccw(p, dest(e), org(e)) > 0; }
  bool left_of(int32_t p, Edge e) const { return ccw(p, org(e), dest(e)) > 0; }

  // Total order along the cut axis. Axis 0: (x, y) ascending. Axis 1 is axis 0 seen in a frame turned by a quarter
  // (x', y') = (y, -x): y ascending, then x descending. A rotation keeps every orientation / in-circle predicate,
  // so the merge below works unchanged on bottom / top halves.
  bool before(int32_t a, int32_t b, int axis) const {
    if (axis == 0) return px[a] != px[b] ? px[a] < px[b] : py[a] < py[b];
    return py[a] != py[b] ? py[a] < py[b] : px[a] > px[b];
  }
  // e: any ccw hull edge (outer face on its right). Returns the ccw hull edge out of the hull's FIRST vertex in
  // the axis order (the order is a linear functional: one local minimum on a convex hull).
  Edge first_ccw(Edge e, int axis) const {
    for (;;) {
      if (before(dest(e), org(e), axis)) {
        e = rprev(e);  // next hull vertex, counter-clockwise
        continue;
      }
      const Edge p = sym(oprev(e));  // the ccw hull edge out of the previous hull vertex
      if (before(org(p), org(e), axis)) {
        e = p;
        continue;
      }
      return e;
    }
  }
  // e: any cw hull edge (outer face on its left). Returns the cw hull edge out of the hull's LAST vertex.
  Edge last_cw(Edge e, int axis) const {
    for (;;) {
      if (before(org(e), dest(e), axis)) {
        e = lnext(e);  // next hull vertex, clockwise
        continue;
      }
      const Edge p = sym(on(e));  // the cw hull edge out of the previous hull vertex
      if (before(org(e), org(p), axis)) {
        e = p;
        continue;
      }
      return e;
    }
  }

  // s[lo, hi): distinct vertex ids (reordered in place). Cuts alternate between the axes (Dwyer 1987: the merged
  // hulls stay compact, far fewer edges are created only to be deleted again than with vertical strips). Returns
  // the ccw convex-hull edge out of the first vertex in the order of `axis` and the cw hull edge out of the last.
  void build(int32_t* s, int lo, int hi, int axis, Edge* le, Edge* re) {
    const int n = hi - lo;
    const auto cmp = [this, axis](int32_t a, int32_t b) { return before(a, b, axis); };
    if (n <= 3) std::sort(s + lo, s + hi, cmp);
    if (n == 2) {
      const Edge a = make_edge(s[lo], s[lo + 1]);
      *le = a;
      *re = sym(a);
      return;
    }
    if (n == 3) {
      const Edge a = make_edge(s[lo], s[lo + 1]), b = make_edge(s[lo + 1], s[lo + 2]);
      splice(sym(a), b);
      const int64_t o = ccw(s[lo], s[lo + 1], s[lo + 2]);
      if (o > 0) {
        connect(b, a);
        *le = a;
        *re = sym(b);
      } else if (o < 0) {
        const Edge c = connect(b, a);
        *le = sym(c);
        *re = c;
      } else {  // collinear
        *le = a;
        *re = sym(b);
      }
      return;
    }
    const int mid = lo + n / 2;
    std::nth_element(s + lo, s + mid, s + hi, cmp);
    Edge ldo, ldi, rdi, rdo;
    build(s, lo, mid, 1 - axis, &ldo, &ldi);
    build(s, mid, hi, 1 - axis, &rdi, &rdo);
    // the halves were built across the other axis: move their handles to this axis' extreme vertices
    ldo = first_ccw(ldo, axis);
    ldi = last_cw(ldi, axis);
    rdi = first_ccw(rdi, axis);
    rdo = last_cw(rdo, axis);
    for (;;) {  // lower common tangent of the two hulls
      if (left_of(org(rdi), ldi)) ldi = lnext(ldi);
      else if (right_of(org(ldi), rdi)) rdi = rprev(rdi);
      else break;
    }
    Edge basel = connect(sym(rdi), ldi);
    if (org(ldi) == org(ldo)) ldo = sym(basel);
    if (org(rdi) == org(rdo)) rdo = basel;
    for (;;) {  // merge: the rising bubble
      Edge lcand = on(sym(basel));
      if (right_of(dest(lcand), basel))
        while (incircle(dest(basel), org(basel), dest(lcand), dest(on(lcand)))) {
          const Edge t = on(lcand);
          remove(lcand);
          lcand = t;
        }
      Edge rcand = oprev(basel);
      if (right_of(dest(rcand), basel))
        while (incircle(dest(basel), org(basel), dest(rcand), dest(oprev(rcand)))) {
          const Edge t = oprev(rcand);
          remove(rcand);
          rcand = t;
        }
      const bool lv = right_of(dest(lcand), basel), rv = right_of(dest(rcand), basel);
      if (!lv && !rv) break;
      if (!lv || (rv && incircle(dest(lcand), org(lcand), org(rcand), dest(rcand))))
        basel = connect(rcand, sym(basel));
      else
        basel = connect(sym(basel), sym(lcand));
    }
    *le = ldo;
    *re = rdo;
  }
};

}  // namespace
}  // namespace fb

extern "C" int flame_b200_delaunay(int32_t V, const float* pos, int32_t tri_capacity, int32_t* n_tris,
                                   int32_t* tris, int32_t edge_capacity, int32_t* n_edges, int32_t* edges) {
  using namespace fb;
  if (V < 0 || (V > 0 && !pos) || !n_tris || !n_edges || tri_capacity < 0 || edge_capacity < 0) {
    set_error("delaunay: bad arguments");
    return FLAME_B200_EINVAL;
  }
  *n_tris = 0;
  *n_edges = 0;
  // per-thread scratch, kept across calls: a frame pipeline triangulates every frame on several host threads, and
  // allocating / page-faulting ~3 MB per call serialises them on the process' memory map
  static thread_local std::vector<int32_t> px, py, order;
  static thread_local std::vector<uint64_t> keys;
  static thread_local Mesh m;
  px.resize(V);
  py.resize(V);
  order.resize(V);
  for (int32_t v = 0; v < V; ++v) {
    const float x = pos[2 * v], y = pos[2 * v + 1];
    if (!(std::fabs(x) < 8192.0f) || !(std::fabs(y) < 8192.0f)) {
      set_error("delaunay: coordinate outside (-8192, 8192) px or not finite");
      return FLAME_B200_EINVAL;
    }
    px[v] = (int32_t)std::lrintf(x * 256.0f);
    py[v] = (int32_t)std::lrintf(y * 256.0f);
    order[v] = v;
  }
  {  // sort by (x, y, id): one packed 64-bit key per vertex (x, y biased to unsigned 22 bits, id 20 bits) when
     // the ids fit, else the comparison sort
    if (V < (1 << 20)) {
      keys.resize(V);
      for (int32_t v = 0; v < V; ++v)
        keys[v] = ((uint64_t)(uint32_t)(px[v] + (1 << 21)) << 42) | ((uint64_t)(uint32_t)(py[v] + (1 << 21)) << 20) |
                  (uint64_t)v;
      std::sort(keys.begin(), keys.end());
      for (int32_t v = 0; v < V; ++v) order[v] = (int32_t)(keys[v] & ((1u << 20) - 1));
    } else {
      std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        return px[a] != px[b] ? px[a] < px[b] : (py[a] != py[b] ? py[a] < py[b] : a < b);
      });
    }
  }
  // coincident points (after rounding): the first one represents them, the others stay unconnected (as
  // Triangle leaves duplicate vertices out)
  int n = 0;
  for (int32_t i = 0; i < V; ++i)
    if (n == 0 || px[order[i]] != px[order[n - 1]] || py[order[i]] != py[order[n - 1]]) order[n++] = order[i];
  if (n < 2) return 0;
  m.quads.clear();
  m.dead.clear();
  m.px = px.data();
  m.py = py.data();
  m.quads.reserve((size_t)8 * n + 16);  // edges ever created (kept + deleted); ~6.8 n for uniform points
  m.dead.reserve((size_t)8 * n + 16);
  Edge le, re;
  m.build(order.data(), 0, n, 0, &le, &re);
  const int32_t quads = (int32_t)m.dead.size();
  int32_t ne = 0, nt = 0;
  for (int32_t q = 0; q < quads; ++q) {
    if (m.dead[q]) continue;
    const Edge e = 4 * q;
    if (ne < edge_capacity && edges) {
      edges[2 * ne] = m.org(e);
      edges[2 * ne + 1] = m.dest(e);
    }
    ++ne;
  }
  // faces: the cycle to the left of a directed edge; interior faces are positively oriented triangles. Each is
  // reported from its lowest-numbered directed edge.
  for (int32_t q = 0; q < quads; ++q) {
    if (m.dead[q]) continue;
    for (int half = 0; half < 2; ++half) {
      const Edge e0 = 4 * q + 2 * half;
      const Edge e1 = m.lnext(e0), e2 = m.lnext(e1);
      if (m.lnext(e2) != e0 || e1 < e0 || e2 < e0) continue;
      const int32_t a = m.org(e0), b = m.org(e1), c = m.org(e2);
      if (m.ccw(a, b, c) <= 0) continue;  // the outer face (or a degenerate cycle)
      if (nt < tri_capacity && tris) {
        tris[3 * nt] = a;
        tris[3 * nt + 1] = b;
        tris[3 * nt + 2] = c;
      }
      ++nt;
    }
  }
  *n_edges = ne;
  *n_tris = nt;
  if ((tris && nt > tri_capacity) || (edges && ne > edge_capacity)) {
    set_error("delaunay: output capacity too small (needs " + std::to_string(nt) + " triangles, " +
              std::to_string(ne) + " edges)");
    return FLAME_B200_ENOMEM;
  }
  return 0;
}
