// C entry point over the reference's Delaunay wrapper (test infrastructure; compiled only when
// /root/reference is present, output in oracle/_ref/). Calls Shewchuk Triangle exactly as
// src/flame/utils/delaunay.cc:31-77 does (switches "zneQB", float coordinates).
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include "flame/external/triangle/triangle.h"

extern "C" int ref_delaunay(int32_t n, const float* xy, int32_t* n_tris, int32_t** tris,
                            int32_t* n_edges, int32_t** edges, int32_t** neighbors) {
  struct triangulateio in, out;
  std::memset(&in, 0, sizeof(in));
  std::memset(&out, 0, sizeof(out));
  in.numberofpoints = n;
  in.pointlist = static_cast<float*>(std::malloc(sizeof(float) * 2 * n));
  std::memcpy(in.pointlist, xy, sizeof(float) * 2 * n);
  char parameters[] = "zneQB";  // delaunay.cc:67
  ::triangulate(parameters, &in, &out, NULL);
  std::free(in.pointlist);
  *n_tris = out.numberoftriangles;
  *n_edges = out.numberofedges;
  *tris = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * 3 * out.numberoftriangles));
  *neighbors = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * 3 * out.numberoftriangles));
  *edges = static_cast<int32_t*>(std::malloc(sizeof(int32_t) * 2 * out.numberofedges));
  for (int i = 0; i < 3 * out.numberoftriangles; ++i) {
    (*tris)[i] = out.trianglelist[i];
    (*neighbors)[i] = out.neighborlist[i];
  }
  for (int i = 0; i < 2 * out.numberofedges; ++i) (*edges)[i] = out.edgelist[i];
  std::free(out.pointlist);
  std::free(out.trianglelist);
  std::free(out.edgelist);
  std::free(out.neighborlist);
  return 0;
}

extern "C" void ref_free(void* p) { std::free(p); }
