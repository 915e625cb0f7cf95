// C entry points over the reference's dense projection and image helpers (test infrastructure;
// oracle/_ref only). utils/rasterization.cc and utils/image_utils.cc are compiled VERBATIM from the
// reference checkout against oracle/shim; this file only adapts flat arrays to their signatures.
#include <cstdint>
#include <vector>

#include "flame/utils/image_utils.h"
#include "flame/utils/rasterization.h"

extern "C" void ref_draw_shaded_triangle_barycentric(int p1x, int p1y, int p2x, int p2y, int p3x,
                                                     int p3y, float v1, float v2, float v3,
                                                     float* img, int rows, int cols) {
  cv::Mat m(rows, cols, CV_32FC1, img);
  flame::utils::DrawShadedTriangleBarycentric(cv::Point(p1x, p1y), cv::Point(p2x, p2y),
                                              cv::Point(p3x, p3y), v1, v2, v3, &m);
}

// utils::interpolateMesh (image_utils.cc:373-396), verbatim. img must have >= 4 floats of slack
// after rows*cols (the 4-wide SSE sweep may touch them).
extern "C" void ref_interpolate_mesh(int32_t T, const int32_t* tris, const float* pos,
                                     const float* values, const uint8_t* vtx_valid,
                                     const uint8_t* tri_valid, float* img, int rows, int cols) {
  int V = 0;
  for (int i = 0; i < 3 * T; ++i) V = tris[i] + 1 > V ? tris[i] + 1 : V;
  std::vector<flame::Triangle> triangles(T);
  std::vector<bool> tv(T, true), vv(V, true);
  for (int t = 0; t < T; ++t) {
    triangles[t] = flame::Triangle(tris[3 * t], tris[3 * t + 1], tris[3 * t + 2]);
    if (tri_valid) tv[t] = tri_valid[t] != 0;
  }
  std::vector<cv::Point2f> vertices(V);
  std::vector<float> vals(V);
  for (int v = 0; v < V; ++v) {
    vertices[v] = cv::Point2f(pos[2 * v], pos[2 * v + 1]);
    vals[v] = values[v];
    if (vtx_valid) vv[v] = vtx_valid[v] != 0;
  }
  cv::Mat m(rows, cols, CV_32FC1, img);
  flame::utils::interpolateMesh(triangles, vertices, vals, vv, tv, &m);
}

// utils::getCentralGradient<uint8_t, float> (image_utils.cc:36-108, the SSE specialisation), verbatim.
extern "C" void ref_central_gradient_u8(int width, int height, const uint8_t* img, float* gradx,
                                        float* grady) {
  flame::utils::getCentralGradient<uint8_t, float>(width, height, img, gradx, grady);
}

// utils::clipLineLiangBarsky (image_utils.cc:269-371), verbatim.
extern "C" int ref_clip_line(float xmin, float xmax, float ymin, float ymax, float x0, float y0,
                             float x1, float y1, float* out4) {
  return flame::utils::clipLineLiangBarsky(xmin, xmax, ymin, ymax, x0, y0, x1, y1, out4, out4 + 1,
                                           out4 + 2, out4 + 3) ? 1 : 0;
}
