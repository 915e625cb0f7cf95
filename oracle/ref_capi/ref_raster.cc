// C entry points over the reference's dense projection (test infrastructure; oracle/_ref only).
// utils::DrawShadedTriangleBarycentric is the reference's rasterization.cc compiled verbatim;
// interpolateMesh below restates the 20-line loop of utils/image_utils.cc:373-396 because
// image_utils.cc as a whole needs the full OpenCV API.
#include <cstdint>

#include "flame/utils/rasterization.h"

extern "C" void ref_draw_shaded_triangle_barycentric(int p1x, int p1y, int p2x, int p2y, int p3x,
                                                     int p3y, float v1, float v2, float v3,
                                                     float* img, int rows, int cols) {
  cv::Mat m(rows, cols, CV_32FC1, img);
  flame::utils::DrawShadedTriangleBarycentric(cv::Point(p1x, p1y), cv::Point(p2x, p2y),
                                              cv::Point(p3x, p3y), v1, v2, v3, &m);
}

// img must have >= 4 floats of slack after rows*cols (the 4-wide SSE sweep may touch them).
extern "C" void ref_interpolate_mesh(int32_t T, const int32_t* tris, const float* pos,
                                     const float* values, const uint8_t* vtx_valid,
                                     const uint8_t* tri_valid, float* img, int rows, int cols) {
  cv::Mat m(rows, cols, CV_32FC1, img);
  for (int ii = 0; ii < T; ++ii) {
    int a = tris[3 * ii], b = tris[3 * ii + 1], c = tris[3 * ii + 2];
    if ((!tri_valid || tri_valid[ii]) &&
        (!vtx_valid || (vtx_valid[a] && vtx_valid[b] && vtx_valid[c]))) {
      // image_utils.cc:385-391: cv::Point2f -> cv::Point (saturate_cast, half-to-even), reversed.
      flame::utils::DrawShadedTriangleBarycentric(
          cv::Point2f(pos[2 * c], pos[2 * c + 1]), cv::Point2f(pos[2 * b], pos[2 * b + 1]),
          cv::Point2f(pos[2 * a], pos[2 * a + 1]), values[c], values[b], values[a], &m);
    }
  }
}
