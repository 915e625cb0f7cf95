// C entry point over the reference's Flame::detectFeatures (test infrastructure; oracle/_ref only). The member
// function is sliced VERBATIM out of flame.cc by oracle/ref_capi/slice_flame.py (its live part is the
// single-pass grid detector, flame.cc:1192-1277; the rest of the body is commented out upstream) and compiled
// against oracle/shim; this file declares just enough of class Flame around it and adapts flat arrays.
#include <cstdint>
#include <cstdio>
#include <limits>
#include <vector>

#include "flame/params.h"
#include "flame/stereo/epipolar_geometry.h"
#include "flame/types.h"
#include "flame/utils/assert.h"
#include "flame/utils/frame.h"
#include "flame/utils/image_utils.h"
#include "flame/utils/stats_tracker.h"

namespace flame {

class Flame {
 public:
  static void detectFeatures(const Params& params, const Matrix3f& K, const Matrix3f& Kinv, const utils::Frame& fref,
                             const utils::Frame& fprev, const utils::Frame& fcmp, const Image1f& idepthmap,
                             const std::vector<cv::Point2f>& curr_feats, Image1f* error,
                             std::vector<cv::Point2f>* features, utils::StatsTracker* stats, Image3b* debug_img);
  // debug drawing (flame.cc:2363+): never reached with debug_draw_detections == false
  static void drawDetections(const Params&, const Image1b&, const Image1f&, float, const std::vector<cv::KeyPoint>&,
                             utils::StatsTracker*, Image3b*) {}
};

#include "flame_slices_detect.inc"

}  // namespace flame

struct ref_detect_params {
  float rescale_factor_max;
  int32_t win_size, do_letterbox;
  float min_grad_mag;
  int32_t detection_win_size;
};

static Sophus::SE3f to_pose(const float p[7]) {
  return Sophus::SE3f(Eigen::Quaternionf(p[0], p[1], p[2], p[3]), flame::Vector3f(p[4], p[5], p[6]));
}

extern "C" int32_t ref_detect_features(const float K[9], const float Kinv[9], const float pose_ref[7],
                                       const float pose_prev[7], int width, int height, const ref_detect_params* p,
                                       const float* gradx, const float* grady, const float* curr_xy, int32_t n_curr,
                                       float* out_xy, int32_t max_out) {
  using namespace flame;
  Params params;
  params.debug_quiet = true;
  params.debug_draw_detections = false;
  params.rescale_factor_max = p->rescale_factor_max;
  params.fparams.win_size = p->win_size;
  params.do_letterbox = p->do_letterbox != 0;
  params.min_grad_mag = p->min_grad_mag;
  params.detection_win_size = p->detection_win_size;
  Matrix3f Km, Kinvm;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      Km(i, j) = K[3 * i + j];
      Kinvm(i, j) = Kinv[3 * i + j];
    }
  utils::Frame fref(1), fprev(1);
  fref.img[0] = Image1b(height, width);  // only its size is read
  fref.gradx[0] = Image1f(height, width, const_cast<float*>(gradx));
  fref.grady[0] = Image1f(height, width, const_cast<float*>(grady));
  fref.pose = to_pose(pose_ref);
  fprev.pose = to_pose(pose_prev);
  std::vector<cv::Point2f> curr(n_curr), feats;
  for (int i = 0; i < n_curr; ++i) curr[i] = cv::Point2f(curr_xy[2 * i], curr_xy[2 * i + 1]);
  Image1f idepthmap, error;
  utils::StatsTracker stats;
  Flame::detectFeatures(params, Km, Kinvm, fref, fprev, fprev, idepthmap, curr, &error, &feats, &stats, nullptr);
  for (size_t i = 0; i < feats.size() && (int32_t)i < max_out; ++i) {
    out_xy[2 * i] = feats[i].x;
    out_xy[2 * i + 1] = feats[i].y;
  }
  return (int32_t)feats.size();
}
