// C entry points over the reference's stereo sources compiled VERBATIM against oracle/shim:
// stereo/inverse_depth_filter.cc, stereo/inverse_depth_meas_model.cc, stereo/line_stereo.h,
// stereo/epipolar_geometry.h, utils/image_utils.{h,cc} (bilinearInterp, clipLineLiangBarsky). Test
// infrastructure only.
//
// Flame::updateFeatureIDepths / Flame::trackFeature are private members of flame::Flame whose
// translation unit (flame.cc) needs all of OpenCV/Boost; their control flow (flame.cc:1308-1496,
// 1536-1752, debug drawing removed) is restated below around the verbatim callees.
#include <cmath>
#include <cstdint>
#include <limits>

#include "flame/stereo/inverse_depth_filter.h"
#include "flame/stereo/inverse_depth_meas_model.h"

using flame::stereo::EpipolarGeometry;
namespace idf = flame::stereo::inverse_depth_filter;

struct ref_feature { uint32_t id, frame_id; float x, y, idepth_mu, idepth_var; int32_t valid;
                     uint32_t num_updates, num_dropouts; int32_t search_status; };
struct ref_frame { float pose[7]; const uint8_t* img_pad; const float* gradx_pad; const float* grady_pad; };
struct ref_ls_params { float max_cost; int do_subpixel; float sample_dist; float second_best_factor; };
struct ref_filter_params { int win_size; float search_sigma, min_grad_mag, idepth_min, idepth_max,
                           epilength_min, epilength_max, process_var_factor, process_fail_var_factor;
                           ref_ls_params sparams; };
struct ref_meas_params { int win_size; float pixel_var, epipolar_line_var; };
struct ref_track_params { ref_filter_params fparams; ref_meas_params zparams; float min_baseline,
                          rescale_factor_min, rescale_factor_max, idepth_var_max; int max_dropouts;
                          float outlier_sigma_thresh; int do_meas_fusion, do_letterbox; };
struct ref_diag { float flow_x, flow_y, mu_meas, var_meas; int32_t stage; };

static Sophus::SE3f pose_of(const float* p) {
  return Sophus::SE3f(Eigen::Quaternionf(p[0], p[1], p[2], p[3]), Eigen::Vector3f(p[4], p[5], p[6]));
}

static void fail(const ref_track_params& P, ref_feature& f) {  // flame.cc:1344-1376
  f.idepth_var *= P.fparams.process_fail_var_factor;
  if (f.idepth_var > P.idepth_var_max) f.valid = 0;
  f.num_dropouts++;
  if (f.num_dropouts > (uint32_t)P.max_dropouts) f.valid = 0;
}

extern "C" int ref_update_feature_idepths(const ref_track_params* P, const float* Kf, const float* Kinvf,
                                          int width, int height, int pad, const ref_frame* frames,
                                          int n_frames, int fnew, int curr_pf, ref_feature* feats,
                                          int n_feats, ref_diag* diag) {
  Eigen::Matrix3f K, Kinv;
  for (int i = 0; i < 9; ++i) { K(i / 3, i % 3) = Kf[i]; Kinv(i / 3, i % 3) = Kinvf[i]; }
  idf::Params fparams;
  fparams.win_size = P->fparams.win_size; fparams.search_sigma = P->fparams.search_sigma;
  fparams.min_grad_mag = P->fparams.min_grad_mag; fparams.idepth_min = P->fparams.idepth_min;
  fparams.idepth_max = P->fparams.idepth_max; fparams.epilength_min = P->fparams.epilength_min;
  fparams.epilength_max = P->fparams.epilength_max;
  fparams.process_var_factor = P->fparams.process_var_factor;
  fparams.process_fail_var_factor = P->fparams.process_fail_var_factor;
  fparams.sparams.max_cost = P->fparams.sparams.max_cost;
  fparams.sparams.do_subpixel = P->fparams.sparams.do_subpixel != 0;
  fparams.sparams.sample_dist = P->fparams.sparams.sample_dist;
  fparams.sparams.second_best_factor = P->fparams.sparams.second_best_factor;
  flame::stereo::InverseDepthMeasModel::Params zparams;
  zparams.win_size = P->zparams.win_size; zparams.pixel_var = P->zparams.pixel_var;
  zparams.epipolar_line_var = P->zparams.epipolar_line_var;

  const int prows = height + 2 * pad, pcols = width + 2 * pad;
  const float qnan = std::numeric_limits<float>::quiet_NaN();
  int num_total_updates = 0;
  for (int ii = 0; ii < n_feats; ++ii) {  // flame.cc:1308
    ref_feature& f = feats[ii];
    ref_diag dg = {qnan, qnan, qnan, qnan, 0};
    EpipolarGeometry<float> epigeo(K, Kinv);
    Sophus::SE3f T_ref_to_new = pose_of(frames[fnew].pose).inverse() * pose_of(frames[f.frame_id].pose);
    epigeo.loadGeometry(T_ref_to_new.unit_quaternion(), T_ref_to_new.translation());
    float baseline = T_ref_to_new.translation().norm();
    if (baseline < P->min_baseline) { if (diag) diag[ii] = dg; continue; }

    // ---- trackFeature (flame.cc:1536-1752) ----
    bool track_success = false;
    int stage = 0;
    cv::Point2f flow;
    cv::Point2f u_cmp;
    float idepth_cmp, var_cmp;
    cv::Point2f fxy(f.x, f.y);
    bool pred_success = idf::predict(epigeo, fparams.process_var_factor, fxy, f.idepth_mu, f.idepth_var,
                                     &u_cmp, &idepth_cmp, &var_cmp);
    if (!pred_success) {
      stage = 1;
    } else {
      int row_offset = 0;
      if (P->do_letterbox) row_offset = height / 3;
      int border = P->rescale_factor_max * fparams.win_size / 2 + 1;
      cv::Rect valid_region(border, border + row_offset, width - 2 * border,
                            height - 2 * border - 2 * row_offset);
      float rescale_factor = 1.0f;
      if ((f.idepth_mu > 0.0f) && (idepth_cmp > 0.0f)) rescale_factor = idepth_cmp / f.idepth_mu;
      if ((rescale_factor <= P->rescale_factor_min) || (rescale_factor >= P->rescale_factor_max)) {
        EpipolarGeometry<float> epipf(K, Kinv);
        Sophus::SE3f T_old_to_new = pose_of(frames[curr_pf].pose).inverse() * pose_of(frames[f.frame_id].pose);
        epipf.loadGeometry(T_old_to_new.unit_quaternion(), T_old_to_new.translation());
        cv::Point2f u_pf;
        float idepth_pf, var_pf;
        bool move_success = idf::predict(epipf, fparams.process_var_factor, fxy, f.idepth_mu,
                                         f.idepth_var, &u_pf, &idepth_pf, &var_pf);
        if (!move_success || !valid_region.contains(u_pf)) {
          f.valid = 0;
          stage = 3;
        } else {
          f.frame_id = curr_pf;
          f.x = u_pf.x; f.y = u_pf.y;
          float old_idepth = f.idepth_mu;
          f.idepth_mu = idepth_pf;
          float var_factor4 = idepth_pf / old_idepth;
          var_factor4 *= var_factor4;
          var_factor4 *= var_factor4;
          if (idepth_pf < 1e-6) var_factor4 = 1;
          f.idepth_var *= var_factor4;
          stage = 2;
        }
      } else {
        cv::Point2f u_start, u_end, epi;
        bool region_success = idf::getSearchRegion(fparams, epigeo, width, height, fxy, f.idepth_mu,
                                                   f.idepth_var, &u_start, &u_end, &epi);
        if (!region_success) {
          stage = 4;
        } else if (!valid_region.contains(fxy)) {
          stage = 5;
        } else {
          cv::Point2f offset(pad, pad);
          cv::Mat1b img_ref(prows, pcols, const_cast<uint8_t*>(frames[f.frame_id].img_pad));
          cv::Mat1b img_cmp(prows, pcols, const_cast<uint8_t*>(frames[fnew].img_pad));
          idf::Status st = idf::search(fparams, epigeo, rescale_factor, img_ref, img_cmp, fxy + offset,
                                       u_start + offset, u_end + offset, &u_cmp);
          f.search_status = st;
          if (st != idf::SUCCESS) {
            stage = 6;
          } else {
            flow = u_cmp - offset;
            track_success = true;
          }
        }
      }
    }
    if (!track_success) { fail(*P, f); dg.stage = stage; if (diag) diag[ii] = dg; continue; }
    dg.flow_x = flow.x; dg.flow_y = flow.y;

    // ---- measurement (flame.cc:1380-1390) ----
    flame::stereo::InverseDepthMeasModel model(K, Kinv, zparams);
    model.loadGeometry(pose_of(frames[f.frame_id].pose), pose_of(frames[fnew].pose));
    cv::Mat1b ip_ref(prows, pcols, const_cast<uint8_t*>(frames[f.frame_id].img_pad));
    cv::Mat1b ip_new(prows, pcols, const_cast<uint8_t*>(frames[fnew].img_pad));
    cv::Mat1f gxr(prows, pcols, const_cast<float*>(frames[f.frame_id].gradx_pad));
    cv::Mat1f gyr(prows, pcols, const_cast<float*>(frames[f.frame_id].grady_pad));
    cv::Mat1f gxn(prows, pcols, const_cast<float*>(frames[fnew].gradx_pad));
    cv::Mat1f gyn(prows, pcols, const_cast<float*>(frames[fnew].grady_pad));
    model.loadPaddedImages(ip_ref, ip_new, gxr, gyr, gxn, gyn);
    float mu_meas, var_meas;
    bool sense_success = model.idepth(cv::Point2f(f.x, f.y), flow, &mu_meas, &var_meas);
    dg.mu_meas = mu_meas; dg.var_meas = var_meas;
    if (!sense_success) { fail(*P, f); dg.stage = 7; if (diag) diag[ii] = dg; continue; }

    float mu_post, var_post;
    bool fuse_success = idf::update(f.idepth_mu, f.idepth_var, mu_meas, var_meas, &mu_post, &var_post,
                                    P->outlier_sigma_thresh);
    if (!fuse_success) { fail(*P, f); dg.stage = 8; if (diag) diag[ii] = dg; continue; }
    if (P->do_meas_fusion) { f.idepth_mu = mu_post; f.idepth_var = var_post; }
    else { f.idepth_mu = mu_meas; f.idepth_var = var_meas; }
    f.valid = 1; f.num_updates++; f.num_dropouts = 0;
    num_total_updates++;
    dg.stage = 9;
    if (diag) diag[ii] = dg;
  }
  return num_total_updates;
}

// Thin wrappers so the known-answer tests can also run against the verbatim EpipolarGeometry.
extern "C" void ref_epigeo_query(const float* Kf, const float* Kinvf, const float* q_wxyz, const float* t,
                                 float ux, float uy, float idepth, float* out /*12*/) {
  Eigen::Matrix3f K, Kinv;
  for (int i = 0; i < 9; ++i) { K(i / 3, i % 3) = Kf[i]; Kinv(i / 3, i % 3) = Kinvf[i]; }
  EpipolarGeometry<float> g(K, Kinv);
  g.loadGeometry(Eigen::Quaternionf(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]),
                 Eigen::Vector3f(t[0], t[1], t[2]));
  cv::Point2f u(ux, uy), a, b, c, d;
  g.minDepthProjection(u, &a);
  g.maxDepthProjection(u, &b);
  g.epiline(u, &c, &d);
  cv::Point2f p = g.project(u, idepth);
  cv::Point2f r;
  g.referenceEpiline(u, &r);
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y; out[4] = c.x; out[5] = c.y;
  out[6] = d.x; out[7] = d.y; out[8] = p.x; out[9] = p.y; out[10] = r.x; out[11] = r.y;
}
