// ref_harness.cpp -- C entry points over the REFERENCE'S OWN hot-path sources.  TEST INFRASTRUCTURE ONLY.
//
// oracle/_ref/libpagk_ref.so = this file + /root/reference/src/{gyro_aided_tracker,patch_match,utils,frame}.cpp and src/ORBextractor.cc compiled
// unmodified, where they lie, against the stand-in headers in oracle/ref_shim/ (OpenCV, Eigen3 and glog are not
// installed here).  See oracle/ref_shim/pagk_cv_shim.hpp for what is real and what is restated.  Only tests/ and
// bench.py's CPU legs load it; it is the checker for oracle/pagk_oracle.cpp (the restatement that travels to the GPU
// box in source form) and for the CUDA path.
//
// Entry points mirror oracle/pagk_oracle.cpp's so one Python driver serves both:
//   pagk_ref_track        -> GyroAidedTracker(frameRef, frameCur, calib, bias, table, type, method, "", half)
//                            .TrackFeatures()                         (src/gyro_aided_tracker.cpp:30-49, 344-426)
//   pagk_ref_patch_match  -> PatchMatch(&tracker, ...).OpticalFlowMultiLevel()   (src/patch_match.cpp:33-142)
// TrackFeatures() hard-codes iterations = 10 and pyramids = 3 (src/gyro_aided_tracker.cpp:276-277).  For any other
// pagk_params the harness runs the same three steps through the public interface: TrackFeatures() as GYRO_PREDICT
// (gyro integration + prediction), PatchMatch with the requested levels, and the result filter of :289-336 written
// out below ("composed" path, pagk_ref_last_path() == 1).
#include "../include/pagk.h"

#include "gyro_aided_tracker.h"
#include "patch_match.h"
#include "pagk_cv_resize.h"
#include "pagk_cv_fast.h"

#include <atomic>
#include <chrono>

// ------------------------------------------------------------------------------------------------------------------
// definitions for the stand-in headers
// ------------------------------------------------------------------------------------------------------------------
namespace pagk_eigen_shim {
std::atomic<long long> g_llt_calls{0};
}

namespace {
void finish_guard(cv::Mat &m) {  // out-of-bounds convention of oracle/pagk_oracle.cpp: guard row + one byte
  std::memcpy(m.data + (size_t)m.rows * m.step, m.data + (size_t)(m.rows - 1) * m.step, m.step);
  m.data[(size_t)(m.rows + 1) * m.step] = m.data[(size_t)m.rows * m.step];
}
cv::Mat g_injected_H, g_injected_F;
thread_local int g_last_path = 0;
}  // namespace

namespace cv {
void resize(const Mat &src, Mat &dst, Size dsize, double, double, int interpolation) {
  if (src.type() != CV_8UC1 || interpolation != INTER_LINEAR) shim_abort("resize of this type / interpolation");
  Mat out(dsize.height, dsize.width, CV_8UC1);
  pagk_cv::resize_half(src.data, src.cols, src.rows, (int)src.step, out.data, out.cols, out.rows);
  finish_guard(out);
  dst = out;
}
void FAST(const Mat &image, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression) {
  if (image.type() != CV_8UC1) shim_abort("FAST on this type");
  const int cap = image.rows * image.cols;
  std::vector<float> xy((size_t)2 * std::max(cap, 1)), rs((size_t)std::max(cap, 1));
  const int n = pagk_cv::fast_detect(image.data, image.cols, image.rows, (int)image.step, threshold, nonmaxSuppression, nullptr, cap,
                                     xy.data(), rs.data());
  keypoints.clear();
  for (int k = 0; k < n; ++k) keypoints.push_back(KeyPoint(xy[2 * k], xy[2 * k + 1], 7.f, -1, rs[k]));
}
void copyMakeBorder(const Mat &src, Mat &dst, int top, int bottom, int left, int right, int borderType) {
  if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101 || src.type() != CV_8UC1) shim_abort("copyMakeBorder of this kind");
  const int R = src.rows + top + bottom, C = src.cols + left + right;
  if (dst.rows != R || dst.cols != C || dst.type() != src.type() || !dst.data) dst.create(R, C, src.type());
  auto refl = [](int p, int n) { if (n == 1) return 0; while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p; return p; };
  // through a temporary: src may be a view of dst (ORBextractor::ComputePyramid, src/ORBextractor.cc:1222)
  std::vector<uchar> tmp((size_t)R * C);
  for (int r = 0; r < R; ++r) {
    const uchar *srow = src.data + (size_t)refl(r - top, src.rows) * src.step;
    for (int c = 0; c < C; ++c) tmp[(size_t)r * C + c] = srow[refl(c - left, src.cols)];
  }
  for (int r = 0; r < R; ++r) std::memcpy(dst.data + (size_t)r * dst.step, &tmp[(size_t)r * C], (size_t)C);
}
Mat findHomography(const std::vector<Point2f> &, const std::vector<Point2f> &, int, double, Mat &) {
  if (g_injected_H.empty()) shim_abort("findHomography (no model injected)");
  return g_injected_H.clone();
}
Mat findFundamentalMat(const std::vector<Point2f> &, const std::vector<Point2f> &, int, double, double, Mat &) {
  if (g_injected_F.empty()) shim_abort("findFundamentalMat (no model injected)");
  return g_injected_F.clone();
}
void calcOpticalFlowPyrLK(const Mat &, const Mat &, const std::vector<Point2f> &, std::vector<Point2f> &, std::vector<uchar> &,
                          std::vector<float> &, Size, int, TermCriteria, int, double) {
  shim_abort("calcOpticalFlowPyrLK");
}
void undistortPoints(const Mat &, Mat &, const Mat &, const Mat &, const Mat &, const Mat &) { shim_abort("undistortPoints"); }
void initUndistortRectifyMap(const Mat &, const Mat &, const Mat &, const Mat &, Size, int, Mat &, Mat &) {
  shim_abort("initUndistortRectifyMap");
}
Mat getOptimalNewCameraMatrix(const Mat &, const Mat &, Size, double, Size, void *) { shim_abort("getOptimalNewCameraMatrix"); }
}  // namespace cv

// src/frame.cpp (Frame::SetPredictKeyPointsAndMask) and src/ORBextractor.cc (ORBextractor::DetectFeatures, the keypoint
// top-up) are compiled too.

// ------------------------------------------------------------------------------------------------------------------
namespace {

struct Scene {  // the caller-owned objects the tracker binds by const reference (include/gyro_aided_tracker.h:180-188)
  CameraParams cam;
  Frame ref, cur;
  IMU::Calib calib;
  cv::Point3f bias;
  cv::Mat table;
};

cv::Mat image_mat(const uint8_t *img, int width, int height, int pitch) {
  cv::Mat m(height, width, CV_8UC1);  // continuous (step == cols), as the GPU path and the restatement store level 0
  for (int y = 0; y < height; ++y) std::memcpy(m.data + (size_t)y * m.step, img + (size_t)y * pitch, (size_t)width);
  finish_guard(m);
  return m;
}

void fill_scene(Scene &s, const uint8_t *img_ref, const uint8_t *img_cur, int width, int height, int pitch, int n_keys,
                const float *keys_ref_un, const float *keys_ref, const float K[9], const float dist[5], int n_dist) {
  s.cam.mK = cv::Mat(3, 3, CV_32F);
  for (int i = 0; i < 9; ++i) s.cam.mK.at<float>(i / 3, i % 3) = K[i];
  s.cam.mDistCoef = cv::Mat(n_dist == 5 ? 5 : 4, 1, CV_32F);
  for (int i = 0; i < (n_dist == 5 ? 5 : 4); ++i) s.cam.mDistCoef.at<float>(i) = dist[i];
  s.cam.width = width; s.cam.height = height;
  s.ref.mpCameraParams = &s.cam; s.cur.mpCameraParams = &s.cam;
  s.ref.mGray = image_mat(img_ref, width, height, pitch);
  s.cur.mGray = image_mat(img_cur, width, height, pitch);
  s.ref.mvKeysUn.resize(n_keys); s.ref.mvKeys.resize(n_keys);
  for (int i = 0; i < n_keys; ++i) {
    s.ref.mvKeysUn[i].pt = cv::Point2f(keys_ref_un[2 * i], keys_ref_un[2 * i + 1]);
    const float *k = keys_ref ? keys_ref : keys_ref_un;
    s.ref.mvKeys[i].pt = cv::Point2f(k[2 * i], k[2 * i + 1]);
  }
}

void fill_scene(Scene &s, const pagk_pair_in &in) {
  fill_scene(s, in.img_ref, in.img_cur, in.width, in.height, in.pitch, in.n_keys, in.keys_ref_un, in.keys_ref, in.K, in.dist,
             in.n_dist);
  s.ref.mTimeStamp = in.t_ref; s.cur.mTimeStamp = in.t_cur;
  s.cur.mvImuFromLastFrame.resize(in.n_imu);
  for (int i = 0; i < in.n_imu; ++i)
    s.cur.mvImuFromLastFrame[i] = IMU::Point(cv::Point3f(0, 0, 0), cv::Point3f(in.imu_w[3 * i], in.imu_w[3 * i + 1], in.imu_w[3 * i + 2]),
                                             in.imu_t[i]);
  s.calib.Tbc = cv::Mat::eye(4, 4, CV_32F);
  for (int i = 0; i < 9; ++i) s.calib.Tbc.at<float>(i / 3, i % 3) = in.Rbc[i];
  s.bias = cv::Point3f(in.bias_g[0], in.bias_g[1], in.bias_g[2]);
  if (in.normalize_table) s.table = cv::Mat(in.height, in.width, CV_32FC2, (void *)in.normalize_table);
}

void put_pts(float *dst, const std::vector<cv::Point2f> &v, int n) {
  if (!dst) return;
  for (int i = 0; i < n; ++i) {
    dst[2 * i] = i < (int)v.size() ? v[i].x : 0.f;
    dst[2 * i + 1] = i < (int)v.size() ? v[i].y : 0.f;
  }
}
void put_quads(float *dst, const std::vector<std::vector<cv::Point2f>> &v, int n) {
  if (!dst) return;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < 4; ++j) {
      const bool has = i < (int)v.size() && j < (int)v[i].size();
      dst[8 * i + 2 * j] = has ? v[i][j].x : 0.f;
      dst[8 * i + 2 * j + 1] = has ? v[i][j].y : 0.f;
    }
}
template <typename T, typename V> void put_vec(T *dst, const V &v, int n) {
  if (!dst) return;
  for (int i = 0; i < n; ++i) dst[i] = i < (int)v.size() ? (T)v[i] : (T)0;
}

void export_tracker(const GyroAidedTracker &t, int n, pagk_pair_out *o) {
  put_pts(o->pt_predict_un, t.mvPtPredictUn, n); put_pts(o->pt_predict, t.mvPtPredict, n);
  put_vec(o->status, t.mvStatus, n);
  put_pts(o->pt_gyro_predict_un, t.mvPtGyroPredictUn, n); put_pts(o->pt_gyro_predict, t.mvPtGyroPredict, n);
  put_pts(o->flows_predict_un, t.mvFlowsPredictUn, n);
  if (o->affine)
    for (int i = 0; i < n; ++i) {
      const bool has = i < (int)t.mvAffineDeformationMatrix.size() && !t.mvAffineDeformationMatrix[i].empty();
      for (int k = 0; k < 4; ++k) o->affine[4 * i + k] = has ? t.mvAffineDeformationMatrix[i].at<float>(k / 2, k % 2) : 0.f;
    }
  put_quads(o->corner_flows, t.mvvFlowsPredictCorners, n);
  put_quads(o->pt_corners_un, t.mvvPtPredictCornersUn, n);
  put_quads(o->pt_corners, t.mvvPtPredictCorners, n);
  put_pts(o->pm_pt_un, t.mvPtPredictAfterPatchMatchedUn, n); put_pts(o->pm_pt, t.mvPtPredictAfterPatchMatched, n);
  put_vec(o->pm_status, t.mvStatusAfterPatchMatched, n);
  put_vec(o->pixel_error, t.mvPixelErrorsOfPatchMatched, n);
  put_vec(o->distance, t.mvDistanceBetweenPredictedAndPatchMatched, n);
  put_vec(o->ncc, t.mvNccAfterPatchMatched, n);
  if (o->iters) std::memset(o->iters, 0, sizeof(int32_t) * (size_t)n);  // not observable from outside the reference
  for (int i = 0; i < 9; ++i) {
    o->Rcl[i] = t.mRcl.empty() ? 0.f : t.mRcl.at<float>(i / 3, i % 3);
    o->KRKinv[i] = t.mKRKinv.empty() ? 0.f : t.mKRKinv.at<float>(i / 3, i % 3);
  }
  o->t_gyro_predict = t.mTimeCostGyroPredict; o->t_opt_flow = t.mTimeCostOptFlow; o->t_filter = t.mTimeCostOptFlowResultFilterOut;
}

bool mode_flags(int e_type, bool &gyro_init, bool &illum, bool &affine, bool &regular) {  // composed path only
  switch (e_type) {  // src/gyro_aided_tracker.cpp:384-414
    case GyroAidedTracker::IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION: gyro_init = false; illum = true; affine = true; regular = false; return true;
    case GyroAidedTracker::GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED: gyro_init = true; illum = false; affine = false; regular = false; return true;
    case GyroAidedTracker::GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION: gyro_init = true; illum = true; affine = false; regular = false; return true;
    case GyroAidedTracker::GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION: gyro_init = true; illum = true; affine = true; regular = false; return true;
    case GyroAidedTracker::GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION_REGULAR: gyro_init = true; illum = true; affine = true; regular = true; return true;
    default: return false;
  }
}

int track_one(const pagk_params &prm, const pagk_pair_in &in, pagk_pair_out *out, int n_threads) {
  cv::shim_num_threads() = n_threads > 0 ? n_threads : 1;
  const int N = in.n_keys;
  Scene s;
  fill_scene(s, in);
  const long long llt0 = pagk_eigen_shim::g_llt_calls.load();
  const bool ref_defaults = prm.iterations == 10 && prm.pyramids == 3 && !prm.calc_ncc && !prm.inverse && prm.lambda == 1.0f &&
                            prm.alpha == 0.5f && prm.max_distance == 25;
  const bool needs_pm = prm.e_type != GyroAidedTracker::GYRO_PREDICT;
  if (prm.e_type == GyroAidedTracker::OPENCV_OPTICAL_FLOW_PYR_LK) { out->n_predict = -1; return PAGK_ERR_UNSUPPORTED; }
  if (in.Rcl_override) { out->n_predict = -1; return PAGK_ERR_UNSUPPORTED; }  // TrackFeatures() always integrates the gyro
  if (prm.e_type == GyroAidedTracker::IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION && !in.keys_ref) { out->n_predict = -1; return PAGK_ERR_INVALID; }
  if (ref_defaults || !needs_pm) {
    g_last_path = 0;
    GyroAidedTracker t(s.ref, s.cur, s.calib, s.bias, s.table, (GyroAidedTracker::eType)prm.e_type,
                       (GyroAidedTracker::ePredictMethod)prm.predict_method, "", prm.half_patch);
    const int n_predict = t.TrackFeatures();
    export_tracker(t, N, out);
    out->n_predict = n_predict;
    out->n_iterations = pagk_eigen_shim::g_llt_calls.load() - llt0;
    return n_predict < 0 ? PAGK_ERR_UNSUPPORTED : PAGK_OK;
  }
  // composed path: the same three steps through the public interface, with the requested levels / iterations
  g_last_path = 1;
  bool gyro_init, illum, affine, regular;
  if (!mode_flags(prm.e_type, gyro_init, illum, affine, regular)) { out->n_predict = -1; return PAGK_ERR_UNSUPPORTED; }
  GyroAidedTracker t(s.ref, s.cur, s.calib, s.bias, s.table, GyroAidedTracker::GYRO_PREDICT,
                     (GyroAidedTracker::ePredictMethod)prm.predict_method, "", prm.half_patch);
  if (gyro_init) {
    t.TrackFeatures();  // IntegrateGyroMeasurements + GyroPredictFeatures
  } else {
    // TrackFeatures() integrates the gyro in every mode; do that on a tracker without keypoints, hand Rcl over with
    // the public SetRcl(), then src/gyro_aided_tracker.cpp:264-270
    pagk_pair_in z = in;
    z.n_keys = 0;
    Scene sz;
    fill_scene(sz, z);
    GyroAidedTracker tg(sz.ref, sz.cur, sz.calib, sz.bias, sz.table, GyroAidedTracker::GYRO_PREDICT, GyroAidedTracker::PIXEL_AWARE_PREDICTION, "", 5);
    tg.TrackFeatures();
    t.SetRcl(tg.mRcl);
    for (int i = 0; i < N; ++i) {
      t.mvPtPredictUn[i] = t.mvKeysRefUn[i].pt;
      t.mvPtPredict[i] = t.mvKeysRef[i].pt;
      t.mvStatus[i] = true;
      t.mvFlowsPredictUn[i] = cv::Point2f(0, 0);
      t.mvAffineDeformationMatrix[i] = cv::Mat::eye(2, 2, CV_32F);
    }
  }
  t.mType = (GyroAidedTracker::eType)prm.e_type;
  t.mbHasGyroPredictInitial = gyro_init; t.mbConsiderIllumination = illum; t.mbConsiderAffineDeformation = affine;
  t.mbRegularizationPenalty = regular;
  const int half = t.mHalfPatchSize;
  // the two spans TrackFeatures() would time itself (mTimeCostOptFlow, mTimeCostOptFlowResultFilterOut, seconds)
  const auto tp0 = std::chrono::steady_clock::now();
  {
    PatchMatch pm(&t, half, prm.iterations, prm.pyramids, gyro_init, prm.inverse != 0, illum, affine, regular, prm.calc_ncc != 0);
    pm.OpticalFlowMultiLevel();
  }
  const auto tp1 = std::chrono::steady_clock::now();
  // result filter, src/gyro_aided_tracker.cpp:289-336
  double sum = 0;
  int cnt = 0;
  for (int i = 0; i < N; ++i)
    if (t.mvStatusAfterPatchMatched[i]) { sum += t.mvPixelErrorsOfPatchMatched[i]; cnt++; }
  const double avg = sum / cnt;
  const double thPix = 4.0 * avg > half ? 4.0 * avg : half;
  const double thDist = half * 4.0;
  int n_predict = 0;
  for (int i = 0; i < N; ++i) {
    if (t.mvStatusAfterPatchMatched[i] && t.mvPixelErrorsOfPatchMatched[i] < thPix &&
        t.mvDistanceBetweenPredictedAndPatchMatched[i] < thDist) {
      t.mvPtPredict[i] = t.mvPtPredictAfterPatchMatched[i];
      t.mvPtPredictUn[i] = t.mvPtPredictAfterPatchMatchedUn[i];
      t.mvStatus[i] = true;
      n_predict++;
    } else {
      t.mvStatus[i] = false;
    }
  }
  const auto tp2 = std::chrono::steady_clock::now();
  export_tracker(t, N, out);
  out->t_opt_flow = std::chrono::duration<float>(tp1 - tp0).count();
  out->t_filter = std::chrono::duration<float>(tp2 - tp1).count();
  out->n_predict = n_predict;
  out->n_iterations = pagk_eigen_shim::g_llt_calls.load() - llt0;
  return PAGK_OK;
}

}  // namespace

extern "C" {

int pagk_ref_last_path(void) { return g_last_path; }

// GeometryValidation's two RANSAC estimators are OpenCV's; tests inject the model they should "find"
void pagk_ref_inject_models(const double *H9, const double *F9) {
  g_injected_H = cv::Mat(); g_injected_F = cv::Mat();
  if (H9) { g_injected_H = cv::Mat(3, 3, CV_64F); for (int i = 0; i < 9; ++i) g_injected_H.at<double>(i / 3, i % 3) = H9[i]; }
  if (F9) { g_injected_F = cv::Mat(3, 3, CV_64F); for (int i = 0; i < 9; ++i) g_injected_F.at<double>(i / 3, i % 3) = F9[i]; }
}

int pagk_ref_track(const pagk_params *prm, const pagk_pair_in *in, pagk_pair_out *out, int n_threads) {
  return track_one(*prm, *in, out, n_threads);
}

int pagk_ref_track_batch(const pagk_params *prm, int n_pairs, const pagk_pair_in *in, pagk_pair_out *out, int n_threads) {
  int rc = PAGK_OK;
  for (int p = 0; p < n_pairs; ++p) {
    const int r = track_one(*prm, in[p], &out[p], n_threads);
    if (r != PAGK_OK) rc = r;
  }
  return rc;
}

// IntegrateGyroMeasurements + SetRcl through TrackFeatures() in GYRO_PREDICT mode on zero keypoints
int pagk_ref_integrate_gyro(const pagk_pair_in *in, float *Rcl, float *KRKinv) {
  pagk_pair_in z = *in;
  z.n_keys = 0;
  Scene s;
  fill_scene(s, z);
  GyroAidedTracker t(s.ref, s.cur, s.calib, s.bias, s.table, GyroAidedTracker::GYRO_PREDICT, GyroAidedTracker::PIXEL_AWARE_PREDICTION, "", 5);
  t.TrackFeatures();
  for (int i = 0; i < 9; ++i) { Rcl[i] = t.mRcl.at<float>(i / 3, i % 3); KRKinv[i] = t.mKRKinv.at<float>(i / 3, i % 3); }
  return PAGK_OK;
}

// GyroAidedTracker::GeometryValidation() itself (src/gyro_aided_tracker.cpp:429-508) with cv::findHomography /
// cv::findFundamentalMat standing in as "return the model the caller supplied".  Only mvStatus and the returned inlier
// count are observable from outside; the two scores are locals of the reference.  GeometryValidation appends a line to
// <saveFolderPath>/trackFeatures.txt and timeCost.txt on every call (:493, :502), hence the scratch folder.
int pagk_ref_geometry_validation(int n_pairs, const pagk_geometry_in *in, pagk_geometry_out *out) {
  static const uint8_t dummy[16] = {0};
  for (int p = 0; p < n_pairs; ++p) {
    const pagk_geometry_in &gi = in[p];
    const float K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, dist[5] = {0, 0, 0, 0, 0};
    Scene s;
    fill_scene(s, dummy, dummy, 2, 2, 2, gi.n_keys, gi.keys_ref_un, nullptr, K, dist, 4);
    s.calib.Tbc = cv::Mat::eye(4, 4, CV_32F);
    s.bias = cv::Point3f(0, 0, 0);
    GyroAidedTracker t(s.ref, s.cur, s.calib, s.bias, s.table, GyroAidedTracker::GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION,
                       GyroAidedTracker::PIXEL_AWARE_PREDICTION, "/tmp/pagk_ref_scratch/", 5);
    for (int i = 0; i < gi.n_keys; ++i) {
      t.mvPtPredictUn[i] = cv::Point2f(gi.pt_predict_un[2 * i], gi.pt_predict_un[2 * i + 1]);
      t.mvStatus[i] = gi.status[i];
    }
    pagk_ref_inject_models(gi.H21, gi.F21);
    int cand = 0;
    for (int i = 0; i < gi.n_keys; ++i) cand += gi.status[i] ? 1 : 0;
    // the reference's locals are sized from sigma = 1.0 (:447); any other value is not reachable through it
    if (gi.sigma != 1.0f) return PAGK_ERR_UNSUPPORTED;
    const int n_inlier = t.GeometryValidation();
    out[p].score_H = out[p].score_F = 0.f; out[p].used_H = -1;
    out[p].n_candidates = cand; out[p].n_inlier = n_inlier;
    if (out[p].status) for (int i = 0; i < gi.n_keys; ++i) out[p].status[i] = t.mvStatus[i];
  }
  pagk_ref_inject_models(nullptr, nullptr);
  return PAGK_OK;
}

// Frame::SetPredictKeyPointsAndMask() itself (src/frame.cpp:115-153) on a current and a last Frame filled from the arrays
int pagk_ref_set_predict_keypoints_and_mask(int n_pairs, const pagk_carry_in *in, pagk_carry_out *out) {
  for (int p = 0; p < n_pairs; ++p) {
    const pagk_carry_in &ci = in[p];
    CameraParams cam;
    cam.width = ci.width; cam.height = ci.height;
    Frame last, cur;
    last.mTimeStamp = ci.t_last; cur.mTimeStamp = ci.t_cur;
    cur.mpLastFrame = &last; cur.mpCameraParams = &cam; last.mpCameraParams = &cam;
    cur.mfx = ci.fx; cur.mfy = ci.fy; cur.mcx = ci.cx; cur.mcy = ci.cy;
    cur.mfx_inv = 1.0 / cur.mfx; cur.mfy_inv = 1.0 / cur.mfy;  // as Frame's constructor computes them (src/frame.cpp:71)
    last.mvKeysNormal.resize(ci.n_keys);
    last.mvFlowVelocityInNormalPlane.resize(ci.n_keys);
    cur.mvStatus.resize(ci.n_keys); cur.mvPtPredict.resize(ci.n_keys); cur.mvPtPredictUn.resize(ci.n_keys);
    for (int i = 0; i < ci.n_keys; ++i) {
      last.mvKeysNormal[i].pt = cv::Point2f(ci.keys_normal_last[2 * i], ci.keys_normal_last[2 * i + 1]);
      cur.mvStatus[i] = ci.status[i];
      cur.mvPtPredict[i] = cv::Point2f(ci.pt_predict[2 * i], ci.pt_predict[2 * i + 1]);
      cur.mvPtPredictUn[i] = cv::Point2f(ci.pt_predict_un[2 * i], ci.pt_predict_un[2 * i + 1]);
    }
    cur.mMask = cv::Mat::ones(ci.height, ci.width, CV_8UC1);  // what the constructor / Reset() leave (src/frame.cpp:89, 104)
    cur.SetPredictKeyPointsAndMask();
    pagk_carry_out &o = out[p];
    o.n_out = (int)cur.mvKeysUn.size();
    for (int k = 0; k < o.n_out; ++k) {
      o.keys[2 * k] = cur.mvKeys[k].pt.x; o.keys[2 * k + 1] = cur.mvKeys[k].pt.y;
      o.keys_un[2 * k] = cur.mvKeysUn[k].pt.x; o.keys_un[2 * k + 1] = cur.mvKeysUn[k].pt.y;
      o.keys_normal[2 * k] = cur.mvKeysNormal[k].pt.x; o.keys_normal[2 * k + 1] = cur.mvKeysNormal[k].pt.y;
      o.index_in_last[k] = cur.mvPtIndexInLastFrame[k];
      o.flow_velocity_last[2 * k] = last.mvFlowVelocityInNormalPlane[k].x;
      o.flow_velocity_last[2 * k + 1] = last.mvFlowVelocityInNormalPlane[k].y;
    }
    if (o.mask)
      for (int y = 0; y < ci.height; ++y) std::memcpy(o.mask + (size_t)y * ci.width, cur.mMask.data + (size_t)y * cur.mMask.step, (size_t)ci.width);
  }
  return PAGK_OK;
}

// ORB_SLAM2::ORBextractor(nfeatures, 1.2f, 1, iniThFAST, minThFAST).DetectFeatures(image, mask, keypoints)
// (src/ORBextractor.cc:1148-1205 with ComputeKeyPointsOctTree :789-876): the keypoint top-up of Frame::DetectKeyPoints
// (src/frame.cpp:155-219).  With one level and nfeatures larger than the number of candidates, DistributeOctTree keeps every
// keypoint (each ends in a node of its own), so the output is the set the per-cell FAST calls produced.
int pagk_ref_orb_detect(const uint8_t *img, int width, int height, int pitch, const uint8_t *mask, int nfeatures, int ini_th, int min_th,
                        int max_out, float *xy, float *response, int *n_out) {
  cv::Mat image = image_mat(img, width, height, pitch);
  cv::Mat m;
  if (mask) { m = cv::Mat(height, width, CV_8UC1); std::memcpy(m.data, mask, (size_t)width * height); }
  else m = cv::Mat::ones(height, width, CV_8UC1);
  ORB_SLAM2::ORBextractor ex(nfeatures, 1.2f, 1, ini_th, min_th);
  std::vector<cv::KeyPoint> keys;
  ex.DetectFeatures(image, m, keys);
  *n_out = (int)keys.size();
  for (int k = 0; k < (int)keys.size() && k < max_out; ++k) { xy[2 * k] = keys[k].pt.x; xy[2 * k + 1] = keys[k].pt.y; response[k] = keys[k].response; }
  return PAGK_OK;
}

// ORBextractor::DistributeOctTree itself (src/ORBextractor.cc:563-787; protected, reached through a derived class) on
// caller-given candidates relative to (min_x, min_y): indices of the kept candidates in its output order.  The index travels
// in cv::KeyPoint::class_id, which the function copies along with the keypoint.
namespace {
struct OctTreeAccess : public ORB_SLAM2::ORBextractor {
  OctTreeAccess(int n) : ORB_SLAM2::ORBextractor(n, 1.2f, 1, 20, 7) {}
  std::vector<cv::KeyPoint> run(const std::vector<cv::KeyPoint> &k, int minX, int maxX, int minY, int maxY, int N) {
    return DistributeOctTree(k, minX, maxX, minY, maxY, N, 0);
  }
};
}  // namespace
int pagk_ref_distribute_octtree(int n, const float *xy, const float *response, int min_x, int max_x, int min_y, int max_y,
                                int n_features, int *out_index, int *n_out) {
  std::vector<cv::KeyPoint> keys((size_t)n);
  for (int k = 0; k < n; ++k) {
    keys[(size_t)k].pt = cv::Point2f(xy[2 * k], xy[2 * k + 1]);
    keys[(size_t)k].response = response[k];
    keys[(size_t)k].class_id = k;
  }
  OctTreeAccess ex(n_features);
  const std::vector<cv::KeyPoint> out = ex.run(keys, min_x, max_x, min_y, max_y, n_features);
  *n_out = (int)out.size();
  for (size_t k = 0; k < out.size(); ++k) out_index[k] = out[k].class_id;
  return PAGK_OK;
}

// PatchMatch(&tracker, ...).OpticalFlowMultiLevel() on caller-given predictions, status and deformation matrices
int pagk_ref_patch_match(const pagk_patch_match_in *in, pagk_pair_out *out, int n_threads) {
  cv::shim_num_threads() = n_threads > 0 ? n_threads : 1;
  const int N = in->n_keys;
  Scene s;
  fill_scene(s, in->img_ref, in->img_cur, in->width, in->height, in->pitch, N, in->keys_ref_un, nullptr, in->K, in->dist, in->n_dist);
  s.calib.Tbc = cv::Mat::eye(4, 4, CV_32F);
  s.bias = cv::Point3f(0, 0, 0);
  GyroAidedTracker t(s.ref, s.cur, s.calib, s.bias, s.table, GyroAidedTracker::GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION,
                     GyroAidedTracker::PIXEL_AWARE_PREDICTION, "", in->half_patch);
  for (int i = 0; i < N; ++i) {
    t.mvPtPredictUn[i] = cv::Point2f(in->pt_predict_un[2 * i], in->pt_predict_un[2 * i + 1]);
    t.mvStatus[i] = in->status[i];
    cv::Mat A(2, 2, CV_32F);
    for (int k = 0; k < 4; ++k) A.at<float>(k / 2, k % 2) = in->affine[4 * i + k];
    t.mvAffineDeformationMatrix[i] = A;
  }
  const long long llt0 = pagk_eigen_shim::g_llt_calls.load();
  if (in->lambda != 1.0f || in->alpha != 0.5f || in->max_distance != 25) return PAGK_ERR_UNSUPPORTED;  // fixed in the PatchMatch ctor (:49-51)
  {
    PatchMatch pm(&t, in->half_patch, in->iterations, in->pyramids, in->has_gyro_predict_initial != 0, in->inverse != 0,
                  in->consider_illumination != 0, in->consider_affine_deformation != 0, in->regularization_penalty != 0,
                  in->calc_ncc != 0);
    pm.OpticalFlowMultiLevel();
  }
  export_tracker(t, N, out);
  out->n_predict = 0;
  out->n_iterations = pagk_eigen_shim::g_llt_calls.load() - llt0;
  return PAGK_OK;
}

}  // extern "C"
