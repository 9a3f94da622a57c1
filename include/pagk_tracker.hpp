// pagk_tracker.hpp -- header-only C++ host shim over the C-ABI (include/pagk.h).
//
// It re-creates the two classes a caller of the reference uses on this path, with the reference's
// method names, constructor argument order and public result members, over OpenCV-free value types:
//
//     reference (include/gyro_aided_tracker.h:109-259)        here (namespace pagk)
//     ------------------------------------------------        ---------------------------------------
//     cv::Point2f / cv::Point3f / cv::KeyPoint                 Point2f / Point3f / KeyPoint
//     cv::Mat  (CV_8UC1 image, borrowed)                       ImageView {data, cols, rows, step}
//     cv::Mat  (3x3 / 4x4 / 2x2 CV_32F)                        Mat3f / Mat4f / Mat2f (row-major arrays)
//     IMU::Point, IMU::Calib (imu_types.h:93-144)              ImuPoint, ImuCalib {Tbc}
//     CameraParams (imu_types.h:34-88)                         CameraParams {mK, mDistCoef[4], width, height}
//     Frame fields read/written by the tracker (frame.h:59-91) Frame
//     GyroAidedTracker, PatchMatch                             GyroAidedTracker, PatchMatch
//
// A reference driver (Examples/Demo/RealSenseD435i.cpp:244-254) ports by type substitution:
//
//     pagk::Device dev(0, /*max_width*/752, /*max_height*/480, /*max_keys*/1024);   // once per process
//     pagk::GyroAidedTracker trk(dev, lastFrame, curFrame, imuCalib, biasg, /*normalizeTable*/nullptr,
//                                pagk::GyroAidedTracker::GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION,
//                                pagk::GyroAidedTracker::PIXEL_AWARE_PREDICTION, saveFolderPath, half_patch_size);
//     int n = trk.TrackFeatures();
//     trk.SetBackToFrame(curFrame);
//
// Semantics kept from the reference: inputs are borrowed (they must outlive the tracker); TrackFeatures()
// returns the number of tracked features or -1 for an unsupported eType; per-feature failure is
// mvStatus[i] == 0; timing members hold seconds.  Differences: the only extra constructor argument is the
// Device (the GPU workspace, reusable across frames); saveFolderPath is accepted and ignored (the
// reference only mkdir -p's it, :54-58); GeometryValidation (RANSAC H/F, :429-508) is outside this path.
// Errors from the C-ABI other than "unsupported" are thrown as pagk::Error.
#ifndef PAGK_TRACKER_HPP_
#define PAGK_TRACKER_HPP_

#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "pagk.h"

namespace pagk {

struct Point2f { float x = 0.f, y = 0.f; Point2f() = default; Point2f(float x_, float y_) : x(x_), y(y_) {} };
struct Point3f { float x = 0.f, y = 0.f, z = 0.f; Point3f() = default; Point3f(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {} };
struct KeyPoint { Point2f pt; float size = 0.f, angle = -1.f, response = 0.f; int octave = 0, class_id = -1; };
struct ImageView { const uint8_t *data = nullptr; int cols = 0, rows = 0, step = 0; };
using Mat2f = std::array<float, 4>;
using Mat3f = std::array<float, 9>;
using Mat4f = std::array<float, 16>;
static_assert(sizeof(Point2f) == 8, "Point2f must be two packed floats");

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error("pagk error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) { if (rc != PAGK_OK) throw Error(rc, pagk_last_error()); }

struct ImuPoint {  // IMU::Point
  Point3f a, w;
  double t = 0.0;
  ImuPoint() = default;
  ImuPoint(const Point3f &acc, const Point3f &gyro, double ts) : a(acc), w(gyro), t(ts) {}
};
struct ImuCalib { Mat4f Tbc{}; };  // IMU::Calib: only Tbc is read by the tracker (src/gyro_aided_tracker.cpp:42)
struct CameraParams {
  Mat3f mK{};
  std::array<float, 4> mDistCoef{};  // k1 k2 p1 p2; the reference's CameraParams holds 4, so the tracker sees k3 = 0
  int width = 0, height = 0;
};
struct Frame {
  double mTimeStamp = 0.0;
  ImageView mGray;
  std::vector<KeyPoint> mvKeys, mvKeysUn;
  std::vector<ImuPoint> mvImuFromLastFrame;
  const CameraParams *mpCameraParams = nullptr;
  // written by GyroAidedTracker::SetBackToFrame
  std::vector<Point2f> mvPtGyroPredictUn, mvPtPredict, mvPtPredictUn;
  std::vector<uint8_t> mvStatus;
  std::vector<float> mvNcc;
  std::vector<std::vector<Point2f>> mvvFlowsPredictCorners;
  Mat3f mRcl{};
};

// The GPU workspace (one pagk_handle).  Create once, reuse for every frame pair.
class Device {
 public:
  Device(int device, int max_width, int max_height, int max_keys, int max_levels = 4, int max_half_patch = 10, int max_pairs = 1) {
    pagk_config c{device, max_width, max_height, max_keys, max_pairs, 64, max_levels, max_half_patch};
    check(pagk_create(&c, &h_));
  }
  ~Device() { pagk_destroy(h_); }
  Device(const Device &) = delete;
  Device &operator=(const Device &) = delete;
  pagk_handle *handle() const { return h_; }

 private:
  pagk_handle *h_ = nullptr;
};

class PatchMatch;

class GyroAidedTracker {
 public:
  enum eType {
    OPENCV_OPTICAL_FLOW_PYR_LK = 0,
    GYRO_PREDICT = 1,
    GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED = 2,
    GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION = 3,
    GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION = 4,
    GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION_REGULAR = 6,
    IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION = 5
  };
  enum ePredictMethod { PIXEL_AWARE_PREDICTION = 1, SINGLE_HOMOGRAPHY = 2 };

  // reference src/gyro_aided_tracker.cpp:30-49
  GyroAidedTracker(Device &dev, const Frame &pFrameRef, const Frame &pFrameCur, const ImuCalib &imuCalib,
                   const Point3f &biasg_, const float *normalizeTable /* H x W x 2 or nullptr */,
                   eType type_ = GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION,
                   ePredictMethod predictMethod_ = PIXEL_AWARE_PREDICTION, std::string saveFolderPath = "",
                   int halfPatchSize_ = 5)
      : mType(type_), mPredictMethod(predictMethod_), mSaveFolderPath(std::move(saveFolderPath)),
        mTimeStamp(pFrameCur.mTimeStamp), mTimeStampRef(pFrameRef.mTimeStamp), mImgGrayRef(pFrameRef.mGray),
        mImgGrayCur(pFrameCur.mGray), mvKeysRef(pFrameRef.mvKeys), mvKeysRefUn(pFrameRef.mvKeysUn),
        mvImuFromLastFrame(pFrameCur.mvImuFromLastFrame), mBias(biasg_), mHalfPatchSize(halfPatchSize_ == 0 ? 5 : halfPatchSize_),
        mK(pFrameCur.mpCameraParams->mK), mDistCoef(pFrameCur.mpCameraParams->mDistCoef),
        mWidth(pFrameCur.mpCameraParams->width), mHeight(pFrameCur.mpCameraParams->height), mN((int)pFrameRef.mvKeys.size()),
        dev_(dev), normalize_table_(normalizeTable) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) mRbc[r * 3 + c] = imuCalib.Tbc[r * 4 + c];
    Initialize();
  }

  void Initialize() {  // reference :51-95 (sizes the result vectors)
    const size_t n = (size_t)mN;
    mvPtPredict.assign(n, Point2f()); mvPtPredictUn.assign(n, Point2f()); mvFlowsPredictUn.assign(n, Point2f());
    mvStatus.assign(n, 0);
    mvvPtPredictCorners.assign(n, {}); mvvPtPredictCornersUn.assign(n, {}); mvvFlowsPredictCorners.assign(n, {});
    mvAffineDeformationMatrix.assign(n, Mat2f{});
    keys_un_.resize(n * 2); keys_.resize(n * 2);
    for (size_t i = 0; i < n; ++i) {
      keys_un_[2 * i] = mvKeysRefUn[i].pt.x; keys_un_[2 * i + 1] = mvKeysRefUn[i].pt.y;
      keys_[2 * i] = mvKeysRef[i].pt.x; keys_[2 * i + 1] = mvKeysRef[i].pt.y;
    }
    imu_t_.clear(); imu_w_.clear();
    for (const ImuPoint &p : mvImuFromLastFrame) { imu_t_.push_back(p.t); imu_w_.push_back(p.w.x); imu_w_.push_back(p.w.y); imu_w_.push_back(p.w.z); }
  }

  void SetRegularizationPenalty(bool flag) { mbRegularizationPenalty = flag; }  // overwritten by TrackFeatures(), as in the reference
  void SetType(eType type_) { mType = type_; }
  void SetRcl(const Mat3f &Rcl_) { mRcl = Rcl_; has_rcl_ = true; }
  Mat3f GetRcl() const { return mRcl; }
  // PatchMatch parameters the reference hard-codes at :276-278; exposed because BASELINE configs use 4-5 levels
  void SetPyramids(int levels) { pyramids_ = levels; }
  void SetIterations(int it) { iterations_ = it; }

  int TrackFeatures() {  // reference :344-426
    pagk_params prm;
    pagk_default_params(&prm);
    prm.e_type = (int)mType; prm.predict_method = (int)mPredictMethod; prm.half_patch = mHalfPatchSize;
    prm.iterations = iterations_; prm.pyramids = pyramids_;
    pagk_pair_in in = make_in();
    Staging st((size_t)mN);
    pagk_pair_out out = st.out();
    const int rc = pagk_track_batch(dev_.handle(), &prm, 1, &in, &out);
    if (rc == PAGK_ERR_UNSUPPORTED) return -1;  // "Unsupport type!!! return -1;" (:415-418)
    check(rc);
    st.store(*this, out, true);
    return out.n_predict;
  }

  // int GeometryValidation(), reference :429-508.  The reference estimates H21 / F21 itself with cv::findHomography and
  // cv::findFundamentalMat (OpenCV RANSAC).  GeometryValidation() does the same on the device (pagk's own RANSAC: not
  // OpenCV's models bit for bit); GeometryValidation(H21, F21) takes what those two OpenCV calls returned for the status-1
  // correspondences (mvKeysRefUn[i].pt -> mvPtPredictUn[i]); scoring, model choice and outlier marking run on the GPU.
  int GeometryValidation(unsigned int seed = 1, float sigma = 1.0f) { return GeometryValidation(nullptr, nullptr, sigma, seed); }
  int GeometryValidation(const double H21[9], const double F21[9], float sigma = 1.0f, unsigned int seed = 1) {
    pagk_geometry_in gi;
    std::vector<float> k1((size_t)mN * 2), k2((size_t)mN * 2);
    for (int i = 0; i < mN; ++i) {
      k1[2 * (size_t)i] = mvKeysRefUn[(size_t)i].pt.x; k1[2 * (size_t)i + 1] = mvKeysRefUn[(size_t)i].pt.y;
      k2[2 * (size_t)i] = mvPtPredictUn[(size_t)i].x; k2[2 * (size_t)i + 1] = mvPtPredictUn[(size_t)i].y;
    }
    gi.n_keys = mN; gi.keys_ref_un = k1.data(); gi.pt_predict_un = k2.data(); gi.status = mvStatus.data();
    for (int i = 0; i < 9; ++i) { gi.H21[i] = H21 ? H21[i] : 0.0; gi.F21[i] = F21 ? F21[i] : 0.0; }
    gi.sigma = sigma; gi.estimate = (H21 && F21) ? 0 : 1; gi.seed = seed; gi.reserved = 0;
    std::vector<uint8_t> st((size_t)mN);
    pagk_geometry_out go;
    go.status = st.data();
    check(pagk_geometry_validation(dev_.handle(), 1, &gi, &go));
    mvStatus = st;
    mGeometryScoreH = go.score_H; mGeometryScoreF = go.score_F; mGeometryUsedH = go.used_H != 0;
    return go.n_inlier;
  }
  float mGeometryScoreH = 0.f, mGeometryScoreF = 0.f;  // locals score_H / score_F of the reference (:448)
  bool mGeometryUsedH = false;

  void SetBackToFrame(Frame &pFrame) const {  // reference :97-111
    pFrame.mvPtGyroPredictUn = mvPtGyroPredictUn;
    pFrame.mvPtPredict = mvPtPredict;
    pFrame.mvPtPredictUn = mvPtPredictUn;
    pFrame.mvStatus = mvStatus;
    pFrame.mvNcc = mvNccAfterPatchMatched;
    pFrame.mvvFlowsPredictCorners = mvvFlowsPredictCorners;
    pFrame.mRcl = mRcl;
  }

  // ---- public members, names as in include/gyro_aided_tracker.h:173-259 ----
  eType mType;
  ePredictMethod mPredictMethod;
  std::string mSaveFolderPath;
  double mTimeStamp, mTimeStampRef;
  const ImageView &mImgGrayRef, &mImgGrayCur;
  const std::vector<KeyPoint> &mvKeysRef, &mvKeysRefUn;
  const std::vector<ImuPoint> &mvImuFromLastFrame;
  const Point3f &mBias;
  std::vector<Point2f> mvPtPredict, mvPtPredictUn, mvPtGyroPredict, mvPtGyroPredictUn;
  std::vector<std::vector<Point2f>> mvvPtPredictCorners, mvvPtPredictCornersUn, mvvFlowsPredictCorners;
  std::vector<uint8_t> mvStatus;
  int mHalfPatchSize;
  std::vector<Mat2f> mvAffineDeformationMatrix;
  std::vector<Point2f> mvPtPredictAfterPatchMatched, mvPtPredictAfterPatchMatchedUn;
  std::vector<uint8_t> mvStatusAfterPatchMatched;
  std::vector<double> mvPixelErrorsOfPatchMatched, mvDistanceBetweenPredictedAndPatchMatched;
  std::vector<float> mvNccAfterPatchMatched;
  std::vector<Point2f> mvFlowsPredictUn;
  float mTimeCostGyroPredict = 0, mTimeCostOptFlow = 0, mTimeCostOptFlowResultFilterOut = 0;
  Mat3f mRbc{}, mRcl{}, mK, mKRKinv{};
  std::array<float, 4> mDistCoef;
  int mWidth, mHeight, mN;
  bool mbHasGyroPredictInitial = true, mbConsiderIllumination = true, mbConsiderAffineDeformation = false,
       mbRegularizationPenalty = false;
  long long mIterations = 0;  // not in the reference: Gauss-Newton passes executed (the throughput unit)

 private:
  friend class PatchMatch;
  struct Staging {  // flat result buffers of one call
    size_t n;
    std::vector<float> p_un, p, g_un, g, fl, aff, cfl, c_un, c, pm_un, pm, ncc;
    std::vector<uint8_t> st, pm_st;
    std::vector<double> perr, dist;
    explicit Staging(size_t n_) : n(n_), p_un(2 * n_), p(2 * n_), g_un(2 * n_), g(2 * n_), fl(2 * n_), aff(4 * n_), cfl(8 * n_),
                                  c_un(8 * n_), c(8 * n_), pm_un(2 * n_), pm(2 * n_), ncc(n_), st(n_), pm_st(n_), perr(n_), dist(n_) {}
    pagk_pair_out out() {
      pagk_pair_out o;
      std::memset(&o, 0, sizeof(o));
      o.pt_predict_un = p_un.data(); o.pt_predict = p.data(); o.status = st.data(); o.pt_gyro_predict_un = g_un.data();
      o.pt_gyro_predict = g.data(); o.flows_predict_un = fl.data(); o.affine = aff.data(); o.corner_flows = cfl.data();
      o.pt_corners_un = c_un.data(); o.pt_corners = c.data(); o.pm_pt_un = pm_un.data(); o.pm_pt = pm.data();
      o.pm_status = pm_st.data(); o.pixel_error = perr.data(); o.distance = dist.data(); o.ncc = ncc.data();
      return o;
    }
    static void pts(std::vector<Point2f> &dst, const std::vector<float> &src, size_t n) {
      dst.resize(n);
      if (n) std::memcpy(dst.data(), src.data(), n * sizeof(Point2f));
    }
    static void corners(std::vector<std::vector<Point2f>> &dst, const std::vector<float> &src, const std::vector<uint8_t> &ok, size_t n) {
      dst.assign(n, {});
      for (size_t i = 0; i < n; ++i)
        if (ok[i]) { dst[i].resize(4); std::memcpy(dst[i].data(), &src[8 * i], 4 * sizeof(Point2f)); }
    }
    void store_pm(GyroAidedTracker &t) const {
      pts(t.mvPtPredictAfterPatchMatched, pm, n); pts(t.mvPtPredictAfterPatchMatchedUn, pm_un, n);
      t.mvStatusAfterPatchMatched = pm_st; t.mvPixelErrorsOfPatchMatched = perr;
      t.mvDistanceBetweenPredictedAndPatchMatched = dist; t.mvNccAfterPatchMatched = ncc;
    }
    void store(GyroAidedTracker &t, const pagk_pair_out &o, bool with_pm) const {
      pts(t.mvPtPredictUn, p_un, n); pts(t.mvPtPredict, p, n); pts(t.mvPtGyroPredictUn, g_un, n); pts(t.mvPtGyroPredict, g, n);
      pts(t.mvFlowsPredictUn, fl, n);
      t.mvStatus = st;
      // the corner vectors and A exist only for features that passed the gyro border test (:131-168); without a gyro
      // prediction (IMAGE_ONLY_..., :264-270) the reference leaves the corner vectors empty and the kernels write A = identity
      std::vector<uint8_t> gyro_ok(n);
      const bool predicted = t.mType != GyroAidedTracker::IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION;
      for (size_t i = 0; i < n; ++i)
        gyro_ok[i] = predicted && (aff[4 * i] != 0.f || aff[4 * i + 1] != 0.f || aff[4 * i + 2] != 0.f || aff[4 * i + 3] != 0.f);
      corners(t.mvvFlowsPredictCorners, cfl, gyro_ok, n); corners(t.mvvPtPredictCornersUn, c_un, gyro_ok, n);
      corners(t.mvvPtPredictCorners, c, gyro_ok, n);
      t.mvAffineDeformationMatrix.resize(n);
      if (n) std::memcpy(t.mvAffineDeformationMatrix.data(), aff.data(), n * sizeof(Mat2f));
      if (with_pm) store_pm(t);
      std::memcpy(t.mRcl.data(), o.Rcl, sizeof(o.Rcl)); std::memcpy(t.mKRKinv.data(), o.KRKinv, sizeof(o.KRKinv));
      t.mTimeCostGyroPredict = o.t_gyro_predict; t.mTimeCostOptFlow = o.t_opt_flow; t.mTimeCostOptFlowResultFilterOut = o.t_filter;
      t.mIterations = o.n_iterations;
    }
  };

  pagk_pair_in make_in() const {
    pagk_pair_in in;
    std::memset(&in, 0, sizeof(in));
    in.img_ref = mImgGrayRef.data; in.img_cur = mImgGrayCur.data;
    in.width = mWidth; in.height = mHeight; in.pitch = mImgGrayCur.step ? mImgGrayCur.step : mWidth;
    in.n_keys = mN; in.keys_ref_un = keys_un_.data(); in.keys_ref = keys_.data();
    in.n_imu = (int)imu_t_.size(); in.imu_t = imu_t_.data(); in.imu_w = imu_w_.data();
    in.t_ref = mTimeStampRef; in.t_cur = mTimeStamp;
    in.bias_g[0] = mBias.x; in.bias_g[1] = mBias.y; in.bias_g[2] = mBias.z;
    std::memcpy(in.K, mK.data(), sizeof(in.K));
    std::memcpy(in.dist, mDistCoef.data(), 4 * sizeof(float)); in.dist[4] = 0.f; in.n_dist = 4;
    std::memcpy(in.Rbc, mRbc.data(), sizeof(in.Rbc));
    in.normalize_table = normalize_table_;
    in.Rcl_override = has_rcl_ ? mRcl.data() : nullptr;
    return in;
  }

  Device &dev_;
  const float *normalize_table_;
  std::vector<float> keys_un_, keys_, imu_w_;
  std::vector<double> imu_t_;
  bool has_rcl_ = false;
  int pyramids_ = 3, iterations_ = 10;
};

// PatchMatch(pMatcher, halfPatchSize, iterations, pyramids, bHasGyroPredictInitial, bInverse, bConsiderIllumination,
//            bConsiderAffineDeformation, bRegularizationPenalty = true, bCalculateNCC = false)
// -- reference include/patch_match.h:44-49.  OpticalFlowMultiLevel() (:79-142) reads mvKeysRefUn, mvPtPredictUn,
// mvStatus and mvAffineDeformationMatrix of the tracker and writes the six SetMatcher vectors back (:370-388).
class PatchMatch {
 public:
  PatchMatch(GyroAidedTracker *pMatcher_, int halfPatchSize_, int iterations_, int pyramids_, bool bHasGyroPredictInitial_,
             bool bInverse_, bool bConsiderIllumination_, bool bConsiderAffineDeformation_, bool bRegularizationPenalty_ = true,
             bool bCalculateNCC_ = false)
      : mpMatcher(pMatcher_), mvGyroPredictStatus(pMatcher_->mvStatus) {
    std::memset(&in_, 0, sizeof(in_));
    in_.half_patch = halfPatchSize_; in_.iterations = iterations_; in_.pyramids = pyramids_;
    in_.has_gyro_predict_initial = bHasGyroPredictInitial_; in_.inverse = bInverse_;
    in_.consider_illumination = bConsiderIllumination_; in_.consider_affine_deformation = bConsiderAffineDeformation_;
    in_.regularization_penalty = bRegularizationPenalty_; in_.calc_ncc = bCalculateNCC_;
    in_.lambda = 1.0f; in_.alpha = 0.5f; in_.max_distance = 25;  // src/patch_match.cpp:48-50
  }

  void OpticalFlowMultiLevel() {
    GyroAidedTracker &t = *mpMatcher;
    const size_t n = (size_t)t.mN;
    std::vector<float> pred(2 * n), aff(4 * n);
    for (size_t i = 0; i < n; ++i) { pred[2 * i] = t.mvPtPredictUn[i].x; pred[2 * i + 1] = t.mvPtPredictUn[i].y; }
    if (n) std::memcpy(aff.data(), t.mvAffineDeformationMatrix.data(), n * sizeof(Mat2f));
    in_.img_ref = t.mImgGrayRef.data; in_.img_cur = t.mImgGrayCur.data;
    in_.width = t.mWidth; in_.height = t.mHeight; in_.pitch = t.mImgGrayCur.step ? t.mImgGrayCur.step : t.mWidth;
    in_.n_keys = t.mN; in_.keys_ref_un = t.keys_un_.data(); in_.pt_predict_un = pred.data();
    in_.status = mvGyroPredictStatus.data(); in_.affine = aff.data();
    std::memcpy(in_.K, t.mK.data(), sizeof(in_.K));
    std::memcpy(in_.dist, t.mDistCoef.data(), 4 * sizeof(float)); in_.dist[4] = 0.f; in_.n_dist = 4;
    GyroAidedTracker::Staging st(n);
    pagk_pair_out out = st.out();
    check(pagk_patch_match(t.dev_.handle(), &in_, &out));
    st.store_pm(t);
    t.mIterations = out.n_iterations;
  }

 private:
  GyroAidedTracker *mpMatcher;
  std::vector<uint8_t> mvGyroPredictStatus;
  pagk_patch_match_in in_;
};

}  // namespace pagk

#endif  // PAGK_TRACKER_HPP_
