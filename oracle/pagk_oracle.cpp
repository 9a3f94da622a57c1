// pagk_oracle.cpp -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
// this library.  Nothing under pixel_aware_gyro_aided_klt_feature_tracker_b200/ links, imports or
// calls it; the product path fails loudly when its CUDA library is missing.
//
// What it restates (file:line relative to /root/reference):
//   src/gyro_aided_tracker.cpp:51-95    Initialize()
//   src/gyro_aided_tracker.cpp:118-256  GyroPredictFeatures / GyroPredictOnePixel
//   src/gyro_aided_tracker.cpp:258-342  GyroPredictFeaturesAndOpticalFlowRefined (orchestration, filter)
//   src/gyro_aided_tracker.cpp:344-426  TrackFeatures (eType -> mode flags)
//   src/gyro_aided_tracker.cpp:511-587  SetRcl / IntegrateGyroMeasurements / IntegrateOneGyroMeasurement
//   src/patch_match.cpp:33-142          PatchMatch ctor, CreatePyramids, OpticalFlowMultiLevel
//   src/patch_match.cpp:167-367         OpticalFlowConsideringIlluminationChange_onePixel
//   src/patch_match.cpp:370-416,433-469 SetMatcher, GetPixelValue, DistortPoints, NCC
//   src/utils.cpp:49-76                 DistortVecPoints
//
// The reference cannot be built as shipped in this image: it needs OpenCV (>= 3.4), Eigen3 and glog, none of which
// are installed, and there is no network.  Two things stand in for it:
//   (1) this file, a restatement that travels to the GPU box in source form;
//   (2) oracle/_ref/libpagk_ref.so (oracle/reference.py, oracle/ref_harness.cpp): the reference's OWN
//       src/gyro_aided_tracker.cpp, src/patch_match.cpp and src/utils.cpp compiled unmodified, where they lie,
//       against stand-in OpenCV / Eigen3 / glog headers (oracle/ref_shim/).  tests/test_reference_build.py holds this
//       file bit-for-bit equal to it on every output (all eTypes, 1-5 levels, 7x7..21x21 patches, distortion,
//       normalize table, border features, flat regions, empty inputs), and tests/golden/lk_frozen.npz freezes its
//       outputs.  So the Gauss-Newton loop, the prediction, the filter and every implicit float/double promotion of
//       the reference's C++ are PINNED against the reference's own code.
// Third-party arithmetic, restated in both (1) and (2):
//   * cv::resize INTER_LINEAR u8        -> pagk_cv_resize.h   PINNED bit-exact against cv2 4.13 fixtures
//   * cv::Mat gemm / invert / scaleAdd  -> small_*()          PINNED bit-exact against cv2 4.13 fixtures
//     (the stand-in cv::MatExpr of (2) is pinned against the same fixtures through the reference's own expressions)
//   * Eigen::Matrix4d::llt().solve(), Vector4d::norm() -> llt_solve4()   PARITY UNPINNED: Eigen 3.3.4 (Ubuntu 18.04,
//                                          README.md:21) restated from its published algorithm
//                                          (llt_inplace<double,Lower>::unblocked, fixed-size triangular
//                                          solver unrollers, SSE2 redux order).  No Eigen here to check.
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math (x86-64 baseline: SSE2, no FMA), which is the
// arithmetic of the reference's own -O3 build (CMakeLists.txt:10-11: no -march, no -ffast-math).
//
// Out-of-bounds convention.  PatchMatch::GetPixelValue (src/patch_match.cpp:391-406) reads the taps
// data[1], data[step], data[step+1] even when the clamped coordinate sits in the last column or row;
// for y in (rows-1, rows) that is a read past the cv::Mat (undefined behaviour in the reference).
// The oracle and the CUDA path both store every pyramid level as a continuous buffer (step == cols)
// followed by one guard row that replicates the last row, plus one more byte (= first byte of the
// guard row).  Column overflow therefore wraps into the next row exactly as in a continuous cv::Mat.

#include "../include/pagk.h"
#include "pagk_cv_resize.h"
#include "pagk_cv_fast.h"

#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>
#include <chrono>

namespace {

struct P2 { float x, y; };

// ---------------------------------------------------------------------------------------------
// Pyramid: cv::resize(src, dst, Size(cols*0.5, rows*0.5)) with the default INTER_LINEAR
// (src/patch_match.cpp:69-70).  OpenCV switches to the INTER_AREA 2x2 fast path when both scale
// factors are exactly 2, otherwise runs the 11-bit fixed point bilinear.
// ---------------------------------------------------------------------------------------------
struct Level {
  int cols = 0, rows = 0;
  std::vector<uint8_t> buf;  // (rows+1)*cols + 1 bytes, see the out-of-bounds convention above
  const uint8_t *data() const { return buf.data(); }
};

void finish_guard(Level &L) {
  const size_t c = (size_t)L.cols, r = (size_t)L.rows;
  memcpy(&L.buf[r * c], &L.buf[(r - 1) * c], c);
  L.buf[(r + 1) * c] = L.buf[r * c];
}

void level_from_image(const uint8_t *img, int cols, int rows, int pitch, Level &L) {
  L.cols = cols; L.rows = rows;
  L.buf.assign((size_t)(rows + 1) * cols + 1, 0);
  for (int y = 0; y < rows; ++y) memcpy(&L.buf[(size_t)y * cols], img + (size_t)y * pitch, (size_t)cols);
  finish_guard(L);
}

using pagk_cv::resize_half;

void build_pyramid(const uint8_t *img, int cols, int rows, int pitch, int levels, std::vector<Level> &pyr) {
  pyr.resize(levels);
  level_from_image(img, cols, rows, pitch, pyr[0]);
  for (int l = 1; l < levels; ++l) {
    const Level &S = pyr[l - 1];
    Level &D = pyr[l];
    D.cols = (int)(S.cols * 0.5);  // cv::Size(cols * mPyramidScale, rows * mPyramidScale), int truncation
    D.rows = (int)(S.rows * 0.5);
    D.buf.assign((size_t)(D.rows + 1) * D.cols + 1, 0);
    resize_half(S.data(), S.cols, S.rows, S.cols, D.buf.data(), D.cols, D.rows);
    finish_guard(D);
  }
}

// PatchMatch::GetPixelValue, src/patch_match.cpp:391-406 (the member; utils.h has a different one)
inline float get_pixel_value(const Level &img, float x, float y) {
  if (x < 0) x = 0;
  if (y < 0) y = 0;
  if (x >= img.cols) x = (float)(img.cols - 1);
  if (y >= img.rows) y = (float)(img.rows - 1);
  const uint8_t *data = img.data() + (size_t)((int)y) * img.cols + (int)x;
  const float xx = x - std::floor(x), yy = y - std::floor(y);
  const float a = 1.0f - xx, b = 1.0f - yy;
  const float top = a * (float)data[0] + xx * (float)data[1];
  const float bot = a * (float)data[img.cols] + xx * (float)data[img.cols + 1];
  return b * top + yy * bot;
}

// ---------------------------------------------------------------------------------------------
// cv::Mat small-matrix arithmetic as OpenCV evaluates it for CV_32F (SURVEY.md appendix B,
// re-verified against cv2 4.13 by tests/golden/make_golden.py).
// ---------------------------------------------------------------------------------------------
// gemm without transpose flags, inner dim 2..4: float accumulation, left to right, no FMA.
void small_gemm_nn(const float *A, const float *B, float *D, int m, int n, int k) {
  std::vector<float> tmp((size_t)m * n);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      float s = A[i * k] * B[j];
      for (int t = 1; t < k; ++t) s = s + A[i * k + t] * B[t * n + j];
      tmp[(size_t)i * n + j] = s;
    }
  memcpy(D, tmp.data(), sizeof(float) * m * n);
}
// gemm with a transpose flag: double accumulation (4-way unrolled partial sums), one cast to float.
// a(i,t), b(t,j) are fetched through the given strides so any transpose combination can be expressed.
void small_gemm_dbl(const float *A, int a_rs, int a_cs, const float *B, int b_rs, int b_cs, float *D, int m, int n,
                    int k) {
  std::vector<float> tmp((size_t)m * n);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      int t = 0;
      for (; t <= k - 4; t += 4) {
        s0 += (double)A[i * a_rs + t * a_cs] * (double)B[t * b_rs + j * b_cs];
        s1 += (double)A[i * a_rs + (t + 1) * a_cs] * (double)B[(t + 1) * b_rs + j * b_cs];
        s2 += (double)A[i * a_rs + (t + 2) * a_cs] * (double)B[(t + 2) * b_rs + j * b_cs];
        s3 += (double)A[i * a_rs + (t + 3) * a_cs] * (double)B[(t + 3) * b_rs + j * b_cs];
      }
      for (; t < k; ++t) s0 += (double)A[i * a_rs + t * a_cs] * (double)B[t * b_rs + j * b_cs];
      s0 += s1 + s2 + s3;
      tmp[(size_t)i * n + j] = (float)s0;
    }
  memcpy(D, tmp.data(), sizeof(float) * m * n);
}
// cv::invert of a 3x3 CV_32F (DECOMP_LU special case): adjugate in double times 1/det, cast to float.
bool small_inv3(const float *S, float *D) {
  const double s00 = S[0], s01 = S[1], s02 = S[2], s10 = S[3], s11 = S[4], s12 = S[5], s20 = S[6], s21 = S[7],
               s22 = S[8];
  double det = s00 * (s11 * s22 - s12 * s21) - s01 * (s10 * s22 - s12 * s20) + s02 * (s10 * s21 - s11 * s20);
  if (det == 0.) { memset(D, 0, 9 * sizeof(float)); return false; }
  const double d = 1. / det;
  float t[9];
  t[0] = (float)((s11 * s22 - s12 * s21) * d);
  t[1] = (float)((s02 * s21 - s01 * s22) * d);
  t[2] = (float)((s01 * s12 - s02 * s11) * d);
  t[3] = (float)((s12 * s20 - s10 * s22) * d);
  t[4] = (float)((s00 * s22 - s02 * s20) * d);
  t[5] = (float)((s02 * s10 - s00 * s12) * d);
  t[6] = (float)((s10 * s21 - s11 * s20) * d);
  t[7] = (float)((s01 * s20 - s00 * s21) * d);
  t[8] = (float)((s00 * s11 - s01 * s10) * d);
  memcpy(D, t, sizeof(t));
  return true;
}
bool small_inv2(const float *S, float *D) {
  const double det = (double)S[0] * S[3] - (double)S[1] * S[2];
  if (det == 0.) { memset(D, 0, 4 * sizeof(float)); return false; }
  const double d = 1. / det;
  const float t0 = (float)((double)S[3] * d), t1 = (float)(-(double)S[1] * d), t2 = (float)(-(double)S[2] * d),
              t3 = (float)((double)S[0] * d);
  D[0] = t0; D[1] = t1; D[2] = t2; D[3] = t3;
  return true;
}

// ---------------------------------------------------------------------------------------------
// Gyro integration, src/gyro_aided_tracker.cpp:521-587 and SetRcl :511-519
// ---------------------------------------------------------------------------------------------
void integrate_one(const float w[3], const float bias[3], double dt, float dRout[9]) {
  const float x = (float)((double)(w[0] - bias[0]) * dt);
  const float y = (float)((double)(w[1] - bias[1]) * dt);
  const float z = (float)((double)(w[2] - bias[2]) * dt);
  const float d2 = x * x + y * y + z * z;
  const float d = std::sqrt(d2);
  const float W[9] = {0.f, -z, y, z, 0.f, -x, -y, x, 0.f};
  if ((double)d < 1e-4) {
    for (int i = 0; i < 9; ++i) dRout[i] = ((i % 4 == 0) ? 1.f : 0.f) + W[i];  // I + W
    return;
  }
  // I + W*sin(d)/d + W*W*(1-cos(d))/d2 as cv::MatExpr lowers it: scaleAdd(W, a, I), gemm(W, W, b), add.
  const float a = (float)((double)std::sin(d) * (1.0 / (double)d));
  const float b = (float)((double)(1.0f - std::cos(d)) * (1.0 / (double)d2));
  float T1[9], S[9];
  for (int i = 0; i < 9; ++i) T1[i] = W[i] * a + ((i % 4 == 0) ? 1.f : 0.f);
  small_gemm_nn(W, W, S, 3, 3, 3);
  for (int i = 0; i < 9; ++i) dRout[i] = T1[i] + S[i] * b;
}

void set_rcl(const float K[9], const float Rcl[9], float KRKinv[9]) {
  float Kinv[9], KR[9];
  small_inv3(K, Kinv);
  small_gemm_nn(K, Rcl, KR, 3, 3, 3);
  small_gemm_nn(KR, Kinv, KRKinv, 3, 3, 3);
}

void integrate_gyro(const pagk_pair_in &in, float Rcl[9], float KRKinv[9]) {
  if (in.Rcl_override) {
    memcpy(Rcl, in.Rcl_override, 9 * sizeof(float));
    set_rcl(in.K, Rcl, KRKinv);
    return;
  }
  float dR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  const int n = in.n_imu - 1;
  for (int i = 0; i < n; ++i) {
    float tstep = 0.f;
    float av[3] = {0.f, 0.f, 0.f};
    const float *w0 = in.imu_w + 3 * i, *w1 = in.imu_w + 3 * (i + 1);
    const double t0 = in.imu_t[i], t1 = in.imu_t[i + 1];
    if ((i == 0) && (i < (n - 1))) {
      const float tab = (float)(t1 - t0);
      const float tini = (float)(t0 - in.t_ref);
      const float r = tini / tab;
      for (int c = 0; c < 3; ++c) av[c] = ((w0[c] + w1[c]) - (w1[c] - w0[c]) * r) * 0.5f;
      tstep = (float)(t1 - in.t_ref);
    } else if (i < (n - 1)) {
      for (int c = 0; c < 3; ++c) av[c] = (w0[c] + w1[c]) * 0.5f;
      tstep = (float)(t1 - t0);
    } else if ((i > 0) && (i == (n - 1))) {
      const float tab = (float)(t1 - t0);
      const float tend = (float)(t1 - in.t_cur);
      const float r = tend / tab;
      for (int c = 0; c < 3; ++c) av[c] = ((w0[c] + w1[c]) - (w1[c] - w0[c]) * r) * 0.5f;
      tstep = (float)(in.t_cur - t0);
    } else if ((i == 0) && (i == (n - 1))) {
      for (int c = 0; c < 3; ++c) av[c] = w0[c];
      tstep = (float)(in.t_cur - in.t_ref);
    }
    float d[9];
    integrate_one(av, in.bias_g, (double)tstep, d);
    small_gemm_nn(dR, d, dR, 3, 3, 3);  // dR_ref_cur *= deltaR
  }
  // Rcl = Rbc.t() * dR.t() * Rbc : first product is a GEMM_1_T|GEMM_2_T gemm (double accumulate),
  // second a plain small gemm (float accumulate).
  float M1[9];
  small_gemm_dbl(in.Rbc, 1, 3, dR, 1, 3, M1, 3, 3, 3);  // a(i,t)=Rbc[t][i], b(t,j)=dR[j][t]
  small_gemm_nn(M1, in.Rbc, Rcl, 3, 3, 3);
  set_rcl(in.K, Rcl, KRKinv);
}

// ---------------------------------------------------------------------------------------------
// Tracker state = the members GyroAidedTracker::Initialize caches (src/gyro_aided_tracker.cpp:51-95)
// ---------------------------------------------------------------------------------------------
struct Cam {
  float fx, fy, cx, cy, fx_inv, fy_inv, k1, k2, p1, p2, k3;
  int width, height;
};
Cam make_cam(const float K[9], const float dist[5], int n_dist, int width, int height) {
  Cam c;
  c.fx = K[0]; c.fy = K[4]; c.cx = K[2]; c.cy = K[5];
  c.fx_inv = (float)(1.0 / c.fx); c.fy_inv = (float)(1.0 / c.fy);
  c.k1 = dist[0]; c.k2 = dist[1]; c.p1 = dist[2]; c.p2 = dist[3];
  c.k3 = (n_dist == 5) ? dist[4] : 0.f;
  c.width = width; c.height = height;
  return c;
}

inline P2 distort_point(const Cam &c, P2 p) {
  const float x = (p.x - c.cx) * c.fx_inv;
  const float y = (p.y - c.cy) * c.fy_inv;
  const float r2 = x * x + y * y;
  const float r4 = r2 * r2;
  const float r6 = r4 * r2;
  const float rad = 1 + c.k1 * r2 + c.k2 * r4 + c.k3 * r6;
  const float xd = x * rad + 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x);
  const float yd = y * rad + c.p1 * (r2 + 2 * y * y) + 2 * c.p2 * x * y;
  P2 o;
  o.x = c.fx * xd + c.cx;
  o.y = c.fy * yd + c.cy;
  return o;
}

// GyroPredictOnePixel, src/gyro_aided_tracker.cpp:194-256
void gyro_predict_one(const Cam &c, const float Rcl[9], const float M[9], int method, const float *ntab, P2 ref,
                      P2 &pred, P2 &pred_dist, P2 &flow) {
  float xn, yn;
  if (ntab) {
    const size_t o = ((size_t)((int)ref.y) * c.width + (int)ref.x) * 2;
    xn = ntab[o]; yn = ntab[o + 1];
  } else {
    xn = (ref.x - c.cx) * c.fx_inv;
    yn = (ref.y - c.cy) * c.fy_inv;
  }
  float lambda;
  if (method == PAGK_PIXEL_AWARE_PREDICTION) {
    const float den = Rcl[6] * xn + Rcl[7] * yn + Rcl[8];
    lambda = (float)(1.0 / (double)den);
  } else {
    lambda = 1.0f;
  }
  pred.x = (M[0] * ref.x + M[1] * ref.y + M[2]) * lambda;
  pred.y = (M[3] * ref.x + M[4] * ref.y + M[5]) * lambda;
  pred_dist = distort_point(c, pred);
  flow.x = pred.x - ref.x;
  flow.y = pred.y - ref.y;
}

struct Work {  // per pair result vectors (always allocated; copied to the caller's non-NULL pointers)
  int N = 0;
  std::vector<P2> pt_predict_un, pt_predict, gyro_un, gyro, flows, pm_un, pm;
  std::vector<uint8_t> status, pm_status;
  std::vector<float> affine, cflows, corners_un, corners, ncc;
  std::vector<double> pix_err, dist;
  std::vector<int32_t> iters;
  void init(int n) {
    N = n;
    P2 z{0.f, 0.f};
    pt_predict_un.assign(n, z); pt_predict.assign(n, z); gyro_un.assign(n, z); gyro.assign(n, z);
    flows.assign(n, z); pm_un.assign(n, z); pm.assign(n, z);
    status.assign(n, 0); pm_status.assign(n, 0);
    affine.assign((size_t)n * 4, 0.f); cflows.assign((size_t)n * 8, 0.f);
    corners_un.assign((size_t)n * 8, 0.f); corners.assign((size_t)n * 8, 0.f);
    ncc.assign(n, 0.f); pix_err.assign(n, 0.0); dist.assign(n, 0.0); iters.assign(n, 0);
  }
};

template <class F>
void parallel_for(int n, int n_threads, F f) {  // static split, mirrors cv::parallel_for_
  if (n_threads <= 1 || n < 2 * n_threads) { f(0, n); return; }
  std::vector<std::thread> th;
  const int chunk = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const int a = t * chunk, b = std::min(n, a + chunk);
    if (a >= b) break;
    th.emplace_back([=] { f(a, b); });
  }
  for (auto &t : th) t.join();
}

// GyroPredictFeatures, src/gyro_aided_tracker.cpp:118-185
int gyro_predict_features(const Cam &c, const float Rcl[9], const float M[9], int method, const float *ntab,
                          const float *keys_un, int half, Work &w, int n_threads) {
  const float hf = (float)half;
  const float corner[4][2] = {{-hf, -hf}, {hf, -hf}, {-hf, hf}, {hf, hf}};  // mvPatchCorners :72-76
  // (B*B^T).inv(): B*B^T = diag(4h^2) exactly; inverse via the 2x2 double adjugate.
  float BBt[4], BBinv[4];
  {
    float Bm[8];
    for (int j = 0; j < 4; ++j) { Bm[j] = corner[j][0]; Bm[4 + j] = corner[j][1]; }
    small_gemm_dbl(Bm, 4, 1, Bm, 1, 4, BBt, 2, 2, 4);  // B * B^T
    small_inv2(BBt, BBinv);
  }
  parallel_for(w.N, n_threads, [&](int i0, int i1) {
    for (int i = i0; i < i1; ++i) {
      P2 ref{keys_un[2 * i], keys_un[2 * i + 1]};
      P2 pun, pd, fl;
      gyro_predict_one(c, Rcl, M, method, ntab, ref, pun, pd, fl);
      if (pun.x < 0 || pun.x >= c.width || pun.y < 0 || pun.y >= c.height) continue;
      if (pd.x < 0 || pd.x >= c.width || pd.y < 0 || pd.y >= c.height) continue;
      w.pt_predict_un[i] = pun;
      w.pt_predict[i] = pd;
      w.status[i] = 1;
      w.flows[i] = fl;
      float C[8];  // matC 2x4: row 0 = x, row 1 = y
      for (int j = 0; j < 4; ++j) {
        P2 cr{ref.x + corner[j][0], ref.y + corner[j][1]};
        P2 cun, cd, cf;
        gyro_predict_one(c, Rcl, M, method, ntab, cr, cun, cd, cf);
        w.corners_un[(size_t)i * 8 + 2 * j] = cun.x; w.corners_un[(size_t)i * 8 + 2 * j + 1] = cun.y;
        w.corners[(size_t)i * 8 + 2 * j] = cd.x; w.corners[(size_t)i * 8 + 2 * j + 1] = cd.y;
        const float vx = cun.x - pun.x, vy = cun.y - pun.y;
        w.cflows[(size_t)i * 8 + 2 * j] = vx; w.cflows[(size_t)i * 8 + 2 * j + 1] = vy;
        C[j] = vx; C[4 + j] = vy;
      }
      // A = matC * B^T * (B*B^T)^-1  (:166-167): C*B^T is a GEMM_2_T gemm (double accumulate),
      // the product with the inverse a plain 2x2 small gemm (float).
      float Bm[8], S[4];
      for (int j = 0; j < 4; ++j) { Bm[j] = corner[j][0]; Bm[4 + j] = corner[j][1]; }
      small_gemm_dbl(C, 4, 1, Bm, 1, 4, S, 2, 2, 4);
      small_gemm_nn(S, BBinv, &w.affine[(size_t)i * 4], 2, 2, 2);
    }
  });
  int n_predict = 0;
  for (int i = 0; i < w.N; ++i) n_predict += w.status[i] ? 1 : 0;
  w.gyro = w.pt_predict;
  w.gyro_un = w.pt_predict_un;
  return n_predict;
}

// ---------------------------------------------------------------------------------------------
// Eigen::Matrix4d H; H.llt().solve(b)   (src/patch_match.cpp:319), Eigen 3.3.4 semantics:
//   llt_inplace<double,Lower>::unblocked (size < 32), then matrixL().solveInPlace, matrixU().solveInPlace
//   through triangular_solver_unroller (fixed size 4) with `.sum()` of fixed-size segments:
//   forward rows of a column-major matrix are strided -> scalar tree  a0 + (a1 + a2);
//   backward rows of the transposed view are contiguous -> SSE2 packets (a0 + a1) + a2.
// m[r][c]; only the lower triangle is read.  The factorisation stops at the first pivot <= 0 and
// leaves the rest untouched; solve() still runs on the partial factor, exactly as Eigen does.
// ---------------------------------------------------------------------------------------------
// Sensitivity study only (tools/eigen_variants.py): Eigen is not in the container, so the operation
// order above is a restatement of Eigen 3.3.4 from memory.  g_llt_variant != 0 switches ONE association to the other
// plausible reading so that the effect of a misreading can be measured; 0 is the order every parity claim refers to.
//   1  pivot:    (m_kk - m_k0^2) - m_k1^2 ...  instead of  m_kk - (m_k0^2 + m_k1^2 + ...)
//   2  column update:  m_ik - (m_i0 m_k0 + m_i1 m_k1 ...)  (dot product first) instead of column by column
//   3  forward substitution sums left to right instead of the scalar tree a0 + (a1 + a2)
//   4  backward substitution sums as a0 + (a1 + a2) instead of the packet order (a0 + a1) + a2
//   5  norm(): ((u0^2 + u1^2) + u2^2) + u3^2 instead of the packet order (u0^2 + u2^2) + (u1^2 + u3^2)
//   6  all of 1..5 together
static int g_llt_variant = 0;

void llt_solve4(const double Hin[4][4], const double bin[4], double x[4]) {
  const int V = g_llt_variant;
  const bool v1 = V == 1 || V == 6, v2 = V == 2 || V == 6, v3 = V == 3 || V == 6, v4 = V == 4 || V == 6;
  double m[4][4];
  memcpy(m, Hin, sizeof(m));
  for (int k = 0; k < 4; ++k) {
    const int rs = 3 - k;
    double piv = m[k][k];
    if (k > 0) {
      if (v1) {
        for (int j = 0; j < k; ++j) piv -= m[k][j] * m[k][j];
      } else {
        double s = m[k][0] * m[k][0];
        for (int j = 1; j < k; ++j) s = s + m[k][j] * m[k][j];
        piv -= s;
      }
    }
    if (piv <= 0.0) break;
    piv = std::sqrt(piv);
    m[k][k] = piv;
    if (k > 0 && rs > 0) {
      if (v2) {
        for (int i = k + 1; i < 4; ++i) {
          double d = m[i][0] * m[k][0];
          for (int j = 1; j < k; ++j) d = d + m[i][j] * m[k][j];
          m[i][k] -= d;
        }
      } else {
        for (int j = 0; j < k; ++j) {
          const double t = -1.0 * m[k][j];
          for (int i = k + 1; i < 4; ++i) m[i][k] += m[i][j] * t;
        }
      }
    }
    for (int i = k + 1; i < 4; ++i) m[i][k] /= piv;
  }
  double r[4] = {bin[0], bin[1], bin[2], bin[3]};
  // L y = b
  r[0] /= m[0][0];
  r[1] -= m[1][0] * r[0];
  r[1] /= m[1][1];
  r[2] -= (m[2][0] * r[0] + m[2][1] * r[1]);
  r[2] /= m[2][2];
  if (v3) r[3] -= ((m[3][0] * r[0] + m[3][1] * r[1]) + m[3][2] * r[2]);
  else r[3] -= (m[3][0] * r[0] + (m[3][1] * r[1] + m[3][2] * r[2]));
  r[3] /= m[3][3];
  // L^T x = y
  r[3] /= m[3][3];
  r[2] -= m[3][2] * r[3];
  r[2] /= m[2][2];
  r[1] -= (m[2][1] * r[2] + m[3][1] * r[3]);
  r[1] /= m[1][1];
  if (v4) r[0] -= (m[1][0] * r[1] + (m[2][0] * r[2] + m[3][0] * r[3]));
  else r[0] -= ((m[1][0] * r[1] + m[2][0] * r[2]) + m[3][0] * r[3]);
  r[0] /= m[0][0];
  x[0] = r[0]; x[1] = r[1]; x[2] = r[2]; x[3] = r[3];
}

// PatchMatch::NCC, src/patch_match.cpp:433-469
float ncc_patch(int half, const Level &ref, const Level &cur, P2 pr, P2 pc, const float *A) {
  float mean_ref = 0.f, mean_cur = 0.f;
  std::vector<float> vr, vc;
  for (int x = -half; x <= half; ++x)
    for (int y = -half; y <= half; ++y) {
      const float a = get_pixel_value(ref, pr.x + x, pr.y + y);
      mean_ref += a;
      vr.push_back(a);
      float b;
      if (!A) {
        b = get_pixel_value(cur, pc.x + x, pc.y + y);
      } else {
        const float wx = A[0] * x + A[1] * y;
        const float wy = A[2] * x + A[3] * y;
        b = get_pixel_value(cur, pc.x + wx, pc.y + wy);
      }
      mean_cur += b;
      vc.push_back(b);
    }
  mean_ref /= vr.size();
  mean_cur /= vc.size();
  float num = 0, d1 = 0, d2 = 0;
  for (size_t i = 0; i < vr.size(); ++i) {
    num += ((vr[i] - mean_ref) * (vc[i] - mean_cur));
    d1 += (vr[i] - mean_ref) * (vr[i] - mean_ref);
    d2 += (vc[i] - mean_cur) * (vc[i] - mean_cur);
  }
  return (float)(num / std::sqrt(d1 * d2 + 1e-10));
}

// ---------------------------------------------------------------------------------------------
// PatchMatch, src/patch_match.cpp
// ---------------------------------------------------------------------------------------------
struct PM {
  int N, half, iterations, pyramids;
  bool gyro_init, inverse, illum, affine, regular, calc_ncc;
  float lambda, alpha, inv_log_max_dist;
  int max_distance;
  double win_size_inv;
  std::vector<Level> pyr1, pyr2;
  std::vector<float> scales;
  const float *keys_un;
  const uint8_t *gyro_status;
  const float *A;  // [N][4]
  std::vector<P2> pt1, pt2;
  std::vector<uint8_t> success;
  std::vector<double> pix_err;
  std::vector<float> ncc;
  std::vector<int32_t> iters;
};

// OpticalFlowConsideringIlluminationChange_onePixel, src/patch_match.cpp:167-367 (forward mode)
void one_pixel(PM &pm, int level, int i) {
  if (!pm.gyro_status[i]) return;
  const int h = pm.half;
  const Level &I1 = pm.pyr1[level], &I2 = pm.pyr2[level];
  P2 pt{pm.pt1[i].x * pm.scales[level], pm.pt1[i].y * pm.scales[level]};
  P2 next;
  if (level == pm.pyramids - 1) {
    next.x = pm.pt2[i].x * pm.scales[level];
    next.y = pm.pt2[i].y * pm.scales[level];
  } else {
    next.x = (float)((double)(pm.pt2[i].x * 1.0f) / 0.5);  // Point2f * 1.0f / (double)mPyramidScale
    next.y = (float)((double)(pm.pt2[i].y * 1.0f) / 0.5);
  }
  float dx = next.x - pt.x, dy = next.y - pt.y;
  float dg = 0.f, db = 0.f;
  float cost = 0.f, lastCost = 0.f;
  bool succ = true;
  const int P = 2 * h + 1, NP = P * P;
  std::vector<float> wxs(NP), wys(NP), vE(NP);
  std::vector<double> vJ((size_t)NP * 4);
  const float *A = pm.A + (size_t)i * 4;
  {
    int idx = 0;
    for (int y = -h; y <= h; ++y)
      for (int x = -h; x <= h; ++x, ++idx) {
        float wx = (float)x, wy = (float)y;
        if (pm.affine) {
          wx = A[0] * x + A[1] * y;
          wy = A[2] * x + A[3] * y;
        }
        wxs[idx] = wx; wys[idx] = wy;
      }
  }
  for (int iter = 0; iter < pm.iterations; ++iter) {
    double H[4][4] = {{0}}, b[4] = {0, 0, 0, 0};
    pm.iters[i] += 1;
    int idx = 0;
    for (int y = -h; y <= h; ++y)
      for (int x = -h; x <= h; ++x, ++idx) {
        const float wx = wxs[idx], wy = wys[idx];
        const float error = get_pixel_value(I2, pt.x + dx + wx, pt.y + dy + wy) + db -
                            (1.0f + dg) * get_pixel_value(I1, pt.x + x, pt.y + y);
        const float Ix = (float)(0.5 * (get_pixel_value(I2, pt.x + dx + wx + 1, pt.y + dy + wy) -
                                        get_pixel_value(I2, pt.x + dx + wx - 1, pt.y + dy + wy)));
        const float Iy = (float)(0.5 * (get_pixel_value(I2, pt.x + dx + wx, pt.y + dy + wy + 1) -
                                        get_pixel_value(I2, pt.x + dx + wx, pt.y + dy + wy - 1)));
        const float de_dg = -get_pixel_value(I1, pt.x, pt.y);
        vJ[(size_t)idx * 4 + 0] = Ix; vJ[(size_t)idx * 4 + 1] = Iy;
        vJ[(size_t)idx * 4 + 2] = de_dg; vJ[(size_t)idx * 4 + 3] = 1;
        vE[idx] = error;
      }
    cost = 0;
    for (int p = 0; p < NP; ++p) {
      const double *J = &vJ[(size_t)p * 4];
      const double e = (double)vE[p];
      for (int r = 0; r < 4; ++r) b[r] += (-J[r]) * e;
      cost += vE[p] * vE[p];
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) H[r][c] += J[r] * J[c];
    }
    if (pm.regular) {
      const double d = std::sqrt(dx * dx + dy * dy);  // float expression, sqrtf, widened
      const float li = pm.lambda * pm.inv_log_max_dist;
      const double e_pen = li * std::log(pm.alpha * d + 1);
      const double jx = li * pm.alpha / (pm.alpha * d + 1) * (dx / d);
      const double jy = li * pm.alpha / (pm.alpha * d + 1) * (dy / d);
      const double JP[4] = {jx, jy, 0, 0};
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) H[r][c] += JP[r] * JP[c];
      for (int r = 0; r < 4; ++r) b[r] += JP[r] * e_pen;
      cost += e_pen * e_pen;
    }
    double up[4];
    llt_solve4(H, b, up);
    if (std::isnan(up[0])) { succ = false; break; }
    if (iter > 0 && cost > lastCost) break;
    dx += up[0];
    dy += up[1];
    if (pm.illum) { dg += up[2]; db += up[3]; }
    lastCost = cost;
    succ = true;
    const double nrm = (g_llt_variant == 5 || g_llt_variant == 6)
                           ? std::sqrt(((up[0] * up[0] + up[1] * up[1]) + up[2] * up[2]) + up[3] * up[3])
                           : std::sqrt((up[0] * up[0] + up[2] * up[2]) + (up[1] * up[1] + up[3] * up[3]));
    if (nrm < 1e-2) break;
  }
  pm.pt2[i].x = pt.x + dx;
  pm.pt2[i].y = pt.y + dy;
  if (level == 0) {
    pm.success[i] = succ;
    pm.pix_err[i] = std::sqrt(lastCost * pm.win_size_inv);
  }
  if (pm.calc_ncc) {
    pm.ncc[i] = ncc_patch(h, pm.pyr1[0], pm.pyr2[0], pm.pt1[i], pm.pt2[i], pm.affine ? A : nullptr);
  } else {
    pm.ncc[i] = 1;
  }
}

int patch_match_run(const pagk_patch_match_in &in, Work &w, int n_threads) {
  PM pm;
  pm.N = in.n_keys;
  pm.half = in.half_patch; pm.iterations = in.iterations; pm.pyramids = in.pyramids;
  pm.gyro_init = in.has_gyro_predict_initial; pm.inverse = in.inverse; pm.illum = in.consider_illumination;
  pm.affine = in.consider_affine_deformation; pm.regular = in.regularization_penalty; pm.calc_ncc = in.calc_ncc;
  pm.lambda = in.lambda; pm.alpha = in.alpha; pm.max_distance = in.max_distance;
  pm.inv_log_max_dist = (float)(1.0 / (double)std::log(pm.alpha * pm.max_distance + 1));
  pm.win_size_inv = (double)(1.0f / (2.0f * pm.half + 1.0f) / (2.0f * pm.half + 1.0f));
  if (pm.inverse) return PAGK_ERR_UNSUPPORTED;
  build_pyramid(in.img_ref, in.width, in.height, in.pitch, pm.pyramids, pm.pyr1);
  build_pyramid(in.img_cur, in.width, in.height, in.pitch, pm.pyramids, pm.pyr2);
  pm.scales.resize(pm.pyramids);
  pm.scales[0] = 1.0f;
  for (int l = 1; l < pm.pyramids; ++l) pm.scales[l] = (float)(pm.scales[l - 1] * 0.5);
  pm.keys_un = in.keys_ref_un; pm.gyro_status = in.status; pm.A = in.affine;
  pm.pt1.resize(pm.N); pm.pt2.resize(pm.N);
  for (int i = 0; i < pm.N; ++i) {
    pm.pt1[i] = P2{in.keys_ref_un[2 * i], in.keys_ref_un[2 * i + 1]};
    pm.pt2[i] = pm.gyro_init ? P2{in.pt_predict_un[2 * i], in.pt_predict_un[2 * i + 1]} : pm.pt1[i];
  }
  pm.success.assign(pm.N, 0); pm.pix_err.assign(pm.N, 0.0); pm.ncc.assign(pm.N, 0.f); pm.iters.assign(pm.N, 0);
  // developer aid: PAGK_ORACLE_LEVEL_ITERS=<file> appends the running pass counts after every level ([level][N] int32 per pair)
  const char *dump_path = std::getenv("PAGK_ORACLE_LEVEL_ITERS");
  for (int level = pm.pyramids - 1; level >= 0; --level) {
    parallel_for(pm.N, n_threads, [&](int i0, int i1) {
      for (int i = i0; i < i1; ++i) one_pixel(pm, level, i);
    });
    if (dump_path) {
      if (FILE *f = std::fopen(dump_path, "ab")) { std::fwrite(pm.iters.data(), sizeof(int32_t), (size_t)pm.N, f); std::fclose(f); }
    }
  }
  // DistortPoints (:409-416) + SetMatcher (:370-388)
  const Cam c = make_cam(in.K, in.dist, in.n_dist, in.width, in.height);
  for (int i = 0; i < pm.N; ++i) {
    w.pm_un[i] = pm.pt2[i];
    w.pm[i] = (in.dist[0] == 0.0f) ? pm.pt2[i] : distort_point(c, pm.pt2[i]);
    w.pm_status[i] = pm.success[i];
    w.pix_err[i] = pm.pix_err[i];
    const float ddx = in.pt_predict_un[2 * i] - pm.pt2[i].x, ddy = in.pt_predict_un[2 * i + 1] - pm.pt2[i].y;
    w.dist[i] = (double)std::sqrt(ddx * ddx + ddy * ddy);
    w.ncc[i] = pm.ncc[i];
    w.iters[i] = pm.iters[i];
  }
  return PAGK_OK;
}

void export_work(const Work &w, pagk_pair_out *o) {
  const size_t n = (size_t)w.N;
  auto cp = [](void *dst, const void *src, size_t bytes) { if (dst && bytes) memcpy(dst, src, bytes); };
  cp(o->pt_predict_un, w.pt_predict_un.data(), n * 8); cp(o->pt_predict, w.pt_predict.data(), n * 8);
  cp(o->status, w.status.data(), n);
  cp(o->pt_gyro_predict_un, w.gyro_un.data(), n * 8); cp(o->pt_gyro_predict, w.gyro.data(), n * 8);
  cp(o->flows_predict_un, w.flows.data(), n * 8);
  cp(o->affine, w.affine.data(), n * 16); cp(o->corner_flows, w.cflows.data(), n * 32);
  cp(o->pt_corners_un, w.corners_un.data(), n * 32); cp(o->pt_corners, w.corners.data(), n * 32);
  cp(o->pm_pt_un, w.pm_un.data(), n * 8); cp(o->pm_pt, w.pm.data(), n * 8);
  cp(o->pm_status, w.pm_status.data(), n);
  cp(o->pixel_error, w.pix_err.data(), n * 8); cp(o->distance, w.dist.data(), n * 8);
  cp(o->ncc, w.ncc.data(), n * 4); cp(o->iters, w.iters.data(), n * 4);
  int64_t tot = 0;
  for (size_t i = 0; i < n; ++i) tot += w.iters[i];
  o->n_iterations = tot;
}

int mode_flags(int e_type, bool &gyro_init, bool &illum, bool &affine, bool &regular) {
  switch (e_type) {  // src/gyro_aided_tracker.cpp:384-414
    case PAGK_IMAGE_ONLY_OPTICAL_FLOW_CONSIDER_ILLUMINATION: gyro_init = false; illum = true; affine = true; regular = false; return 0;
    case PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED: gyro_init = true; illum = false; affine = false; regular = false; return 0;
    case PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION: gyro_init = true; illum = true; affine = false; regular = false; return 0;
    case PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION: gyro_init = true; illum = true; affine = true; regular = false; return 0;
    case PAGK_GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION_REGULAR: gyro_init = true; illum = true; affine = true; regular = true; return 0;
    default: return -1;
  }
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// TrackFeatures(), src/gyro_aided_tracker.cpp:344-426
int track_one(const pagk_params &prm, const pagk_pair_in &in, pagk_pair_out *out, int n_threads) {
  const int half = prm.half_patch == 0 ? 5 : prm.half_patch;
  Work w;
  w.init(in.n_keys);
  integrate_gyro(in, out->Rcl, out->KRKinv);
  const Cam c = make_cam(in.K, in.dist, in.n_dist, in.width, in.height);
  out->t_gyro_predict = out->t_opt_flow = out->t_filter = 0.f;
  out->n_iterations = 0;
  if (prm.e_type == PAGK_GYRO_PREDICT) {
    const double t0 = now_s();
    out->n_predict = gyro_predict_features(c, out->Rcl, out->KRKinv, prm.predict_method, in.normalize_table,
                                           in.keys_ref_un, half, w, n_threads);
    out->t_gyro_predict = (float)(now_s() - t0);
    export_work(w, out);
    return PAGK_OK;
  }
  bool gyro_init, illum, affine, regular;
  if (mode_flags(prm.e_type, gyro_init, illum, affine, regular) != 0) {
    out->n_predict = -1;
    return PAGK_ERR_UNSUPPORTED;
  }
  if (!gyro_init && !in.keys_ref) { out->n_predict = -1; return PAGK_ERR_INVALID; }
  // Step 1 (:261-271)
  double t0 = now_s();
  if (gyro_init) {
    gyro_predict_features(c, out->Rcl, out->KRKinv, prm.predict_method, in.normalize_table, in.keys_ref_un, half, w,
                          n_threads);
  } else {
    for (int i = 0; i < w.N; ++i) {
      w.pt_predict_un[i] = P2{in.keys_ref_un[2 * i], in.keys_ref_un[2 * i + 1]};
      w.pt_predict[i] = P2{in.keys_ref[2 * i], in.keys_ref[2 * i + 1]};
      w.status[i] = 1;
      w.flows[i] = P2{0.f, 0.f};
      w.affine[(size_t)i * 4 + 0] = 1.f; w.affine[(size_t)i * 4 + 1] = 0.f;
      w.affine[(size_t)i * 4 + 2] = 0.f; w.affine[(size_t)i * 4 + 3] = 1.f;
    }
  }
  out->t_gyro_predict = (float)(now_s() - t0);
  // Step 2 (:275-286)
  t0 = now_s();
  pagk_patch_match_in pin;
  memset(&pin, 0, sizeof(pin));
  pin.img_ref = in.img_ref; pin.img_cur = in.img_cur;
  pin.width = in.width; pin.height = in.height; pin.pitch = in.pitch;
  pin.n_keys = in.n_keys; pin.keys_ref_un = in.keys_ref_un;
  pin.status = w.status.data(); pin.affine = w.affine.data();
  memcpy(pin.K, in.K, sizeof(pin.K)); memcpy(pin.dist, in.dist, sizeof(pin.dist)); pin.n_dist = in.n_dist;
  pin.half_patch = half; pin.iterations = prm.iterations; pin.pyramids = prm.pyramids;
  pin.has_gyro_predict_initial = gyro_init; pin.inverse = prm.inverse; pin.consider_illumination = illum;
  pin.consider_affine_deformation = affine; pin.regularization_penalty = regular; pin.calc_ncc = prm.calc_ncc;
  pin.lambda = prm.lambda; pin.alpha = prm.alpha; pin.max_distance = prm.max_distance;
  std::vector<P2> pred_un_snapshot = w.pt_predict_un;
  pin.pt_predict_un = reinterpret_cast<const float *>(pred_un_snapshot.data());
  const int rc = patch_match_run(pin, w, n_threads);
  if (rc != PAGK_OK) { out->n_predict = -1; return rc; }
  out->t_opt_flow = (float)(now_s() - t0);
  // Step 3 (:289-336)
  t0 = now_s();
  double sum = 0;
  int cnt = 0;
  for (int i = 0; i < w.N; ++i)
    if (w.pm_status[i]) { sum += w.pix_err[i]; cnt++; }
  const double avg = sum / cnt;
  const double thPix = 4.0 * avg > half ? 4.0 * avg : half;
  const double thDist = half * 4.0;
  int n_predict = 0;
  for (int i = 0; i < w.N; ++i) {
    if (w.pm_status[i] && w.pix_err[i] < thPix && w.dist[i] < thDist) {
      w.pt_predict[i] = w.pm[i];
      w.pt_predict_un[i] = w.pm_un[i];
      w.status[i] = 1;
      n_predict++;
    } else {
      w.status[i] = 0;
    }
  }
  out->t_filter = (float)(now_s() - t0);
  out->n_predict = n_predict;
  export_work(w, out);
  return PAGK_OK;
}

// ---------------------------------------------------------------------------------------------
// GeometryValidation without the two RANSAC estimators, src/gyro_aided_tracker.cpp:429-508, 589-768.
// cv::Mat H12 = H21.inv() on a 3x3 CV_64F is OpenCV's closed form: adjugate times 1/det in double.
// ---------------------------------------------------------------------------------------------
void inv3_f64(const double *S, double *D) {
  const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (det == 0.) { for (int i = 0; i < 9; ++i) D[i] = 0.; return; }
  const double d = 1. / det;
  double t[9];
  t[0] = (S[4] * S[8] - S[5] * S[7]) * d; t[1] = (S[2] * S[7] - S[1] * S[8]) * d; t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
  t[3] = (S[5] * S[6] - S[3] * S[8]) * d; t[4] = (S[0] * S[8] - S[2] * S[6]) * d; t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
  t[6] = (S[3] * S[7] - S[4] * S[6]) * d; t[7] = (S[1] * S[6] - S[0] * S[7]) * d; t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  memcpy(D, t, sizeof(t));
}

// CheckHomography, :589-678
float check_homography(const double *H21, const std::vector<P2> &p1, const std::vector<P2> &p2, float sigma, std::vector<uint8_t> &inl) {
  double H12[9];
  inv3_f64(H21, H12);
  const double h11 = H21[0], h12 = H21[1], h13 = H21[2], h21 = H21[3], h22 = H21[4], h23 = H21[5], h31 = H21[6], h32 = H21[7], h33 = H21[8];
  const double h11inv = H12[0], h12inv = H12[1], h13inv = H12[2], h21inv = H12[3], h22inv = H12[4], h23inv = H12[5],
               h31inv = H12[6], h32inv = H12[7], h33inv = H12[8];
  const size_t N = p1.size();
  inl.assign(N, 0);
  float score = 0;
  const float th = 5.99;
  const float invSigmaSquare = 1.0 / (sigma * sigma);
  for (size_t i = 0; i < N; i++) {
    bool bIn = true;
    const float u1 = p1[i].x, v1 = p1[i].y, u2 = p2[i].x, v2 = p2[i].y;
    const float w1in2inv = 1.0 / (h31 * u1 + h32 * v1 + h33);
    const float u1in2 = (h11 * u1 + h12 * v1 + h13) * w1in2inv;
    const float v1in2 = (h21 * u1 + h22 * v1 + h23) * w1in2inv;
    const float squareDist2 = (u2 - u1in2) * (u2 - u1in2) + (v2 - v1in2) * (v2 - v1in2);
    const float chiSquare2 = squareDist2 * invSigmaSquare;
    if (chiSquare2 > th) bIn = false; else score += th - chiSquare2;
    const float w2in1inv = 1.0 / (h31inv * u2 + h32inv * v2 + h33inv);
    const float u2in1 = (h11inv * u2 + h12inv * v2 + h13inv) * w2in1inv;
    const float v2in1 = (h21inv * u2 + h22inv * v2 + h23inv) * w2in1inv;
    const float squareDist1 = (u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1);
    const float chiSquare1 = squareDist1 * invSigmaSquare;
    if (chiSquare1 > th) bIn = false; else score += th - chiSquare1;
    inl[i] = bIn ? 1 : 0;
  }
  return score;
}

// CheckFundamental, :680-768
float check_fundamental(const double *F21, const std::vector<P2> &p1, const std::vector<P2> &p2, float sigma, std::vector<uint8_t> &inl) {
  const double f11 = F21[0], f12 = F21[1], f13 = F21[2], f21 = F21[3], f22 = F21[4], f23 = F21[5], f31 = F21[6], f32 = F21[7], f33 = F21[8];
  const size_t N = p1.size();
  inl.assign(N, 0);
  float score = 0;
  const float th = 3.84;
  const float thScore = 5.99;
  const float invSigmaSquare = 1.0 / (sigma * sigma);
  for (size_t i = 0; i < N; i++) {
    bool bIn = true;
    const float u1 = p1[i].x, v1 = p1[i].y, u2 = p2[i].x, v2 = p2[i].y;
    const float a2 = f11 * u1 + f12 * v1 + f13;
    const float b2 = f21 * u1 + f22 * v1 + f23;
    const float c2 = f31 * u1 + f32 * v1 + f33;
    const float num2 = a2 * u2 + b2 * v2 + c2;
    const float squareDist2 = num2 * num2 / (a2 * a2 + b2 * b2);
    const float chiSquare2 = squareDist2 * invSigmaSquare;
    if (chiSquare2 > th) bIn = false; else score += thScore - chiSquare2;
    const float a1 = u2 * f11 + v2 * f21 + f31;
    const float b1 = u2 * f12 + v2 * f22 + f32;
    const float c1 = u2 * f13 + v2 * f23 + f33;
    const float num1 = a1 * u1 + b1 * v1 + c1;
    const float squareDist1 = num1 * num1 / (a1 * a1 + b1 * b1);
    const float chiSquare1 = squareDist1 * invSigmaSquare;
    if (chiSquare1 > th) bIn = false; else score += thScore - chiSquare1;
    inl[i] = bIn ? 1 : 0;
  }
  return score;
}

// GeometryValidation, :429-508
void geometry_validation(const pagk_geometry_in &in, pagk_geometry_out *out) {
  std::vector<P2> p1, p2;
  std::vector<int> idx;
  std::vector<uint8_t> st(in.status, in.status + in.n_keys);
  for (int i = 0; i < in.n_keys; ++i)
    if (st[i]) {
      idx.push_back(i);
      p1.push_back(P2{in.keys_ref_un[2 * i], in.keys_ref_un[2 * i + 1]});
      p2.push_back(P2{in.pt_predict_un[2 * i], in.pt_predict_un[2 * i + 1]});
    }
  out->score_H = out->score_F = 0.f; out->used_H = 0; out->n_inlier = 0;
  out->n_candidates = (int)p1.size();
  if (p1.size() > 8) {
    std::vector<uint8_t> inH, inF;
    const float sH = check_homography(in.H21, p1, p2, in.sigma, inH);
    const float sF = check_fundamental(in.F21, p1, p2, in.sigma, inF);
    const float RH = sH / (sF + sH);
    const bool useH = RH > 0.45;
    const std::vector<uint8_t> &inl = useH ? inH : inF;
    int cnt = 0;
    for (size_t k = 0; k < idx.size(); ++k) {
      if (!inl[k]) st[idx[k]] = 0; else cnt++;
    }
    out->score_H = sH; out->score_F = sF; out->used_H = useH ? 1 : 0; out->n_inlier = cnt;
  }
  if (out->status) memcpy(out->status, st.data(), (size_t)in.n_keys);
}

// ---------------------------------------------------------------------------------------------
// Frame::SetPredictKeyPointsAndMask, src/frame.cpp:115-153
// ---------------------------------------------------------------------------------------------
void set_predict_keypoints_and_mask(const pagk_carry_in &in, pagk_carry_out *out) {
  const int half_path_size = 7;
  const float mfx_inv = 1.0 / in.fx, mfy_inv = 1.0 / in.fy;  // Frame ctor, src/frame.cpp:71
  if (out->mask) memset(out->mask, 1, (size_t)in.width * in.height);
  int cnt = 0;
  for (int i = 0; i < in.n_keys; ++i) {
    if (!in.status[i]) continue;
    const P2 pt_pred{in.pt_predict[2 * i], in.pt_predict[2 * i + 1]};
    const P2 pt_pred_un{in.pt_predict_un[2 * i], in.pt_predict_un[2 * i + 1]};
    P2 pt_pred_normal;
    pt_pred_normal.x = (pt_pred_un.x - in.cx) * mfx_inv;
    pt_pred_normal.y = (pt_pred_un.y - in.cy) * mfy_inv;
    out->keys[2 * cnt] = pt_pred.x; out->keys[2 * cnt + 1] = pt_pred.y;
    out->keys_un[2 * cnt] = pt_pred_un.x; out->keys_un[2 * cnt + 1] = pt_pred_un.y;
    out->keys_normal[2 * cnt] = pt_pred_normal.x; out->keys_normal[2 * cnt + 1] = pt_pred_normal.y;
    out->index_in_last[cnt] = i;
    const float dnx = pt_pred_normal.x - in.keys_normal_last[2 * i], dny = pt_pred_normal.y - in.keys_normal_last[2 * i + 1];
    const double dt = in.t_cur - in.t_last;
    out->flow_velocity_last[2 * cnt] = (float)(dnx / dt);  // cv::Point2f / double: saturate_cast<float>(a.x / b)
    out->flow_velocity_last[2 * cnt + 1] = (float)(dny / dt);
    cnt++;
    if (out->mask) {
      const int _x = std::min(std::max(0, int(pt_pred_un.x) - half_path_size), in.width - 2 * half_path_size);
      const int _y = std::min(std::max(0, int(pt_pred_un.y) - half_path_size), in.height - 2 * half_path_size);
      for (int r = 0; r < 2 * half_path_size; ++r) memset(out->mask + (size_t)(_y + r) * in.width + _x, 0, 2 * half_path_size);
    }
  }
  out->n_out = cnt;
}

// ---------------------------------------------------------------------------------------------
// cv::FAST, TYPE_9_16 (OpenCV modules/features2d/src/fast.cpp, FAST_t<16> and cornerScore<16>), restated from its
// published algorithm.  PINNED bit-exact against cv2 4.13 (tests/golden/fast.npz): positions, order and responses.
// ---------------------------------------------------------------------------------------------
using pagk_cv::fast_detect;

// ---------------------------------------------------------------------------------------------
// ORBextractor::ComputeKeyPointsOctTree for one level, up to vToDistributeKeys (src/ORBextractor.cc:789-852), plus the
// final mask filter of DetectFeatures (:1200-1203).  The cell arithmetic is the reference's (floats and all).
// ---------------------------------------------------------------------------------------------
int orb_cell_detect(const uint8_t *img, int cols, int rows, int step, int ini_th, int min_th, const uint8_t *mask, int max_out,
                    float *xy, float *response) {
  const int EDGE_THRESHOLD = 19;
  const float W = 30;
  const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
  const int maxBorderX = cols - EDGE_THRESHOLD + 3, maxBorderY = rows - EDGE_THRESHOLD + 3;
  const float width = (maxBorderX - minBorderX), height = (maxBorderY - minBorderY);
  const int nCols = width / W, nRows = height / W;
  if (nCols < 1 || nRows < 1) return 0;
  const int wCell = ceil(width / nCols), hCell = ceil(height / nRows);
  int n = 0;
  std::vector<float> cxy, crs;
  for (int i = 0; i < nRows; i++) {
    const float iniY = minBorderY + i * hCell;
    float maxY = iniY + hCell + 6;
    if (iniY >= maxBorderY - 3) continue;
    if (maxY > maxBorderY) maxY = maxBorderY;
    for (int j = 0; j < nCols; j++) {
      const float iniX = minBorderX + j * wCell;
      float maxX = iniX + wCell + 6;
      if (iniX >= maxBorderX - 6) continue;
      if (maxX > maxBorderX) maxX = maxBorderX;
      const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
      const int cap = std::max(cw * ch, 1);
      cxy.resize((size_t)2 * cap); crs.resize((size_t)cap);
      const uint8_t *view = img + (size_t)y0 * step + x0;
      int k = fast_detect(view, cw, ch, step, ini_th, true, nullptr, cap, cxy.data(), crs.data());
      if (k == 0) k = fast_detect(view, cw, ch, step, min_th, true, nullptr, cap, cxy.data(), crs.data());
      for (int q = 0; q < k; ++q) {
        const float px = cxy[2 * q] + j * wCell + minBorderX, py = cxy[2 * q + 1] + i * hCell + minBorderY;
        if (mask && !mask[(size_t)(int)py * cols + (int)px]) continue;
        if (n < max_out) { xy[2 * n] = px; xy[2 * n + 1] = py; response[n] = crs[q]; }
        ++n;
      }
    }
  }
  return n;
}

// ---------------------------------------------------------------------------------------------
// cv::remap, INTER_LINEAR, CV_8UC1, CV_32FC1 maps, BORDER_CONSTANT 0 (OpenCV modules/imgproc/src/imgwarp.cpp: remap ->
// remapBilinear with the INTER_BITS = 5 coordinate grid and the INTER_REMAP_COEF_BITS = 15 weight table), restated from its
// published algorithm.  PINNED bit-exact against cv2 4.13 (tests/golden/remap.npz).
// ---------------------------------------------------------------------------------------------
void remap_linear(const uint8_t *src, int cols, int rows, int step, const float *mx, const float *my, int dcols, int drows, uint8_t *dst) {
  auto px = [&](int y, int x) -> int { return (x >= 0 && x < cols && y >= 0 && y < rows) ? src[(size_t)y * step + x] : 0; };
  auto sat16 = [](int v) { return v < -32768 ? -32768 : v > 32767 ? 32767 : v; };
  for (int y = 0; y < drows; ++y)
    for (int x = 0; x < dcols; ++x) {
      const size_t o = (size_t)y * dcols + x;
      const int sx = (int)lrintf(mx[o] * 32.0f), sy = (int)lrintf(my[o] * 32.0f);  // cvRound(x * INTER_TAB_SIZE)
      const int ix = sat16(sx >> 5), iy = sat16(sy >> 5), fx = sx & 31, fy = sy & 31;
      const int w00 = (32 - fy) * (32 - fx) * 32, w01 = (32 - fy) * fx * 32, w10 = fy * (32 - fx) * 32, w11 = fy * fx * 32;
      const int v = (px(iy, ix) * w00 + px(iy, ix + 1) * w01 + px(iy + 1, ix) * w10 + px(iy + 1, ix + 1) * w11 + (1 << 14)) >> 15;
      dst[o] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
}

}  // namespace

extern "C" {

int pagk_oracle_resize_half(const uint8_t *src, int cols, int rows, int step, uint8_t *dst) {
  resize_half(src, cols, rows, step, dst, (int)(cols * 0.5), (int)(rows * 0.5));
  return PAGK_OK;
}

// level `level` of the pyramid of one image, continuous cols x rows
int pagk_oracle_pyramid_level(const uint8_t *img, int width, int height, int pitch, int level, uint8_t *dst) {
  std::vector<Level> pyr;
  build_pyramid(img, width, height, pitch, level + 1, pyr);
  memcpy(dst, pyr[level].data(), (size_t)pyr[level].cols * pyr[level].rows);
  return PAGK_OK;
}

float pagk_oracle_get_pixel_value(const uint8_t *img, int cols, int rows, int pitch, float x, float y) {
  Level L;
  level_from_image(img, cols, rows, pitch, L);
  return get_pixel_value(L, x, y);
}

// sensitivity study switch (see llt_solve4); not thread-safe against running track calls, 0 restores the oracle proper
void pagk_oracle_set_llt_variant(int v) { g_llt_variant = v; }

void pagk_oracle_llt_solve(const double *H16, const double *b4, double *x4) {
  double H[4][4];
  memcpy(H, H16, sizeof(H));
  llt_solve4(H, b4, x4);
}

// A = matC * B^T * (B*B^T)^-1 for given corner flows C[4][2] (test hook for the cv::Mat expression at
// src/gyro_aided_tracker.cpp:166-167)
void pagk_oracle_affine_from_corners(const float *cflows8, int half, float *A4) {
  const float hf = (float)half;
  const float corner[4][2] = {{-hf, -hf}, {hf, -hf}, {-hf, hf}, {hf, hf}};
  float Bm[8], C[8], BBt[4], BBinv[4], S[4];
  for (int j = 0; j < 4; ++j) { Bm[j] = corner[j][0]; Bm[4 + j] = corner[j][1]; C[j] = cflows8[2 * j]; C[4 + j] = cflows8[2 * j + 1]; }
  small_gemm_dbl(Bm, 4, 1, Bm, 1, 4, BBt, 2, 2, 4);
  small_inv2(BBt, BBinv);
  small_gemm_dbl(C, 4, 1, Bm, 1, 4, S, 2, 2, 4);
  small_gemm_nn(S, BBinv, A4, 2, 2, 2);
}

int pagk_oracle_integrate_gyro(const pagk_pair_in *in, float *Rcl, float *KRKinv) {
  integrate_gyro(*in, Rcl, KRKinv);
  return PAGK_OK;
}

int pagk_oracle_gyro_predict(const pagk_params *prm, const pagk_pair_in *in, pagk_pair_out *out) {
  const int half = prm->half_patch == 0 ? 5 : prm->half_patch;
  Work w;
  w.init(in->n_keys);
  integrate_gyro(*in, out->Rcl, out->KRKinv);
  const Cam c = make_cam(in->K, in->dist, in->n_dist, in->width, in->height);
  out->n_predict = gyro_predict_features(c, out->Rcl, out->KRKinv, prm->predict_method, in->normalize_table,
                                         in->keys_ref_un, half, w, 1);
  export_work(w, out);
  return PAGK_OK;
}

int pagk_oracle_patch_match(const pagk_patch_match_in *in, pagk_pair_out *out, int n_threads) {
  Work w;
  w.init(in->n_keys);
  const int rc = patch_match_run(*in, w, n_threads);
  if (rc != PAGK_OK) return rc;
  export_work(w, out);
  return PAGK_OK;
}

int pagk_oracle_geometry_validation(int n_pairs, const pagk_geometry_in *in, pagk_geometry_out *out) {
  for (int p = 0; p < n_pairs; ++p) geometry_validation(in[p], &out[p]);
  return PAGK_OK;
}

int pagk_oracle_set_predict_keypoints_and_mask(int n_pairs, const pagk_carry_in *in, pagk_carry_out *out) {
  for (int p = 0; p < n_pairs; ++p) set_predict_keypoints_and_mask(in[p], &out[p]);
  return PAGK_OK;
}

int pagk_oracle_fast_detect(const uint8_t *img, int width, int height, int pitch, int threshold, int nonmax, const uint8_t *mask,
                            int max_out, float *xy, float *response, int *n_out) {
  *n_out = fast_detect(img, width, height, pitch, threshold, nonmax != 0, mask, max_out, xy, response);
  return PAGK_OK;
}

int pagk_oracle_orb_cell_detect(const uint8_t *img, int width, int height, int pitch, int ini_th, int min_th, const uint8_t *mask,
                                int max_out, float *xy, float *response, int *n_out) {
  *n_out = orb_cell_detect(img, width, height, pitch, ini_th, min_th, mask, max_out, xy, response);
  return PAGK_OK;
}

int pagk_oracle_remap_linear(const uint8_t *src, int width, int height, int pitch, const float *map_x, const float *map_y,
                             int dst_width, int dst_height, uint8_t *dst) {
  remap_linear(src, width, height, pitch, map_x, map_y, dst_width, dst_height, dst);
  return PAGK_OK;
}

int pagk_oracle_track(const pagk_params *prm, const pagk_pair_in *in, pagk_pair_out *out, int n_threads) {
  return track_one(*prm, *in, out, n_threads);
}

// Pairs run one after another, each with n_threads over its features, like the reference's
// per-frame cv::parallel_for_ (src/patch_match.cpp:103).
int pagk_oracle_track_batch(const pagk_params *prm, int n_pairs, const pagk_pair_in *in, pagk_pair_out *out,
                            int n_threads) {
  int rc = PAGK_OK;
  for (int p = 0; p < n_pairs; ++p) {
    const int r = track_one(*prm, in[p], &out[p], n_threads);
    if (r != PAGK_OK) rc = r;
  }
  return rc;
}

}  // extern "C"
