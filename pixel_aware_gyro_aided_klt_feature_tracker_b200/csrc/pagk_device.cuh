// pagk_device.cuh -- device-side data layout and arithmetic helpers shared by the pagk kernels.
//
// Arithmetic contract: every FP32 expression below is evaluated operator by
// operator, round-to-nearest, with NO fused multiply-add -- the translation unit is compiled with
// --fmad=false -- because the reference is an x86-64 SSE2 build (CMakeLists.txt:10-11) and its
// 4x4 normal matrix is structurally singular, so the result depends on the last bit.  The only FMAs
// are the explicit fma() calls that accumulate H and b in double: those products are exact
// (float x float in double), so DFMA rounds exactly like the reference's separate multiply and add.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define PAGK_MAX_LEVELS 8

// One pyramid level inside an image slot.  The reference's levels are continuous cv::Mats (step == cols), and the
// +1 / +step taps of PatchMatch::GetPixelValue (reference src/patch_match.cpp:399-403) read one column past the row
// end (the first byte of the next row) and one row past the image.  Here a level has a row pitch that is a multiple
// of four bytes (the 4-byte asynchronous copies that stage the alignment kernel's windows need aligned rows):
//   pitch == cols (cols % 4 == 0)   the level is continuous, column `cols` of row y IS pixel (0, y + 1)
//   pitch  > cols                   column `cols` of row y is an explicit WRAP byte holding pixel (0, min(y + 1, rows - 1))
// and in both cases row `rows` is a guard row (a copy of row rows - 1, with its own wrap byte).  So the byte at
// (x, y) for 0 <= x <= cols, 0 <= y <= rows is what a continuous cv::Mat followed by a copy of its last row yields.
// Past that the allocation continues (at least 32 rows and 64 bytes of slack) so that a staged window may overhang.
struct PagkLevelGeom {
  int cols, rows;
  int pitch;              // row pitch in elements, a multiple of 4, >= cols (+1 when it is not cols)
  unsigned int offset;    // bytes from the start of the image slot, 256-byte aligned
};

struct PagkGeom {
  int levels;
  int width, height;
  unsigned long long slot_bytes;  // bytes per image slot (all levels), 256-byte aligned
  PagkLevelGeom lv[PAGK_MAX_LEVELS];
};

// What GyroAidedTracker::Initialize / SetRcl cache per tracker (src/gyro_aided_tracker.cpp:64-70, 511-519)
struct PagkPairConst {
  float M[9];  // mKRKinv
  float r31, r32, r33;
  float fx, fy, cx, cy, fx_inv, fy_inv;
  float k1, k2, p1, p2, k3;
  int n_keys;
  int has_table;  // mNormalizeTable non-empty
};

struct PagkPairResult {
  int n_predict;
  int cnt_pm_ok;
  long long n_iterations;
  double avg_pixel_error;
};

// GeometryValidation (reference src/gyro_aided_tracker.cpp:429-508): the two models of one pair and what comes back
struct PagkGeoModel {
  double H21[9], H12[9], F21[9];
  float sigma;
  int n_keys;
};
struct PagkGeoResult {
  float score_H, score_F;
  int used_H, n_candidates, n_inlier, pad;
};

// Frame::SetPredictKeyPointsAndMask (reference src/frame.cpp:115-153): per pair constants and the survivor count
struct PagkCarryConst {
  float cx, cy, fx_inv, fy_inv;
  double dt;  // mTimeStamp - mpLastFrame->mTimeStamp
  int n_keys, width, height, pad;
};

// Mode of one run (what TrackFeatures derives from eType, :384-414, plus the PatchMatch ctor args)
struct PagkMode {
  int half, iterations, levels;
  int gyro_init, illum, affine, regular, calc_ncc;
  int predict_method;
  float lambda, alpha, inv_log_max_dist;
  double win_size_inv;
  float bb_inv;  // (B*B^T)^-1 diagonal, f32(1/(4 h^2)) through the double adjugate
};

// Device result vectors, structure of arrays, [max_pairs][max_keys]
struct PagkOutPtrs {
  float2 *pt_predict_un, *pt_predict, *pt_gyro_un, *pt_gyro, *flows;
  float4 *affine;
  float2 *cflows, *corners_un, *corners;  // [..][4]
  float2 *pm_un, *pm;
  unsigned char *status, *pm_status, *gyro_status;
  double *pix_err, *dist;
  float *ncc;
  int *iters;
};

// PatchMatch::GetPixelValue, reference src/patch_match.cpp:391-406, on a continuous level.
__device__ __forceinline__ float pagk_sample(const unsigned char *__restrict__ img, int pitch, int cols, int rows, float x,
                                             float y) {
  if (x < 0.f) x = 0.f;
  if (y < 0.f) y = 0.f;
  if (x >= (float)cols) x = (float)(cols - 1);
  if (y >= (float)rows) y = (float)(rows - 1);
  const int ix = (int)x, iy = (int)y;
  const unsigned char *d = img + (size_t)iy * pitch + ix;
  const float xx = x - floorf(x), yy = y - floorf(y);
  const float a = 1.0f - xx, b = 1.0f - yy;
  const float top = a * (float)__ldg(d) + xx * (float)__ldg(d + 1);
  const float bot = a * (float)__ldg(d + pitch) + xx * (float)__ldg(d + pitch + 1);
  return b * top + yy * bot;
}

// radial-tangential distortion, reference src/gyro_aided_tracker.cpp:221-231 and src/utils.cpp:60-73
__device__ __forceinline__ float2 pagk_distort(const PagkPairConst &c, float2 p) {
  const float x = (p.x - c.cx) * c.fx_inv;
  const float y = (p.y - c.cy) * c.fy_inv;
  const float r2 = x * x + y * y;
  const float r4 = r2 * r2;
  const float r6 = r4 * r2;
  const float rad = ((1.0f + c.k1 * r2) + c.k2 * r4) + c.k3 * r6;
  const float xd = (x * rad + ((2.0f * c.p1) * x) * y) + c.p2 * (r2 + (2.0f * x) * x);
  const float yd = (y * rad + c.p1 * (r2 + (2.0f * y) * y)) + ((2.0f * c.p2) * x) * y;
  return make_float2(c.fx * xd + c.cx, c.fy * yd + c.cy);
}

// IEEE double division / square root (kept as named wrappers: out-of-line versions were measured slower)
__device__ __forceinline__ double pagk_ddiv(double a, double b) { return a / b; }
__device__ __forceinline__ double pagk_dsqrt(double a) { return sqrt(a); }
// PatchMatch::GetPixelValue out of line, for the rare per-sample paths of the production kernel
static __device__ __noinline__ float pagk_sample_call(const unsigned char *__restrict__ img, int pitch, int cols, int rows, float x, float y) {
  return pagk_sample(img, pitch, cols, rows, x, y);
}

// Eigen::Matrix4d::llt().solve(b) as Eigen 3.3 evaluates it for a fixed 4x4 (see oracle/pagk_oracle.cpp
// llt_solve4 for the derivation of the operation order).  Lower triangle of the symmetric H is passed
// as h00,h10,h11,h20,h21,h22,h30,h31,h32,h33.  --fmad=false keeps every multiply and add separate.
// Factorisation half: H = L L^T in place on the lower triangle (stops at the first pivot <= 0 like Eigen).
__device__ __forceinline__ void pagk_llt_factor4(double &h00, double &h10, double &h11, double &h20, double &h21,
                                                 double &h22, double &h30, double &h31, double &h32, double &h33) {
  bool go = true;
  {  // k = 0
    double piv = h00;
    if (piv <= 0.0) go = false;
    if (go) {
      piv = pagk_dsqrt(piv);
      h00 = piv;
      h10 = pagk_ddiv(h10, piv); h20 = pagk_ddiv(h20, piv); h30 = pagk_ddiv(h30, piv);
    }
  }
  if (go) {  // k = 1
    double piv = h11 - h10 * h10;
    if (piv <= 0.0) go = false;
    if (go) {
      piv = pagk_dsqrt(piv);
      h11 = piv;
      const double t = -1.0 * h10;
      h21 += h20 * t; h31 += h30 * t;
      h21 = pagk_ddiv(h21, piv); h31 = pagk_ddiv(h31, piv);
    }
  }
  if (go) {  // k = 2
    double piv = h22 - (h20 * h20 + h21 * h21);
    if (piv <= 0.0) go = false;
    if (go) {
      piv = pagk_dsqrt(piv);
      h22 = piv;
      const double t0 = -1.0 * h20, t1 = -1.0 * h21;
      h32 += h30 * t0;
      h32 += h31 * t1;
      h32 = pagk_ddiv(h32, piv);
    }
  }
  if (go) {  // k = 3
    double piv = h33 - ((h30 * h30 + h31 * h31) + h32 * h32);
    if (!(piv <= 0.0)) h33 = pagk_dsqrt(piv);  // a NaN pivot is not <= 0: Eigen goes on with sqrt(NaN)
  }
}

// Substitution half: L y = b, then L^T x = y, in Eigen's fixed-size unroller order.
__device__ __forceinline__ void pagk_llt_subst4(double h00, double h10, double h11, double h20, double h21, double h22,
                                                double h30, double h31, double h32, double h33, double b0, double b1,
                                                double b2, double b3, double &x0, double &x1, double &x2, double &x3) {
  double r0 = pagk_ddiv(b0, h00);
  double r1 = pagk_ddiv(b1 - h10 * r0, h11);
  double r2 = pagk_ddiv(b2 - (h20 * r0 + h21 * r1), h22);
  double r3 = pagk_ddiv(b3 - (h30 * r0 + (h31 * r1 + h32 * r2)), h33);
  r3 = pagk_ddiv(r3, h33);
  r2 = pagk_ddiv(r2 - h32 * r3, h22);
  r1 = pagk_ddiv(r1 - (h21 * r2 + h31 * r3), h11);
  r0 = pagk_ddiv(r0 - ((h10 * r1 + h20 * r2) + h30 * r3), h00);
  x0 = r0; x1 = r1; x2 = r2; x3 = r3;
}

__device__ __forceinline__ void pagk_llt_solve4(double h00, double h10, double h11, double h20, double h21, double h22,
                                                double h30, double h31, double h32, double h33, double b0, double b1,
                                                double b2, double b3, double &x0, double &x1, double &x2, double &x3) {
  pagk_llt_factor4(h00, h10, h11, h20, h21, h22, h30, h31, h32, h33);
  pagk_llt_subst4(h00, h10, h11, h20, h21, h22, h30, h31, h32, h33, b0, b1, b2, b3, x0, x1, x2, x3);
}
