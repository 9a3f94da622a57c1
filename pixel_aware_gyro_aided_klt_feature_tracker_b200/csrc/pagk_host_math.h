// pagk_host_math.h -- host arithmetic of the product path that is too small for a kernel:
// IntegrateGyroMeasurements / IntegrateOneGyroMeasurement / SetRcl
// (reference src/gyro_aided_tracker.cpp:511-587), about ten 3x3 products per frame pair.
//
// The reference does this with float cv::Mat expressions; OpenCV evaluates them as documented in
// SURVEY.md appendix B.2 (checked against cv2 4.13):
//   * A*B on CV_32F without a transpose flag, inner dimension 2..4: float accumulation, left to right
//   * A.t()*B.t(): the general gemm, double accumulation, one rounding to float
//   * W*s + I  is scaleAdd(W, (float)s, I); (W*W)*s scales the float product by (float)s
//   * M.inv() of a 3x3: adjugate in double times 1/det, rounded to float
// libm sinf/cosf/sqrtf are used exactly as the reference does.  This file is compiled by the host
// compiler with -ffp-contract=off (no FMA), like the reference's SSE2 build.
#pragma once
#include <cmath>
#include <cstring>

namespace pagk_host {

struct Mat3 {
  float v[3][3];
};

inline Mat3 identity3() {
  Mat3 m;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) m.v[r][c] = (r == c) ? 1.0f : 0.0f;
  return m;
}

inline Mat3 from_array(const float *a) {
  Mat3 m;
  std::memcpy(m.v, a, sizeof(m.v));
  return m;
}

// float-accumulating product (OpenCV's small-matrix gemm)
inline Mat3 mul_f32(const Mat3 &a, const Mat3 &b) {
  Mat3 d;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const float p0 = a.v[r][0] * b.v[0][c];
      const float p1 = a.v[r][1] * b.v[1][c];
      const float p2 = a.v[r][2] * b.v[2][c];
      d.v[r][c] = (p0 + p1) + p2;
    }
  return d;
}

// a^T * b^T with double accumulation (OpenCV's general gemm with GEMM_1_T | GEMM_2_T)
inline Mat3 mul_tt_f64(const Mat3 &a, const Mat3 &b) {
  Mat3 d;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      double s = 0.0;
      for (int k = 0; k < 3; ++k) s += (double)a.v[k][r] * (double)b.v[c][k];
      d.v[r][c] = (float)s;
    }
  return d;
}

inline Mat3 inverse3(const Mat3 &m) {
  const double a = m.v[0][0], b = m.v[0][1], c = m.v[0][2];
  const double d = m.v[1][0], e = m.v[1][1], f = m.v[1][2];
  const double g = m.v[2][0], h = m.v[2][1], i = m.v[2][2];
  const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
  Mat3 o;
  if (det == 0.0) {
    std::memset(o.v, 0, sizeof(o.v));
    return o;
  }
  const double s = 1.0 / det;
  o.v[0][0] = (float)((e * i - f * h) * s); o.v[0][1] = (float)((c * h - b * i) * s); o.v[0][2] = (float)((b * f - c * e) * s);
  o.v[1][0] = (float)((f * g - d * i) * s); o.v[1][1] = (float)((a * i - c * g) * s); o.v[1][2] = (float)((c * d - a * f) * s);
  o.v[2][0] = (float)((d * h - e * g) * s); o.v[2][1] = (float)((b * g - a * h) * s); o.v[2][2] = (float)((a * e - b * d) * s);
  return o;
}

// IntegrateOneGyroMeasurement, reference src/gyro_aided_tracker.cpp:564-587
inline Mat3 delta_rotation(const float gyro[3], const float bias[3], double dt) {
  const float x = (float)((double)(gyro[0] - bias[0]) * dt);
  const float y = (float)((double)(gyro[1] - bias[1]) * dt);
  const float z = (float)((double)(gyro[2] - bias[2]) * dt);
  const float d2 = (x * x + y * y) + z * z;
  const float d = sqrtf(d2);
  Mat3 W;
  W.v[0][0] = 0.f; W.v[0][1] = -z; W.v[0][2] = y;
  W.v[1][0] = z; W.v[1][1] = 0.f; W.v[1][2] = -x;
  W.v[2][0] = -y; W.v[2][1] = x; W.v[2][2] = 0.f;
  const Mat3 I = identity3();
  Mat3 R;
  if ((double)d < 1e-4) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) R.v[r][c] = I.v[r][c] + W.v[r][c];
    return R;
  }
  const float ka = (float)((double)sinf(d) * (1.0 / (double)d));
  const float kb = (float)((double)(1.0f - cosf(d)) * (1.0 / (double)d2));
  const Mat3 WW = mul_f32(W, W);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const float first = W.v[r][c] * ka + I.v[r][c];
      const float second = WW.v[r][c] * kb;
      R.v[r][c] = first + second;
    }
  return R;
}

// IntegrateGyroMeasurements, reference src/gyro_aided_tracker.cpp:521-562
inline Mat3 integrate(int n_imu, const double *t, const float *w, double t_ref, double t_cur, const float bias[3],
                      const Mat3 &Rbc) {
  Mat3 dR = identity3();
  const int n = n_imu - 1;
  for (int i = 0; i < n; ++i) {
    const float *wa = w + 3 * i, *wb = w + 3 * (i + 1);
    float av[3] = {0.f, 0.f, 0.f};
    float tstep = 0.f;
    const bool first = (i == 0), last = (i == n - 1);
    if (first && !last) {
      const float tab = (float)(t[i + 1] - t[i]);
      const float tini = (float)(t[i] - t_ref);
      const float q = tini / tab;
      for (int k = 0; k < 3; ++k) av[k] = ((wa[k] + wb[k]) - (wb[k] - wa[k]) * q) * 0.5f;
      tstep = (float)(t[i + 1] - t_ref);
    } else if (!last) {
      for (int k = 0; k < 3; ++k) av[k] = (wa[k] + wb[k]) * 0.5f;
      tstep = (float)(t[i + 1] - t[i]);
    } else if (!first) {
      const float tab = (float)(t[i + 1] - t[i]);
      const float tend = (float)(t[i + 1] - t_cur);
      const float q = tend / tab;
      for (int k = 0; k < 3; ++k) av[k] = ((wa[k] + wb[k]) - (wb[k] - wa[k]) * q) * 0.5f;
      tstep = (float)(t_cur - t[i]);
    } else {
      for (int k = 0; k < 3; ++k) av[k] = wa[k];
      tstep = (float)(t_cur - t_ref);
    }
    dR = mul_f32(dR, delta_rotation(av, bias, (double)tstep));
  }
  return mul_f32(mul_tt_f64(Rbc, dR), Rbc);  // Rbc^T * dR^T * Rbc
}

// SetRcl, reference src/gyro_aided_tracker.cpp:511-519: K * Rcl * K.inv()
inline Mat3 krkinv(const Mat3 &K, const Mat3 &Rcl) { return mul_f32(mul_f32(K, Rcl), inverse3(K)); }

// (B*B^T).inv() diagonal for the patch corners (+-h, +-h): B*B^T = diag(4h^2); 2x2 inverse through
// the double adjugate: (float)(s11 * (1/det)).
inline float bbt_inverse_diag(int half) {
  const float s = 4.0f * (float)half * (float)half;
  const double det = (double)s * (double)s;
  return (float)((double)s * (1.0 / det));
}

}  // namespace pagk_host

// cv::Mat::inv() of a 3x3 CV_64F (GeometryValidation's H12 = H21.inv(), reference src/gyro_aided_tracker.cpp:597):
// OpenCV's closed form, adjugate times 1/det in double; an all-zero result for a singular matrix.
inline void pagk_inv3_f64(const double *S, double *D) {
  const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (det == 0.) { for (int i = 0; i < 9; ++i) D[i] = 0.; return; }
  const double d = 1. / det;
  double t[9];
  t[0] = (S[4] * S[8] - S[5] * S[7]) * d; t[1] = (S[2] * S[7] - S[1] * S[8]) * d; t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
  t[3] = (S[5] * S[6] - S[3] * S[8]) * d; t[4] = (S[0] * S[8] - S[2] * S[6]) * d; t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
  t[6] = (S[3] * S[7] - S[4] * S[6]) * d; t[7] = (S[1] * S[6] - S[0] * S[7]) * d; t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  for (int i = 0; i < 9; ++i) D[i] = t[i];
}
