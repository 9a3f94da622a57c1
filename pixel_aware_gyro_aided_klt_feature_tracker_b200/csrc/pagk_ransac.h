// pagk_ransac.h -- the two robust estimators of GyroAidedTracker::GeometryValidation, from scratch:
//   cv::findHomography(vPts1, vPts2, cv::RANSAC, 3)               reference src/gyro_aided_tracker.cpp:597
//   cv::findFundamentalMat(vPts1, vPts2, CV_FM_RANSAC, 3., 0.99)  reference src/gyro_aided_tracker.cpp:691
//
// These are OpenCV library calls whose hypotheses come from OpenCV's own RNG: no implementation outside OpenCV reproduces
// their models bit for bit, so this one is accepted statistically (inlier sets and chi-square scores against cv2 fixtures,
// tests/test_ransac.py) and is kept apart from the bit-exact path.  What it shares with OpenCV is the method: minimal
// samples (4 points, normalised DLT / 8 points, normalised eight-point algorithm with the rank-2 constraint), the
// inlier tests (forward reprojection error / the larger of the two epipolar distances, both against 3 px), a refit on
// all inliers of the best hypothesis, and for the homography a Gauss-Newton polish of the reprojection error where OpenCV
// runs Levenberg-Marquardt.  The hypotheses come from a counter-based generator: (seed, pair, hypothesis, draw) -> index,
// so a run is reproducible and every hypothesis is independent of every other (one thread each).
//
// Everything here is plain double arithmetic in functions that compile for the host and the device (PAGK_HD): the device
// kernel (pagk_kernels.cu) and the host harness of the tests (tests/cpp/ransac_host.cpp) run the same code.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define PAGK_HD __host__ __device__ __forceinline__
#else
#define PAGK_HD inline
#endif

namespace pagk_ransac {

constexpr int kHypotheses = 1024;     // per model and frame pair
constexpr double kThreshold2 = 9.0;   // (3 px)^2, both estimators

// counter-based generator: a 64-bit mix (splitmix64 finaliser) of the four counters
PAGK_HD uint32_t rnd(uint32_t seed, uint32_t pair, uint32_t hyp, uint32_t draw) {
  uint64_t z = ((uint64_t)seed << 32) ^ ((uint64_t)pair * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)hyp << 20) ^ (uint64_t)draw;
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}

// k distinct indices in [0, n), n >= k
template <int K>
PAGK_HD void sample(uint32_t seed, uint32_t pair, uint32_t hyp, int n, int *idx) {
  uint32_t draw = 0;
  for (int j = 0; j < K; ++j) {
    for (int attempt = 0; attempt < 64; ++attempt) {
      const int c = (int)(((uint64_t)rnd(seed, pair, hyp, draw++) * (uint64_t)n) >> 32);
      bool dup = false;
      for (int i = 0; i < j; ++i) dup |= (idx[i] == c);
      idx[j] = c;
      if (!dup) break;
    }
  }
}

// cyclic Jacobi on a symmetric N x N matrix (row-major, destroyed); V gets the eigenvectors as columns.  Returns the
// index of the smallest eigenvalue (the eigenvalues end on the diagonal of A).
template <int N>
PAGK_HD int jacobi_smallest(double *A, double *V) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i * N + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < N; ++i) {
      diag += A[i * N + i] * A[i * N + i];
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    }
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {  // columns p and q of A
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {  // rows p and q of A
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < N; ++i)
    if (A[i * N + i] < A[best * N + best]) best = i;
  return best;
}

// Hartley normalisation of a point set given as (sum x, sum y, sum of distances to the centroid needs a second pass):
// the callers compute centroid and mean distance themselves and pass the similarity (s, cx, cy): x' = s (x - cx)
struct Sim { double s, cx, cy; };

PAGK_HD void mat3_mul(const double *a, const double *b, double *c) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

// the two DLT rows of one correspondence (x, y) -> (u, v), both already normalised
PAGK_HD void h_rows(double x, double y, double u, double v, double *r0, double *r1) {
  r0[0] = -x; r0[1] = -y; r0[2] = -1.0; r0[3] = 0.0; r0[4] = 0.0; r0[5] = 0.0; r0[6] = u * x; r0[7] = u * y; r0[8] = u;
  r1[0] = 0.0; r1[1] = 0.0; r1[2] = 0.0; r1[3] = -x; r1[4] = -y; r1[5] = -1.0; r1[6] = v * x; r1[7] = v * y; r1[8] = v;
}
// the eight-point row of one correspondence: x2^T F x1 = 0
PAGK_HD void f_row(double x, double y, double u, double v, double *r) {
  r[0] = u * x; r[1] = u * y; r[2] = u; r[3] = v * x; r[4] = v * y; r[5] = v; r[6] = x; r[7] = y; r[8] = 1.0;
}

// H (normalised coordinates, from the null vector h) back to pixels: H = T2^-1 Hn T1, scaled to H[8] = 1
PAGK_HD bool h_denormalise(const double *h, Sim t1, Sim t2, double *H) {
  const double T1[9] = {t1.s, 0, -t1.s * t1.cx, 0, t1.s, -t1.s * t1.cy, 0, 0, 1};
  const double T2i[9] = {1.0 / t2.s, 0, t2.cx, 0, 1.0 / t2.s, t2.cy, 0, 0, 1};
  double tmp[9];
  mat3_mul(h, T1, tmp);
  mat3_mul(T2i, tmp, H);
  if (!(fabs(H[8]) > 1e-12)) return false;
  const double inv = 1.0 / H[8];
  for (int i = 0; i < 9; ++i) H[i] *= inv;
  return true;
}

// rank-2 constraint and back to pixels: F = T2^T (Fn (I - v v^T)) T1 with v the right singular vector of the smallest
// singular value of Fn; scaled to F[8] = 1 when that entry is not (numerically) zero, as OpenCV does
PAGK_HD bool f_finish(const double *f, Sim t1, Sim t2, double *F) {
  double FtF[9], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) FtF[i * 3 + j] = f[i] * f[j] + f[3 + i] * f[3 + j] + f[6 + i] * f[6 + j];
  const int k = jacobi_smallest<3>(FtF, V);
  const double v0 = V[k], v1 = V[3 + k], v2 = V[6 + k];
  double F2[9];
  for (int i = 0; i < 3; ++i) {
    const double fv = f[i * 3] * v0 + f[i * 3 + 1] * v1 + f[i * 3 + 2] * v2;
    F2[i * 3] = f[i * 3] - fv * v0; F2[i * 3 + 1] = f[i * 3 + 1] - fv * v1; F2[i * 3 + 2] = f[i * 3 + 2] - fv * v2;
  }
  const double T1[9] = {t1.s, 0, -t1.s * t1.cx, 0, t1.s, -t1.s * t1.cy, 0, 0, 1};
  const double T2t[9] = {t2.s, 0, 0, 0, t2.s, 0, -t2.s * t2.cx, -t2.s * t2.cy, 1};
  double tmp[9];
  mat3_mul(F2, T1, tmp);
  mat3_mul(T2t, tmp, F);
  double nrm = 0.0;
  for (int i = 0; i < 9; ++i) nrm += F[i] * F[i];
  if (!(nrm > 0.0) || nrm != nrm) return false;
  const double sc = fabs(F[8]) > 1e-12 * sqrt(nrm) ? 1.0 / F[8] : 1.0 / sqrt(nrm);
  for (int i = 0; i < 9; ++i) F[i] *= sc;
  return true;
}

// cv::Mat::inv() of a 3 x 3 CV_64F: OpenCV's closed form, adjugate times 1/det; zeros for a singular matrix
PAGK_HD void inv3(const double *S, double *D) {
  const double det = S[0] * (S[4] * S[8] - S[5] * S[7]) - S[1] * (S[3] * S[8] - S[5] * S[6]) + S[2] * (S[3] * S[7] - S[4] * S[6]);
  if (det == 0.) { for (int i = 0; i < 9; ++i) D[i] = 0.; return; }
  const double d = 1. / det;
  double t[9];
  t[0] = (S[4] * S[8] - S[5] * S[7]) * d; t[1] = (S[2] * S[7] - S[1] * S[8]) * d; t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
  t[3] = (S[5] * S[6] - S[3] * S[8]) * d; t[4] = (S[0] * S[8] - S[2] * S[6]) * d; t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
  t[6] = (S[3] * S[7] - S[4] * S[6]) * d; t[7] = (S[1] * S[6] - S[0] * S[7]) * d; t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
  for (int i = 0; i < 9; ++i) D[i] = t[i];
}

// forward reprojection error of the homography, squared (OpenCV's HomographyEstimatorCallback::computeError)
PAGK_HD double h_error(const double *H, double x, double y, double u, double v) {
  const double w = H[6] * x + H[7] * y + H[8];
  const double iw = fabs(w) > 1e-300 ? 1.0 / w : 0.0;
  const double du = (H[0] * x + H[1] * y + H[2]) * iw - u, dv = (H[3] * x + H[4] * y + H[5]) * iw - v;
  return du * du + dv * dv;
}
// the larger of the two squared point-to-epipolar-line distances (OpenCV's FMEstimatorCallback::computeError)
PAGK_HD double f_error(const double *F, double x, double y, double u, double v) {
  const double a = F[0] * x + F[1] * y + F[2], b = F[3] * x + F[4] * y + F[5], c = F[6] * x + F[7] * y + F[8];
  const double d2 = u * a + v * b + c, s2 = 1.0 / (a * a + b * b);
  const double a1 = F[0] * u + F[3] * v + F[6], b1 = F[1] * u + F[4] * v + F[7], c1 = F[2] * u + F[5] * v + F[8];
  const double d1 = x * a1 + y * b1 + c1, s1 = 1.0 / (a1 * a1 + b1 * b1);
  const double e2 = d2 * d2 * s2, e1 = d1 * d1 * s1;
  return e1 > e2 ? e1 : e2;
}

// three of the sample's points (nearly) on a line in either image: a degenerate homography sample
PAGK_HD bool h_sample_degenerate(const double *x, const double *y, const double *u, const double *v) {
  for (int a = 0; a < 4; ++a)
    for (int b = a + 1; b < 4; ++b)
      for (int c = b + 1; c < 4; ++c) {
        const double t1 = (x[b] - x[a]) * (y[c] - y[a]) - (y[b] - y[a]) * (x[c] - x[a]);
        const double t2 = (u[b] - u[a]) * (v[c] - v[a]) - (v[b] - v[a]) * (u[c] - u[a]);
        if (fabs(t1) < 1e-3 || fabs(t2) < 1e-3) return true;
        if ((t1 > 0) != (t2 > 0)) return true;  // the orientation of a triangle flips: not a plane seen from one side
      }
  return false;
}

// similarity that moves the centroid to the origin and the mean distance to sqrt(2)
PAGK_HD Sim hartley(const double *x, const double *y, int n) {
  double cx = 0, cy = 0;
  for (int i = 0; i < n; ++i) { cx += x[i]; cy += y[i]; }
  cx /= n; cy /= n;
  double d = 0;
  for (int i = 0; i < n; ++i) d += sqrt((x[i] - cx) * (x[i] - cx) + (y[i] - cy) * (y[i] - cy));
  d /= n;
  Sim t;
  t.cx = cx; t.cy = cy; t.s = d > 1e-12 ? 1.4142135623730951 / d : 1.0;
  return t;
}

// homography of a minimal sample (4 correspondences)
PAGK_HD bool h_from_4(const double *x, const double *y, const double *u, const double *v, double *H) {
  if (h_sample_degenerate(x, y, u, v)) return false;
  const Sim t1 = hartley(x, y, 4), t2 = hartley(u, v, 4);
  double M[81], V[81];
  for (int i = 0; i < 81; ++i) M[i] = 0.0;
  for (int k = 0; k < 4; ++k) {
    double r0[9], r1[9];
    h_rows(t1.s * (x[k] - t1.cx), t1.s * (y[k] - t1.cy), t2.s * (u[k] - t2.cx), t2.s * (v[k] - t2.cy), r0, r1);
    for (int i = 0; i < 9; ++i)
      for (int j = 0; j < 9; ++j) M[i * 9 + j] += r0[i] * r0[j] + r1[i] * r1[j];
  }
  const int k = jacobi_smallest<9>(M, V);
  double h[9];
  for (int i = 0; i < 9; ++i) h[i] = V[i * 9 + k];
  return h_denormalise(h, t1, t2, H);
}

// fundamental matrix of a minimal sample of the eight-point algorithm
PAGK_HD bool f_from_8(const double *x, const double *y, const double *u, const double *v, double *F) {
  const Sim t1 = hartley(x, y, 8), t2 = hartley(u, v, 8);
  double M[81], V[81];
  for (int i = 0; i < 81; ++i) M[i] = 0.0;
  for (int k = 0; k < 8; ++k) {
    double r[9];
    f_row(t1.s * (x[k] - t1.cx), t1.s * (y[k] - t1.cy), t2.s * (u[k] - t2.cx), t2.s * (v[k] - t2.cy), r);
    for (int i = 0; i < 9; ++i)
      for (int j = 0; j < 9; ++j) M[i * 9 + j] += r[i] * r[j];
  }
  const int k = jacobi_smallest<9>(M, V);
  double f[9];
  for (int i = 0; i < 9; ++i) f[i] = V[i * 9 + k];
  return f_finish(f, t1, t2, F);
}

// one Gauss-Newton step's normal equations entry for the homography polish (8 parameters, H[8] = 1): the Jacobian rows of
// the forward reprojection residual (du, dv) of one correspondence
PAGK_HD void h_jacobian(const double *H, double x, double y, double u, double v, double *ju, double *jv, double *ru, double *rv) {
  const double w = H[6] * x + H[7] * y + H[8], iw = 1.0 / w;
  const double px = (H[0] * x + H[1] * y + H[2]) * iw, py = (H[3] * x + H[4] * y + H[5]) * iw;
  ju[0] = x * iw; ju[1] = y * iw; ju[2] = iw; ju[3] = 0; ju[4] = 0; ju[5] = 0; ju[6] = -px * x * iw; ju[7] = -px * y * iw;
  jv[0] = 0; jv[1] = 0; jv[2] = 0; jv[3] = x * iw; jv[4] = y * iw; jv[5] = iw; jv[6] = -py * x * iw; jv[7] = -py * y * iw;
  *ru = px - u; *rv = py - v;
}

// solve the symmetric positive definite 8 x 8 system A d = b (Gaussian elimination with partial pivoting; A destroyed)
PAGK_HD bool solve8(double *A, double *b, double *d) {
  constexpr int N = 8;
  for (int c = 0; c < N; ++c) {
    int p = c;
    for (int r = c + 1; r < N; ++r)
      if (fabs(A[r * N + c]) > fabs(A[p * N + c])) p = r;
    if (!(fabs(A[p * N + c]) > 1e-300)) return false;
    if (p != c) {
      for (int k = 0; k < N; ++k) { const double t = A[c * N + k]; A[c * N + k] = A[p * N + k]; A[p * N + k] = t; }
      const double t = b[c]; b[c] = b[p]; b[p] = t;
    }
    for (int r = c + 1; r < N; ++r) {
      const double f = A[r * N + c] / A[c * N + c];
      for (int k = c; k < N; ++k) A[r * N + k] -= f * A[c * N + k];
      b[r] -= f * b[c];
    }
  }
  for (int r = N - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < N; ++k) s -= A[r * N + k] * d[k];
    d[r] = s / A[r * N + r];
  }
  return true;
}

}  // namespace pagk_ransac
