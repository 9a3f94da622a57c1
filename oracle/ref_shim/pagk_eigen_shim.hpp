// pagk_eigen_shim.hpp -- stand-in for the part of Eigen3 the reference's hot-path sources use.
// TEST INFRASTRUCTURE ONLY.  Eigen is not installed here and there is no network; this header lets the
// reference's unmodified src/patch_match.cpp compile (see pagk_cv_shim.hpp for the whole arrangement).
//
// Used by the reference: Eigen::Matrix4d / Vector4d (Zero(), 4-scalar constructor, unary minus,
// `vector * scalar`, `+=`, `J * J.transpose()`, `H.llt().solve(b)`, `update[0]`, `update.norm()`),
// src/patch_match.cpp:212-214, 256-264, 293-296, 310-312, 319, 343; Matrix<double,5,*> and
// Matrix<double,15,15> appear only as typedefs / a return type.
//
// Arithmetic restated from Eigen 3.3.4 (the version of Ubuntu 18.04, README.md:21), x86-64 SSE2, no FMA.
// PARITY UNPINNED for this header: no Eigen exists here to check it against.
//   * element-wise expressions (`b += -J * e`, `H += J * J.transpose()`): one product and one sum per
//     coefficient, evaluated lazily by Eigen coefficient by coefficient -- no reassociation possible;
//   * LLT: llt_inplace<double, Lower>::unblocked (size < 32 never takes the blocked path):
//       for k: x = m(k,k) - (m(k,0)^2 + m(k,1)^2 + ...)   (squaredNorm of a strided row: sequential)
//              if (x <= 0) stop, leaving the remaining columns untouched (info() = NumericalIssue; the
//              reference never looks at info() and solve() runs on the partial factor)
//              m(k,k) = sqrt(x);  A21 -= A20 * A10^T  (dynamic-size gemv, one column at a time:
//              m(i,k) += m(i,j) * (-1 * m(k,j)) for j = 0..k-1);  A21 /= sqrt(x)
//   * solve: matrixL().solveInPlace then matrixU().solveInPlace through triangular_solver_unroller
//     (fixed size 4): rhs(i) -= (row segment . rhs segment).sum();  rhs(i) /= diag.  The `.sum()` of a
//     3-term segment is a0 + (a1 + a2) for the strided rows of L (scalar unroller splits in halves) and
//     (a0 + a1) + a2 for the contiguous rows of L^T (one SSE2 packet, then the tail);
//   * norm(): sqrt of squaredNorm, SSE2 packet reduction over 4 doubles: (x0^2 + x2^2) + (x1^2 + x3^2).
#pragma once

#include <atomic>
#include <cmath>
#include <cstring>

// every Gauss-Newton pass of the reference calls H.llt() exactly once (src/patch_match.cpp:319): the number of calls is
// the feature x iteration count of the throughput metric, which the reference itself does not expose
namespace pagk_eigen_shim { extern std::atomic<long long> g_llt_calls; }

namespace Eigen {

template <typename M> class LLT;

template <typename T, int R, int C> class Matrix {
 public:
  T d[R * C];  // column-major, as Eigen's default
  Matrix() {}
  Matrix(T a0, T a1, T a2, T a3) {
    static_assert(R * C == 4, "4-scalar constructor is for 4-vectors");
    d[0] = a0; d[1] = a1; d[2] = a2; d[3] = a3;
  }
  static Matrix Zero() { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = T(0); return m; }
  T &operator()(int r, int c) { return d[c * R + r]; }
  const T &operator()(int r, int c) const { return d[c * R + r]; }
  T &operator()(int i) { return d[i]; }
  const T &operator()(int i) const { return d[i]; }
  T &operator[](int i) { return d[i]; }
  const T &operator[](int i) const { return d[i]; }
  Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = -d[i]; return m; }
  Matrix operator*(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] * s; return m; }
  friend Matrix operator*(T s, const Matrix &a) { return a * s; }
  Matrix operator+(const Matrix &o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] + o.d[i]; return m; }
  Matrix operator-(const Matrix &o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] - o.d[i]; return m; }
  Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; ++i) d[i] = d[i] + o.d[i]; return *this; }
  Matrix &operator-=(const Matrix &o) { for (int i = 0; i < R * C; ++i) d[i] = d[i] - o.d[i]; return *this; }
  Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> m;
    for (int r = 0; r < R; ++r) for (int c = 0; c < C; ++c) m(c, r) = (*this)(r, c);
    return m;
  }
  // outer product only (inner dimension 1): each coefficient is a single product
  template <int K> Matrix<T, R, K> operator*(const Matrix<T, C, K> &o) const {
    static_assert(C == 1, "the stand-in implements only vector * row-vector");
    Matrix<T, R, K> m;
    for (int r = 0; r < R; ++r) for (int k = 0; k < K; ++k) m(r, k) = (*this)(r, 0) * o(0, k);
    return m;
  }
  T norm() const {
    static_assert(R * C == 4, "norm() is restated for 4-vectors");
    return std::sqrt((d[0] * d[0] + d[2] * d[2]) + (d[1] * d[1] + d[3] * d[3]));
  }
  LLT<Matrix> llt() const { pagk_eigen_shim::g_llt_calls.fetch_add(1, std::memory_order_relaxed); return LLT<Matrix>(*this); }
};

template <typename T, int N> class LLT<Matrix<T, N, N>> {
 public:
  typedef Matrix<T, N, N> Mat;
  typedef Matrix<T, N, 1> Vec;
  Mat m;
  int failed_at;
  explicit LLT(const Mat &a) : m(a), failed_at(-1) {
    for (int k = 0; k < N; ++k) {
      const int rs = N - k - 1;
      T x = m(k, k);
      if (k > 0) {
        T s = m(k, 0) * m(k, 0);
        for (int j = 1; j < k; ++j) s = s + m(k, j) * m(k, j);
        x -= s;
      }
      if (x <= T(0)) { failed_at = k; break; }
      m(k, k) = x = std::sqrt(x);
      if (k > 0 && rs > 0)
        for (int j = 0; j < k; ++j) {
          const T t = T(-1) * m(k, j);
          for (int i = k + 1; i < N; ++i) m(i, k) = m(i, k) + m(i, j) * t;
        }
      for (int i = k + 1; i < N; ++i) m(i, k) = m(i, k) / x;
    }
  }
  Vec solve(const Vec &b) const {
    static_assert(N == 4, "solve() is restated for the 4x4 case the reference uses");
    Vec r = b;
    // L y = b, rows of L are strided in a column-major matrix: scalar redux, halves
    r[0] /= m(0, 0);
    r[1] -= m(1, 0) * r[0];
    r[1] /= m(1, 1);
    r[2] -= (m(2, 0) * r[0] + m(2, 1) * r[1]);
    r[2] /= m(2, 2);
    r[3] -= (m(3, 0) * r[0] + (m(3, 1) * r[1] + m(3, 2) * r[2]));
    r[3] /= m(3, 3);
    // L^T x = y, row i of L^T is column i of L (contiguous): packet of two, then the tail
    r[3] /= m(3, 3);
    r[2] -= m(3, 2) * r[3];
    r[2] /= m(2, 2);
    r[1] -= (m(2, 1) * r[2] + m(3, 1) * r[3]);
    r[1] /= m(1, 1);
    r[0] -= ((m(1, 0) * r[1] + m(2, 0) * r[2]) + m(3, 0) * r[3]);
    r[0] /= m(0, 0);
    return r;
  }
};

typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 3, 1> Vector3d;

}  // namespace Eigen
