// Host harness of csrc/pagk_ransac.h for the tests: the same estimator code the device kernel runs (pagk_ransac_kernel,
// csrc/pagk_kernels.cu), executed sequentially -- same generator, same hypotheses, same order of every sum.  Test
// infrastructure: built by tests/test_ransac.py with g++, never loaded by the product.
#include <cstring>
#include <vector>
#include "../../pixel_aware_gyro_aided_klt_feature_tracker_b200/csrc/pagk_ransac.h"

using namespace pagk_ransac;

namespace {
struct Pts { std::vector<double> x, y, u, v; };

int best_hypothesis(const Pts &p, int model, unsigned seed, unsigned pair, double *Mdl) {
  const int M = (int)p.x.size();
  int best_cnt = -1, best_hyp = -1;
  for (int hyp = 0; hyp < kHypotheses; ++hyp) {
    double m[9];
    bool ok;
    if (model == 0) {
      int id[4]; sample<4>(seed, pair, (unsigned)hyp, M, id);
      double x[4], y[4], u[4], v[4];
      for (int k = 0; k < 4; ++k) { x[k] = p.x[id[k]]; y[k] = p.y[id[k]]; u[k] = p.u[id[k]]; v[k] = p.v[id[k]]; }
      ok = h_from_4(x, y, u, v, m);
    } else {
      int id[8]; sample<8>(seed, pair, (unsigned)hyp + kHypotheses, M, id);
      double x[8], y[8], u[8], v[8];
      for (int k = 0; k < 8; ++k) { x[k] = p.x[id[k]]; y[k] = p.y[id[k]]; u[k] = p.u[id[k]]; v[k] = p.v[id[k]]; }
      ok = f_from_8(x, y, u, v, m);
    }
    if (!ok) continue;
    int cnt = 0;
    for (int k = 0; k < M; ++k) cnt += ((model == 0 ? h_error(m, p.x[k], p.y[k], p.u[k], p.v[k]) : f_error(m, p.x[k], p.y[k], p.u[k], p.v[k])) <= kThreshold2);
    if (cnt > best_cnt) { best_cnt = cnt; best_hyp = hyp; memcpy(Mdl, m, sizeof(m)); }
  }
  return best_cnt;
}

void estimate(const Pts &p, int model, unsigned seed, unsigned pair, double *out) {
  const int M = (int)p.x.size();
  double Mdl[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (model == 1) memset(Mdl, 0, sizeof(Mdl));
  const int cnt = best_hypothesis(p, model, seed, pair, Mdl);
  const int need = model == 0 ? 4 : 8;
  if (cnt >= need) {
    std::vector<char> inl((size_t)M);
    for (int k = 0; k < M; ++k) inl[k] = (model == 0 ? h_error(Mdl, p.x[k], p.y[k], p.u[k], p.v[k]) : f_error(Mdl, p.x[k], p.y[k], p.u[k], p.v[k])) <= kThreshold2;
    double s[4] = {0, 0, 0, 0};
    int n = 0;
    for (int k = 0; k < M; ++k) if (inl[k]) { s[0] += p.x[k]; s[1] += p.y[k]; s[2] += p.u[k]; s[3] += p.v[k]; ++n; }
    for (double &q : s) q /= n;
    double d1 = 0, d2 = 0;
    for (int k = 0; k < M; ++k) if (inl[k]) {
      d1 += sqrt((p.x[k] - s[0]) * (p.x[k] - s[0]) + (p.y[k] - s[1]) * (p.y[k] - s[1]));
      d2 += sqrt((p.u[k] - s[2]) * (p.u[k] - s[2]) + (p.v[k] - s[3]) * (p.v[k] - s[3]));
    }
    d1 /= n; d2 /= n;
    const Sim t1 = {d1 > 1e-12 ? 1.4142135623730951 / d1 : 1.0, s[0], s[1]}, t2 = {d2 > 1e-12 ? 1.4142135623730951 / d2 : 1.0, s[2], s[3]};
    double A[81];
    for (int r = 0; r < 9; ++r)
      for (int c = r; c < 9; ++c) {
        double acc = 0.0;
        for (int k = 0; k < M; ++k) if (inl[k]) {
          const double xn = t1.s * (p.x[k] - t1.cx), yn = t1.s * (p.y[k] - t1.cy), un = t2.s * (p.u[k] - t2.cx), vn = t2.s * (p.v[k] - t2.cy);
          if (model == 0) { double r0[9], r1[9]; h_rows(xn, yn, un, vn, r0, r1); acc += r0[r] * r0[c] + r1[r] * r1[c]; }
          else { double rr[9]; f_row(xn, yn, un, vn, rr); acc += rr[r] * rr[c]; }
        }
        A[r * 9 + c] = acc; A[c * 9 + r] = acc;
      }
    double V[81], vec[9], m2[9];
    const int k = jacobi_smallest<9>(A, V);
    for (int i = 0; i < 9; ++i) vec[i] = V[i * 9 + k];
    if (model == 0 ? h_denormalise(vec, t1, t2, m2) : f_finish(vec, t1, t2, m2)) memcpy(Mdl, m2, sizeof(m2));
    if (model == 0)
      for (int it = 0; it < 5; ++it) {
        double JtJ[64], Jtr[8], d[8];
        for (int r = 0; r < 8; ++r) {
          for (int c = r; c < 8; ++c) {
            double acc = 0.0;
            for (int q = 0; q < M; ++q) if (inl[q]) { double ju[8], jv[8], ru, rv; h_jacobian(Mdl, p.x[q], p.y[q], p.u[q], p.v[q], ju, jv, &ru, &rv); acc += ju[r] * ju[c] + jv[r] * jv[c]; }
            JtJ[r * 8 + c] = acc; JtJ[c * 8 + r] = acc;
          }
          double acc = 0.0;
          for (int q = 0; q < M; ++q) if (inl[q]) { double ju[8], jv[8], ru, rv; h_jacobian(Mdl, p.x[q], p.y[q], p.u[q], p.v[q], ju, jv, &ru, &rv); acc += ju[r] * ru + jv[r] * rv; }
          Jtr[r] = acc;
        }
        if (solve8(JtJ, Jtr, d)) for (int i = 0; i < 8; ++i) Mdl[i] -= d[i];
      }
  }
  memcpy(out, Mdl, sizeof(Mdl));
}
}  // namespace

extern "C" int pagk_ransac_host(int n, const float *pts1, const float *pts2, const unsigned char *status, unsigned seed, unsigned pair,
                                double *H21, double *F21) {
  Pts p;
  for (int i = 0; i < n; ++i)
    if (status[i]) { p.x.push_back(pts1[2 * i]); p.y.push_back(pts1[2 * i + 1]); p.u.push_back(pts2[2 * i]); p.v.push_back(pts2[2 * i + 1]); }
  if (p.x.size() <= 8) return 1;
  estimate(p, 0, seed, pair, H21);
  estimate(p, 1, seed, pair, F21);
  return 0;
}
