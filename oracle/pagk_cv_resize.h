// pagk_cv_resize.h -- cv::resize(src, dst, Size(cols*0.5, rows*0.5)) with the default INTER_LINEAR on CV_8UC1
// (src/patch_match.cpp:69-70 of the reference).  TEST INFRASTRUCTURE ONLY; shared by the restatement
// (pagk_oracle.cpp) and by the stand-in cv::resize of the reference build (ref_harness.cpp).
// OpenCV switches to the INTER_AREA 2x2 fast path when both scale factors are exactly 2, otherwise it runs the
// 11-bit fixed-point bilinear.  PINNED bit-exact against cv2 4.13 (tests/golden/pyramid.npz).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <vector>

namespace pagk_cv {
inline void resize_half(const uint8_t *src, int scols, int srows, int sstep, uint8_t *dst, int dcols, int drows) {
  if (scols == 2 * dcols && srows == 2 * drows) {
    for (int y = 0; y < drows; ++y) {
      const uint8_t *r0 = src + (size_t)(2 * y) * sstep, *r1 = r0 + sstep;
      uint8_t *d = dst + (size_t)y * dcols;
      for (int x = 0; x < dcols; ++x)
        d[x] = (uint8_t)((r0[2 * x] + r0[2 * x + 1] + r1[2 * x] + r1[2 * x + 1] + 2) >> 2);
    }
    return;
  }
  const double sx = (double)scols / dcols, sy = (double)srows / drows;
  std::vector<int> xofs(dcols), a0(dcols), a1(dcols);
  for (int dx = 0; dx < dcols; ++dx) {
    float fx = (float)((dx + 0.5) * sx - 0.5);
    int ix = (int)std::floor(fx);
    fx -= ix;
    if (ix < 0) { ix = 0; fx = 0.f; }
    if (ix >= scols - 1) { ix = scols - 1; fx = 0.f; }
    xofs[dx] = ix;
    a0[dx] = (int)(short)std::lrint((1.f - fx) * 2048.f);
    a1[dx] = (int)(short)std::lrint(fx * 2048.f);
  }
  std::vector<int> t0(dcols), t1(dcols);
  for (int dy = 0; dy < drows; ++dy) {
    float fy = (float)((dy + 0.5) * sy - 0.5);
    int iy = (int)std::floor(fy);
    fy -= iy;
    const int y0 = std::min(std::max(iy, 0), srows - 1), y1 = std::min(std::max(iy + 1, 0), srows - 1);
    const int b0 = (int)(short)std::lrint((1.f - fy) * 2048.f), b1 = (int)(short)std::lrint(fy * 2048.f);
    const uint8_t *r0 = src + (size_t)y0 * sstep, *r1 = src + (size_t)y1 * sstep;
    for (int dx = 0; dx < dcols; ++dx) {
      const int ix = xofs[dx], ix1 = std::min(ix + 1, scols - 1);
      t0[dx] = r0[ix] * a0[dx] + r0[ix1] * a1[dx];
      t1[dx] = r1[ix] * a0[dx] + r1[ix1] * a1[dx];
    }
    uint8_t *d = dst + (size_t)dy * dcols;
    for (int dx = 0; dx < dcols; ++dx) {
      const int v = (((b0 * (t0[dx] >> 4)) >> 16) + ((b1 * (t1[dx] >> 4)) >> 16) + 2) >> 2;
      d[dx] = (uint8_t)std::min(std::max(v, 0), 255);
    }
  }
}
}  // namespace pagk_cv
