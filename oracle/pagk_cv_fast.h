// pagk_cv_fast.h -- cv::FAST(image, keypoints, threshold, nonmaxSuppression, TYPE_9_16), OpenCV modules/features2d/src/fast.cpp
// (FAST_t<16>, cornerScore<16>), restated from its published algorithm.  TEST INFRASTRUCTURE ONLY; shared by the restatement
// (pagk_oracle.cpp) and by the stand-in cv::FAST of the reference build (ref_harness.cpp).
// PINNED bit-exact against cv2 4.13 (tests/golden/fast.npz): positions, OpenCV's row-major order, responses.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstddef>
#include <vector>

namespace pagk_cv {
const int kFastDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kFastDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

inline int fast_corner_score(const uint8_t *ptr, const int *pixel, int threshold) {
  const int K = 8, N = K * 3 + 1;
  const int v = ptr[0];
  short d[N];
  for (int k = 0; k < N; k++) d[k] = (short)(v - ptr[pixel[k]]);
  int a0 = threshold;
  for (int k = 0; k < 16; k += 2) {
    int a = std::min((int)d[k + 1], (int)d[k + 2]);
    a = std::min(a, (int)d[k + 3]);
    if (a <= a0) continue;
    a = std::min(a, (int)d[k + 4]); a = std::min(a, (int)d[k + 5]); a = std::min(a, (int)d[k + 6]);
    a = std::min(a, (int)d[k + 7]); a = std::min(a, (int)d[k + 8]);
    a0 = std::max(a0, std::min(a, (int)d[k]));
    a0 = std::max(a0, std::min(a, (int)d[k + 9]));
  }
  int b0 = -a0;
  for (int k = 0; k < 16; k += 2) {
    int b = std::max((int)d[k + 1], (int)d[k + 2]);
    b = std::max(b, (int)d[k + 3]); b = std::max(b, (int)d[k + 4]); b = std::max(b, (int)d[k + 5]);
    if (b >= b0) continue;
    b = std::max(b, (int)d[k + 6]); b = std::max(b, (int)d[k + 7]); b = std::max(b, (int)d[k + 8]);
    b0 = std::min(b0, std::max(b, (int)d[k]));
    b0 = std::min(b0, std::max(b, (int)d[k + 9]));
  }
  return -b0 - 1;
}

inline int fast_detect(const uint8_t *img, int cols, int rows, int step, int threshold, bool nonmax, const uint8_t *mask, int max_out,
                float *xy, float *response) {
  const int K = 8, N = 25;
  int pixel[25];
  for (int k = 0; k < 16; ++k) pixel[k] = kFastDx[k] + kFastDy[k] * step;
  for (int k = 16; k < 25; ++k) pixel[k] = pixel[k - 16];
  threshold = std::min(std::max(threshold, 0), 255);
  std::vector<uint8_t> score((size_t)rows * cols, 0), corner((size_t)rows * cols, 0);
  for (int i = 3; i < rows - 3; ++i)
    for (int j = 3; j < cols - 3; ++j) {
      const uint8_t *ptr = img + (size_t)i * step + j;
      const int v = ptr[0];
      bool is = false;
      {  // nine contiguous pixels of the circle darker than v - threshold ...
        const int vt = v - threshold;
        int count = 0;
        for (int k = 0; k < N && !is; k++) { if (ptr[pixel[k]] < vt) { if (++count > K) is = true; } else count = 0; }
      }
      if (!is) {  // ... or brighter than v + threshold
        const int vt = v + threshold;
        int count = 0;
        for (int k = 0; k < N && !is; k++) { if (ptr[pixel[k]] > vt) { if (++count > K) is = true; } else count = 0; }
      }
      if (is) {
        corner[(size_t)i * cols + j] = 1;
        if (nonmax) score[(size_t)i * cols + j] = (uint8_t)fast_corner_score(ptr, pixel, threshold);
      }
    }
  int n = 0;
  for (int i = 3; i < rows - 3; ++i)
    for (int j = 3; j < cols - 3; ++j) {
      if (!corner[(size_t)i * cols + j]) continue;
      const uint8_t *sc = &score[(size_t)i * cols + j];
      const int s0 = sc[0];
      if (nonmax && !(s0 > sc[1] && s0 > sc[-1] && s0 > sc[-cols - 1] && s0 > sc[-cols] && s0 > sc[-cols + 1] && s0 > sc[cols - 1] &&
                      s0 > sc[cols] && s0 > sc[cols + 1]))
        continue;
      if (mask && !mask[(size_t)i * cols + j]) continue;
      if (n < max_out) { xy[2 * n] = (float)j; xy[2 * n + 1] = (float)i; response[n] = (float)s0; }
      ++n;
    }
  return n;
}
}  // namespace pagk_cv
