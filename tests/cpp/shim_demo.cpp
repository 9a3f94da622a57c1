// shim_demo.cpp -- a reference-style driver written against include/pagk_tracker.hpp.
// Mirrors the call sequence of Examples/Demo/RealSenseD435i.cpp:244-254 on one synthetic frame pair and
// prints the results; tests/test_cpp_shim.py rebuilds the same inputs in numpy and checks them against
// the CPU oracle.  Build: g++ -std=c++17 -I include tests/cpp/shim_demo.cpp -L <csrc> -lpagk_cuda
#include <cstdio>
#include <cstdint>
#include <vector>

#include "pagk_tracker.hpp"

static std::vector<uint8_t> noise_image(int w, int h, uint64_t seed) {  // 5x5 box blur of LCG noise, integer only
  std::vector<uint32_t> n((size_t)(w + 4) * (h + 4));
  uint64_t s = seed;
  for (auto &v : n) { s = s * 6364136223846793005ULL + 1442695040888963407ULL; v = (uint32_t)(s >> 56); }
  std::vector<uint8_t> img((size_t)w * h);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint32_t acc = 0;
      for (int dy = 0; dy < 5; ++dy)
        for (int dx = 0; dx < 5; ++dx) acc += n[(size_t)(y + dy) * (w + 4) + x + dx];
      img[(size_t)y * w + x] = (uint8_t)(acc / 25);
    }
  return img;
}

int main() {
  const int W = 240, H = 180, N = 64;
  const std::vector<uint8_t> big = noise_image(W + 8, H + 8, 12345);
  std::vector<uint8_t> ref((size_t)W * H), cur((size_t)W * H);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      ref[(size_t)y * W + x] = big[(size_t)(y + 4) * (W + 8) + x + 4];
      cur[(size_t)y * W + x] = big[(size_t)(y + 3) * (W + 8) + x + 2];  // cur(x, y) = ref(x - 2, y - 1)
    }
  pagk::CameraParams cam;
  cam.mK = {200.f, 0.f, 120.f, 0.f, 200.f, 90.f, 0.f, 0.f, 1.f};
  cam.mDistCoef = {0.f, 0.f, 0.f, 0.f};
  cam.width = W; cam.height = H;
  pagk::Frame lastFrame, curFrame;
  lastFrame.mTimeStamp = 10.0; curFrame.mTimeStamp = 10.05;
  lastFrame.mGray = {ref.data(), W, H, W}; curFrame.mGray = {cur.data(), W, H, W};
  lastFrame.mpCameraParams = curFrame.mpCameraParams = &cam;
  for (int i = 0; i < N; ++i) {
    pagk::KeyPoint kp;
    kp.pt = pagk::Point2f(30.f + (float)((i * 37) % 180) + 0.25f * (float)(i % 4), 30.f + (float)((i * 53) % 120) + 0.5f * (float)(i % 2));
    lastFrame.mvKeys.push_back(kp); lastFrame.mvKeysUn.push_back(kp);
  }
  for (int k = 0; k < 12; ++k)  // stationary gyro, 200 Hz
    curFrame.mvImuFromLastFrame.emplace_back(pagk::Point3f(0, 0, 9.8f), pagk::Point3f(0, 0, 0), 9.998 + 0.005 * k);
  pagk::ImuCalib imuCalib;
  imuCalib.Tbc = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  const pagk::Point3f biasg(0, 0, 0);
  try {
    pagk::Device dev(0, W, H, N);
    pagk::GyroAidedTracker trk(dev, lastFrame, curFrame, imuCalib, biasg, nullptr,
                               pagk::GyroAidedTracker::GYRO_PREDICT_WITH_OPTICAL_FLOW_REFINED_CONSIDER_ILLUMINATION_DEFORMATION,
                               pagk::GyroAidedTracker::PIXEL_AWARE_PREDICTION, "", 5);
    const int n = trk.TrackFeatures();
    trk.SetBackToFrame(curFrame);
    std::printf("n_predict %d iterations %lld\n", n, trk.mIterations);
    for (int i = 0; i < N; ++i)
      std::printf("%d %d %.9g %.9g\n", i, (int)curFrame.mvStatus[i], curFrame.mvPtPredictUn[i].x, curFrame.mvPtPredictUn[i].y);
    // PatchMatch on its own, as GyroPredictFeaturesAndOpticalFlowRefined uses it (src/gyro_aided_tracker.cpp:280-283)
    pagk::GyroAidedTracker trk2(dev, lastFrame, curFrame, imuCalib, biasg, nullptr, pagk::GyroAidedTracker::GYRO_PREDICT);
    trk2.TrackFeatures();
    pagk::PatchMatch pm(&trk2, 5, 10, 3, true, false, true, true, false);
    pm.OpticalFlowMultiLevel();
    int ok = 0;
    for (int i = 0; i < N; ++i) ok += trk2.mvStatusAfterPatchMatched[i];
    std::printf("patch_match_ok %d\n", ok);
    // GeometryValidation() with the models the caller's cv::findHomography / cv::findFundamentalMat would hand over: here the
    // true motion of the synthetic pair (a shift by (2, 1)) and an arbitrary fundamental matrix
    const double H21[9] = {1, 0, 2, 0, 1, 1, 0, 0, 1}, F21[9] = {0, -1e-3, 0.2, 1e-3, 0, -0.3, -0.2, 0.3, 0.01};
    const int n_in = trk.GeometryValidation(H21, F21);
    std::printf("geometry %d %d %.9g %.9g\n", n_in, (int)trk.mGeometryUsedH, trk.mGeometryScoreH, trk.mGeometryScoreF);
    trk.SetType(pagk::GyroAidedTracker::OPENCV_OPTICAL_FLOW_PYR_LK);
    std::printf("unsupported %d\n", trk.TrackFeatures());
  } catch (const pagk::Error &e) {
    std::printf("error %d %s\n", e.code, e.what());
    return e.code == PAGK_ERR_NO_DEVICE ? 3 : 1;
  }
  return 0;
}
