// Dumps what include/pagk_sequence.hpp parses from a dataset directory, for tests/test_sequence.py to compare with the
// Python mirror:  sequence_dump <settings.yaml> <datasetDir> <keypointDir>
#include <cstdio>

#include "pagk_sequence.hpp"

int main(int argc, char **argv) {
  if (argc != 4) return 2;
  try {
    const pagk::Settings s = pagk::loadConfigureFile(argv[1]);
    std::printf("settings %a %a %a %a %a %a %a %a %d %d %d %d %a %d %d %d %s|%s|%s\n", s.camera.mK[0], s.camera.mK[4], s.camera.mK[2],
                s.camera.mK[5], s.camera.mDistCoef[0], s.camera.mDistCoef[1], s.camera.mDistCoef[2], s.camera.mDistCoef[3],
                (int)s.has_k3, s.camera.width, s.camera.height, s.fps, s.threshold_of_predict_new_keypoint, s.keypoint_number,
                s.half_patch_size, (int)s.loadDetectedKeypoints, s.dataset.c_str(), s.datasetDir.c_str(), s.detectedKeypointsFile.c_str());
    std::printf("tbc");
    for (float v : s.imuCalib.Tbc) std::printf(" %a", v);
    std::printf("\n");
    pagk::ImageFileList list(argv[2]);
    pagk::ImuFile imu(std::string(argv[2]) + "/imu.txt");
    pagk::ImuFeed feed(imu);
    std::string path;
    double time_cur = 0, time_prev = 0;
    std::vector<pagk::ImuPoint> vImuMeas;
    const auto table = pagk::loadTimeCorrespondences(std::string(argv[3]) + "/corresponds.txt");
    while (list.getNextFrame(path, time_cur)) {
      feed.window(time_prev, time_cur, vImuMeas);
      const int idx = pagk::findTimeCorrespondenIndex(table, time_cur);
      std::printf("frame %a %s %zu %d\n", time_cur, path.c_str(), vImuMeas.size(), idx);
      for (const auto &m : vImuMeas) std::printf("imu %a %a %a %a %a %a %a\n", m.t, m.a.x, m.a.y, m.a.z, m.w.x, m.w.y, m.w.z);
      if (idx >= 0) {
        const auto pts = pagk::loadDetectedKeypoints(std::string(argv[3]) + "/" + table[idx].second + ".txt");
        double sx = 0, sy = 0;
        for (const auto &p : pts) { sx += p.x; sy += p.y; }
        std::printf("keypoints %zu %a %a\n", pts.size(), sx, sy);
      }
      time_prev = time_cur;
    }
  } catch (const std::exception &e) {
    std::printf("error %s\n", e.what());
    return 1;
  }
  return 0;
}
