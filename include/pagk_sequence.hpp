// pagk_sequence.hpp -- header-only C++ readers for the files the reference's demo driver consumes, next to the
// tracker shim (pagk_tracker.hpp).  Host parsing only; no OpenCV.  SURVEY.md section 8f rank 4.
//
//     reference                                                             here (namespace pagk)
//     --------------------------------------------------------------------  ----------------------------------------
//     loadConfigureFile(file)            include/common.h:49-103             Settings loadConfigureFile(file)
//     getNextFrame()                     Examples/Demo/RealSenseD435i.cpp:74-100    ImageFileList::getNextFrame(path, time)
//     getNextIMU(IMU::Point&)            Examples/Demo/RealSenseD435i.cpp:102-141   ImuFile::getNextIMU(ImuPoint&)
//     the IMU loop of main()             Examples/Demo/RealSenseD435i.cpp:196-217   ImuFeed::window(time_prev, time_cur, vImuMeas)
//     corresponds.txt loop of main()     Examples/Demo/RealSenseD435i.cpp:168-182   loadTimeCorrespondences(path)
//     findTimeCorrespondenIndex(v, t)    include/common.h:105-114            findTimeCorrespondenIndex(v, t)
//     Frame::LoadDetectedKeypointFromFile (parsing half) src/frame.cpp:222-240  loadDetectedKeypoints(path)
//
// Decoding the PNGs stays with the caller (the reference uses cv::imread; any decoder that yields 8-bit gray works).
// The Python mirror is pixel_aware_gyro_aided_klt_feature_tracker_b200/sequence.py; tests/test_sequence.py holds both
// to the same values on the same files.
#ifndef PAGK_SEQUENCE_HPP_
#define PAGK_SEQUENCE_HPP_

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "pagk_tracker.hpp"

namespace pagk {

// ---------------------------------------------------------------------------------------------- settings
struct Settings {
  CameraParams camera;                 // mK from Camera.fx/fy/cx/cy, mDistCoef = k1 k2 p1 p2
  float k3 = 0.f; bool has_k3 = false;  // Camera.k3 only when it is a real number (include/common.h:80-81)
  int fps = 0;
  ImuCalib imuCalib;                   // Tbc
  int keypoint_number = 0;
  float threshold_of_predict_new_keypoint = 0.f;
  int half_patch_size = 5;
  bool loadDetectedKeypoints = false;
  std::string detectedKeypointsFile, dataset, datasetDir;
  int imu_frequency = 200;
  std::map<std::string, std::string> raw;  // every scalar `key: value` of the file, comments and quotes stripped
};

namespace detail {
inline std::string trim(const std::string &s) {
  const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
  return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}
inline std::string strip_comment(const std::string &s) {  // '#' outside double quotes starts a comment
  bool q = false;
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '"') q = !q;
    if (s[i] == '#' && !q) return s.substr(0, i);
  }
  return s;
}
inline std::string unquote(const std::string &s) {
  return s.size() >= 2 && s.front() == '"' && s.back() == '"' ? s.substr(1, s.size() - 2) : s;
}
}  // namespace detail

// The YAML subset the reference's settings files use: `key: scalar` lines and one flow sequence (`Tbc: [ ... ]`, possibly
// over several lines).
inline Settings loadConfigureFile(const std::string &file) {
  std::ifstream fin(file.c_str());
  if (!fin.is_open()) throw std::runtime_error("loadConfigureFile: cannot open " + file);
  Settings s;
  std::map<std::string, std::vector<double>> seqs;
  std::string line;
  while (std::getline(fin, line)) {
    line = detail::trim(detail::strip_comment(line));
    if (line.empty() || line[0] == '%' || line == "---") continue;
    const size_t c = line.find(':');
    if (c == std::string::npos) continue;
    const std::string key = detail::trim(line.substr(0, c));
    std::string val = detail::trim(line.substr(c + 1));
    if (val.empty() || val[0] == '[') {  // flow sequence, read up to the closing bracket
      std::string body = val;
      while (body.find(']') == std::string::npos && std::getline(fin, line)) body += " " + detail::strip_comment(line);
      const size_t a = body.find('['), b = body.find(']');
      if (a == std::string::npos || b == std::string::npos) throw std::runtime_error("loadConfigureFile: bad sequence for " + key);
      std::stringstream ss(body.substr(a + 1, b - a - 1));
      std::string item;
      while (std::getline(ss, item, ',')) {
        item = detail::trim(item);
        if (!item.empty()) seqs[key].push_back(std::atof(item.c_str()));
      }
    } else {
      s.raw[key] = detail::unquote(val);
    }
  }
  auto need = [&](const char *k) -> const std::string & {
    auto it = s.raw.find(k);
    if (it == s.raw.end()) throw std::runtime_error(std::string("loadConfigureFile: missing '") + k + "' in " + file);
    return it->second;
  };
  auto f = [&](const char *k) { return (float)std::atof(need(k).c_str()); };  // `float fx = fSettings["Camera.fx"]`
  auto opt = [&](const char *k, const std::string &d) { auto it = s.raw.find(k); return it == s.raw.end() ? d : it->second; };
  s.camera.mK = {f("Camera.fx"), 0.f, f("Camera.cx"), 0.f, f("Camera.fy"), f("Camera.cy"), 0.f, 0.f, 1.f};
  s.camera.mDistCoef = {f("Camera.k1"), f("Camera.k2"), f("Camera.p1"), f("Camera.p2")};
  const std::string k3 = opt("Camera.k3", "");
  if (k3.find_first_of(".eE") != std::string::npos) { s.k3 = (float)std::atof(k3.c_str()); s.has_k3 = true; }  // node.isReal()
  s.camera.width = std::atoi(need("Camera.width").c_str());
  s.camera.height = std::atoi(need("Camera.height").c_str());
  s.fps = std::atoi(need("Camera.fps").c_str());
  const auto tb = seqs.find("Tbc");
  if (tb == seqs.end() || tb->second.size() != 16) throw std::runtime_error("loadConfigureFile: Tbc needs 16 values in " + file);
  for (int i = 0; i < 16; ++i) s.imuCalib.Tbc[i] = (float)tb->second[i];
  s.keypoint_number = std::atoi(need("KeyPointNumber").c_str());
  s.threshold_of_predict_new_keypoint = f("ThresholdOfPredictNewKeyPoint");
  s.half_patch_size = std::atoi(opt("HalfPatchSize", "5").c_str());
  s.loadDetectedKeypoints = std::atoi(opt("LoadDetectedKeypoints", "0").c_str()) == 1;
  s.detectedKeypointsFile = opt("DetectedKeypointsFile", "");
  s.dataset = opt("dataset", "");
  s.datasetDir = opt("datasetDir", "");
  s.imu_frequency = std::atoi(opt("IMU.Frequency", "200").c_str());
  return s;
}

// ---------------------------------------------------------------------------------------------- image list, IMU log
class ImageFileList {  // getNextFrame(): <datasetDir>/image_file_list.txt, one path per line, appended to datasetDir as it stands
 public:
  explicit ImageFileList(const std::string &datasetDir) : dir_(datasetDir), fin_((datasetDir + "/image_file_list.txt").c_str()) {
    if (!fin_.is_open()) throw std::runtime_error("ImageFileList: cannot open " + datasetDir + "/image_file_list.txt");
  }
  // the next image path and its timestamp: the file name between the last '/' and ".png" is nanoseconds
  bool getNextFrame(std::string &strImage, double &time_cur) {
    std::string line;
    while (std::getline(fin_, line)) {
      if (!line.empty() && line.back() == '\r') line.pop_back();
      if (line.empty()) continue;
      strImage = dir_ + line;
      const std::string::size_type pos1 = line.rfind("/"), pos2 = line.rfind(".png");
      if (pos2 == std::string::npos) throw std::runtime_error("image_file_list.txt: no '.png' in " + line);
      time_cur = std::stol(line.substr(pos1 + 1, pos2 - pos1 - 1)) * 1e-9;
      return true;
    }
    return false;
  }

 private:
  std::string dir_;
  std::ifstream fin_;
};

class ImuFile {  // getNextIMU(): `t_ns ax ay az wx wy wz` per line
 public:
  explicit ImuFile(const std::string &path) : fin_(path.c_str()) {
    if (!fin_.is_open()) throw std::runtime_error("ImuFile: cannot open " + path);
  }
  bool getNextIMU(ImuPoint &imu) {
    std::string line;
    while (std::getline(fin_, line)) {
      std::istringstream sin(line);
      double a[3], w[3];
      std::string str_time;
      if (!(sin >> str_time)) continue;  // blank line
      if (!(sin >> a[0] >> a[1] >> a[2] >> w[0] >> w[1] >> w[2])) throw std::runtime_error("imu.txt: expected 7 fields: " + line);
      imu.a = Point3f((float)a[0], (float)a[1], (float)a[2]);
      imu.w = Point3f((float)w[0], (float)w[1], (float)w[2]);
      imu.t = std::stol(str_time) * 1e-9;
      return true;
    }
    return false;
  }

 private:
  std::ifstream fin_;
};

// The IMU loop of the demo's main(): `last_imu` is read ahead; per frame pair skip what is older than time_prev - delay,
// hand out what is older than time_cur - delay.  valid_imu goes false when the file runs out and stays false.
class ImuFeed {
 public:
  explicit ImuFeed(ImuFile &f) : f_(f) { have_ = f_.getNextIMU(last_imu_); }
  void window(double time_prev, double time_cur, std::vector<ImuPoint> &vImuMeas, double delay = 0.0) {
    vImuMeas.clear();
    if (time_prev != 0 && have_) {
      while (last_imu_.t < time_prev - delay && f_.getNextIMU(last_imu_)) continue;
      while (last_imu_.t < time_cur - delay && valid_imu_) {
        vImuMeas.push_back(last_imu_);
        valid_imu_ = f_.getNextIMU(last_imu_);
      }
    }
  }
  bool valid_imu() const { return valid_imu_; }

 private:
  ImuFile &f_;
  ImuPoint last_imu_;
  bool have_ = false, valid_imu_ = true;
};

// ---------------------------------------------------------------------------------------------- detected-keypoint files
inline std::vector<std::pair<double, std::string>> loadTimeCorrespondences(const std::string &path) {
  std::vector<std::pair<double, std::string>> v;
  std::ifstream fin(path.c_str());
  if (!fin.is_open()) throw std::runtime_error("loadTimeCorrespondences: cannot open " + path);
  std::string line;
  while (std::getline(fin, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty()) continue;
    const std::string::size_type p_dot = line.find(",");
    if (p_dot == std::string::npos) throw std::runtime_error("corresponds.txt: no ',' in " + line);
    v.push_back(std::make_pair(std::atof(line.substr(0, p_dot).c_str()), line.substr(p_dot + 2, line.size() - p_dot)));
  }
  return v;
}

inline int findTimeCorrespondenIndex(const std::vector<std::pair<double, std::string>> &vpTimeString, double t) {
  for (size_t i = 0; i < vpTimeString.size(); i++)
    if (std::abs(t - vpTimeString[i].first) < 0.0001) return (int)i;  // found
  return -1;                                                          // not found
}

// `index, x, y` per line; the point is (field 1, field 2)
inline std::vector<Point2f> loadDetectedKeypoints(const std::string &path) {
  std::vector<Point2f> v;
  std::ifstream fin(path.c_str());
  if (!fin.is_open()) throw std::runtime_error("loadDetectedKeypoints: cannot open " + path);
  std::string line;
  while (std::getline(fin, line)) {
    std::istringstream sin(line);
    std::vector<double> data;
    std::string field;
    while (std::getline(sin, field, ',')) data.push_back(std::atof(field.c_str()));
    if (data.empty()) continue;
    if (data.size() < 3) throw std::runtime_error("keypoint file: expected 'index, x, y': " + line);
    v.push_back(Point2f((float)data[1], (float)data[2]));
  }
  return v;
}

}  // namespace pagk

#endif  // PAGK_SEQUENCE_HPP_
