// stand-in for glog (include/gyro_aided_tracker.h:38 includes "../Thirdparty/glog/include/glog/logging.h";
// the reference ships glog only as Thirdparty/glog-master.tar.xz).  LOG(x) streams to stderr.
#pragma once
#include <iostream>
namespace pagk_glog_shim {
struct Line {
  ~Line() { std::cerr << std::endl; }
  template <typename T> Line &operator<<(const T &v) { std::cerr << v; return *this; }
};
}  // namespace pagk_glog_shim
#define LOG(severity) ::pagk_glog_shim::Line() << "[" #severity "] "
