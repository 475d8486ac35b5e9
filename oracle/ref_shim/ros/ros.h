// STAND-IN for the handful of roscpp names src/lvba_system.cpp / include/dataset_io.h use (NOT ROS; test infrastructure,
// see ../mini_eigen.h).  Parameters come from a process-wide table the driver fills (ros::param_table()); publishers swallow
// their messages; ok() is always true.
#pragma once
#include <cstdio>
#include <map>
#include <string>
#include <vector>
namespace ros {
struct ParamValue { double num = 0; std::string str; std::vector<double> vec; int kind = 0; };
inline std::map<std::string, ParamValue>& param_table() { static std::map<std::string, ParamValue> t; return t; }
struct Time {
  double sec = 0;
  static Time now() { return Time(); }
  double toSec() const { return sec; }
};
struct Duration { Duration(double = 0) {} };
class Publisher {
 public:
  template <typename M> void publish(const M&) const {}
  int getNumSubscribers() const { return 0; }
};
class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  bool ok() const { return true; }
  template <typename M> Publisher advertise(const std::string&, int, bool = false) { return Publisher(); }
  template <typename T> bool param(const std::string& name, T& out, const T& dflt) const {
    auto it = param_table().find(name);
    if (it == param_table().end()) { out = dflt; return false; }
    assign(out, it->second);
    return true;
  }
 private:
  static void assign(bool& o, const ParamValue& v) { o = v.num != 0; }
  static void assign(int& o, const ParamValue& v) { o = (int)v.num; }
  static void assign(double& o, const ParamValue& v) { o = v.num; }
  static void assign(float& o, const ParamValue& v) { o = (float)v.num; }
  static void assign(std::string& o, const ParamValue& v) { o = v.str; }
  static void assign(std::vector<double>& o, const ParamValue& v) { o = v.vec; }
  static void assign(std::vector<float>& o, const ParamValue& v) { o.assign(v.vec.begin(), v.vec.end()); }
};
inline void spin() {}
inline void spinOnce() {}
inline void init(int&, char**, const std::string&) {}
}  // namespace ros
#define ROS_INFO(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_WARN(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_ERROR(...) do { std::printf(__VA_ARGS__); std::printf("\n"); } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { } while (0)
