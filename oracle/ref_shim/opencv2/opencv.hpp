// STAND-IN for the OpenCV names include/utils.hpp mentions (NOT OpenCV; test infrastructure, see ../mini_eigen.h).
// cv::Mat is a real (tiny) single-channel image so that fetchDepthBilinear (include/utils.hpp:246-274) can be run from its own
// source; the image-processing calls of preprocessLowTextureBGR (:426-446, feature extraction, out of scope) are declared
// and abort if reached.
#pragma once
#include <array>
#include <cstdint>
#include <cstdlib>
#include <iomanip>
#include <iostream>
#include <memory>
#include <regex>
#include <sstream>
#include <string>
#include <vector>
#define CV_32FC1 5
#define CV_16UC1 2
namespace cv {
struct Size { int width, height; Size(int w = 0, int h = 0) : width(w), height(h) {} };
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_(std::make_shared<std::vector<uint8_t>>((size_t)r * c * (type == CV_32FC1 ? 4 : 2))) {}
  bool empty() const { return rows == 0 || cols == 0; }
  int type() const { return type_; }
  template <typename T> T& at(int y, int x) { return reinterpret_cast<T*>(buf_->data())[(size_t)y * cols + x]; }
  template <typename T> const T& at(int y, int x) const { return reinterpret_cast<const T*>(buf_->data())[(size_t)y * cols + x]; }
  Mat clone() const { Mat m(*this); if (buf_) m.buf_ = std::make_shared<std::vector<uint8_t>>(*buf_); return m; }
 private:
  int type_ = CV_32FC1;
  std::shared_ptr<std::vector<uint8_t>> buf_;
};
template <typename T> using Ptr = std::shared_ptr<T>;
struct CLAHE { void apply(const Mat&, Mat&) { std::abort(); } };
enum { INTER_CUBIC = 2, COLOR_BGR2Lab = 44, COLOR_Lab2BGR = 56 };
[[noreturn]] inline void out_of_scope() { std::abort(); }
inline void resize(const Mat&, Mat&, Size, double, double, int) { out_of_scope(); }
inline void cvtColor(const Mat&, Mat&, int) { out_of_scope(); }
inline void split(const Mat&, std::vector<Mat>&) { out_of_scope(); }
inline void merge(const std::vector<Mat>&, Mat&) { out_of_scope(); }
inline Ptr<CLAHE> createCLAHE(double, Size) { out_of_scope(); }
inline void GaussianBlur(const Mat&, Mat&, Size, double) { out_of_scope(); }
inline void addWeighted(const Mat&, double, const Mat&, double, double, Mat&) { out_of_scope(); }
}  // namespace cv
