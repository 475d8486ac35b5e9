// STAND-IN for the OpenCV names include/utils.hpp and src/lvba_system.cpp mention (NOT OpenCV; test infrastructure, see
// ../mini_eigen.h).  cv::Mat is a real (tiny) single-channel float / uint16 image — what generateDepthWithVoxel writes and
// fetchDepthBilinear (include/utils.hpp:246-274) reads — so those run from their own source; image decoding, drawing and the
// image-processing calls (visualisation, feature extraction: out of scope) are declared, do nothing or abort if reached.
#pragma once
#include <array>
#include <cmath>
#include <fstream>
#include <set>
#include <cstdint>
#include <cstdlib>
#include <cstring>
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
template <typename T> struct Point_ { T x, y; Point_() : x(0), y(0) {} template <typename A, typename B> Point_(A a, B b) : x((T)a), y((T)b) {}
  template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
  Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); } Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); } };
typedef Point_<int> Point; typedef Point_<float> Point2f; typedef Point_<double> Point2d;
template <typename T> double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} static Scalar all(double a) { return Scalar(a, a, a, a); } };
struct Vec3b { unsigned char v[3]; Vec3b() : v{0, 0, 0} {} Vec3b(unsigned char a, unsigned char b, unsigned char c) : v{a, b, c} {} unsigned char& operator[](int i) { return v[i]; } const unsigned char& operator[](int i) const { return v[i]; } };
struct Rect { int x, y, width, height; Rect(int a = 0, int b = 0, int c = 0, int d = 0) : x(a), y(b), width(c), height(d) {} };
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)
#define CV_8UC3 16
#define CV_8UC1 0
#define CV_64F 6
#define CV_32F 5
#define CV_32FC2 13
#define CV_16SC2 11
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_(std::make_shared<std::vector<uint8_t>>((size_t)r * c * elem(type))) {}
  Mat(int r, int c, int type, const Scalar& s) : Mat(r, c, type) { fill(s.v[0]); }
  Mat(Size sz, int type, const Scalar& s = Scalar()) : Mat(sz.height, sz.width, type) { fill(s.v[0]); }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  static Mat zeros(Size sz, int type) { return Mat(sz.height, sz.width, type); }
  static int elem(int type) { return type == CV_32FC1 ? 4 : type == CV_16UC1 ? 2 : type == CV_64F ? 8 : type == CV_8UC3 ? 3 : type == CV_8UC1 ? 1 : 8; }
  void fill(double v) {
    if (type_ == CV_32FC1) for (size_t i = 0; i < (size_t)rows * cols; ++i) reinterpret_cast<float*>(buf_->data())[i] = (float)v;
    else if (type_ == CV_8UC3 || type_ == CV_8UC1) std::memset(buf_->data(), (int)v, buf_->size());
  }
  Mat& setTo(const Scalar& s) { fill(s.v[0]); return *this; }
  unsigned char* data = nullptr;          // only handed to SiftGPU (out of scope)
  Size size() const { return Size(cols, rows); }
  int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
  void copyTo(Mat& o) const { o = clone(); }
  void copyTo(Mat&& o) const { o = clone(); }
  void convertTo(Mat& o, int, double = 1, double = 0) const { o = clone(); }
  Mat operator()(const Rect&) const { std::abort(); }
  template <typename T> T* ptr(int y) { return reinterpret_cast<T*>(buf_->data()) + (size_t)y * cols; }
  template <typename T> const T* ptr(int y) const { return reinterpret_cast<const T*>(buf_->data()) + (size_t)y * cols; }
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
template <typename T> class Mat_ : public Mat { public: Mat_() {} Mat_(int r, int c) : Mat(r, c, CV_64F) {}
  static Mat_ eye(int r, int c) { return Mat_(r, c); }
  T& operator()(int y, int x) { return this->template at<T>(y, x); }
  template <typename S> Mat_& operator<<(S) { return *this; } template <typename S> Mat_& operator,(S) { return *this; } };
enum { INTER_LINEAR = 1, INTER_CUBIC = 2, COLOR_BGR2Lab = 44, COLOR_Lab2BGR = 56, COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8, IMREAD_COLOR = 1, IMREAD_GRAYSCALE = 0, IMREAD_UNCHANGED = -1,
       LINE_AA = 16, LINE_8 = 8, FILLED = -1, FONT_HERSHEY_SIMPLEX = 0, BORDER_CONSTANT = 0 };
struct RNG { explicit RNG(unsigned long long = 0) {} int uniform(int a, int) { return a; } double uniform(double a, double) { return a; } };
// image files, drawing, remapping: visualisation / export only — nothing on the path reads what they produce
// imread: no decoder here.  The tests of the COLMAP export (VisualizeOptComparison, src/lvba_system.cpp:1932-2143) ask for "an image of this size in
// one colour" through stub_image(): every file then reads as rows x cols pixels of (grey, grey, grey); by default files read as empty (as before).
struct StubImage { int rows = 0, cols = 0, grey = 128; };
inline StubImage& stub_image() { static StubImage s; return s; }
inline Mat imread(const std::string&, int = IMREAD_COLOR) {
  const StubImage& s = stub_image();
  return s.rows > 0 ? Mat(s.rows, s.cols, CV_8UC3, Scalar(s.grey)) : Mat();
}
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline void circle(Mat&, Point, int, const Scalar&, int = 1, int = LINE_8, int = 0) {}
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = LINE_8, int = 0) {}
inline void rectangle(Mat&, Point, Point, const Scalar&, int = 1, int = LINE_8, int = 0) {}
inline void rectangle(Mat&, Rect, const Scalar&, int = 1, int = LINE_8, int = 0) {}
inline void putText(Mat&, const std::string&, Point, int, double, const Scalar&, int = 1, int = LINE_8, bool = false) {}
inline void hconcat(const Mat&, const Mat&, Mat&) {}
inline void vconcat(const Mat&, const Mat&, Mat&) {}
inline void remap(const Mat& src, Mat& dst, const Mat&, const Mat&, int, int = BORDER_CONSTANT, const Scalar& = Scalar()) { dst = src.clone(); }   // undistortion of an image that is only written to disk
inline void initUndistortRectifyMap(const Mat&, const Mat&, const Mat&, const Mat&, Size, int, Mat&, Mat&) {}
[[noreturn]] inline void out_of_scope() { std::abort(); }
inline void resize(const Mat&, Mat&, Size, double = 0, double = 0, int = INTER_LINEAR) { out_of_scope(); }
inline void cvtColor(const Mat&, Mat&, int) { out_of_scope(); }
inline void split(const Mat&, std::vector<Mat>&) { out_of_scope(); }
inline void merge(const std::vector<Mat>&, Mat&) { out_of_scope(); }
inline Ptr<CLAHE> createCLAHE(double, Size) { out_of_scope(); }
inline void GaussianBlur(const Mat&, Mat&, Size, double) { out_of_scope(); }
inline void addWeighted(const Mat&, double, const Mat&, double, double, Mat&) { out_of_scope(); }
}  // namespace cv
