// Stand-in for pcl::PointCloud (NOT PCL; test infrastructure): a vector of points with the members the reference's
// hot-path headers use.
#pragma once
#include <cctype>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>
namespace pcl {
template <typename PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void reserve(size_t n) { points.reserve(n); }
  void clear() { points.clear(); width = height = 0; }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  void swap(PointCloud& o) { points.swap(o.points); std::swap(width, o.width); std::swap(height, o.height); std::swap(is_dense, o.is_dense); }
  PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (uint32_t)points.size(); height = 1; return *this; }
  template <typename... A> void emplace_back(A&&... a) { points.emplace_back(std::forward<A>(a)...); width = (uint32_t)points.size(); height = 1; }
  void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
};
}  // namespace pcl
