// Stand-in for the old (non-templated) Sophus::SE3 include/utils.hpp:480-492 names (NOT Sophus; test infrastructure).
#pragma once
#include "../mini_eigen.h"
namespace Sophus {
class SE3 {
  Eigen::Matrix3d R_;
  Eigen::Vector3d t_;
 public:
  SE3() : R_(Eigen::Matrix3d::Identity()) {}
  SE3(const Eigen::Matrix3d& R, const Eigen::Vector3d& t) : R_(R), t_(t) {}
  SE3(const Eigen::Quaterniond& q, const Eigen::Vector3d& t) : R_(q.toRotationMatrix()), t_(t) {}
  Eigen::Matrix3d rotation_matrix() const { return R_; }
  const Eigen::Vector3d& translation() const { return t_; }
  Eigen::Vector3d& translation() { return t_; }
  SE3 operator*(const SE3& o) const { return SE3(R_ * o.R_, R_ * o.t_ + t_); }          // composition, as Sophus defines it
  Eigen::Vector3d operator*(const Eigen::Vector3d& p) const { return R_ * p + t_; }
  SE3 inverse() const { const Eigen::Matrix3d Rt = R_.transpose(); return SE3(Rt, -(Rt * t_)); }
};
}  // namespace Sophus
