// Stand-in for the PCL point types the reference's hot-path headers name (NOT PCL; test infrastructure).
// Only the fields include/BALM/tools.hpp, include/BALM/bavoxel.hpp and include/utils.hpp touch exist.
#pragma once
#include <cstdint>
namespace pcl {
struct PointXYZINormal {
  union { float data[4]; struct { float x, y, z; }; };
  union { float data_n[4]; struct { float normal_x, normal_y, normal_z; }; };
  float intensity, curvature;
  PointXYZINormal() : data{0, 0, 0, 1}, data_n{0, 0, 0, 0}, intensity(0), curvature(0) {}
};
struct PointXYZRGB {
  union { float data[4]; struct { float x, y, z; }; };
  uint8_t b, g, r, a;
  PointXYZRGB() : data{0, 0, 0, 1}, b(0), g(0), r(0), a(255) {}
};
struct PointXYZRGBA : PointXYZRGB {};
struct PointXYZI {
  union { float data[4]; struct { float x, y, z; }; };
  float intensity;
  PointXYZI() : data{0, 0, 0, 1}, intensity(0) {}
};
struct PointXYZ {
  union { float data[4]; struct { float x, y, z; }; };
  PointXYZ() : data{0, 0, 0, 1} {}
  PointXYZ(float a, float b, float c) : data{a, b, c, 1} {}
};
}  // namespace pcl
