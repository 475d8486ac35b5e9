// Stand-in (NOT PCL; test infrastructure).  loadPCDFile reads the two uncompressed encodings of the PCD format (DATA ascii / DATA binary, float32
// fields x y z [intensity] at the offsets the header declares) — enough for src/dataset_io.cpp to run on the files oracle/dataset_writer.py writes;
// binary_compressed returns -1 (the reference then skips the file, :270-273).  PCL's own parser is NOT what runs here.  The save functions
// (visualisation / export of the reference) do nothing.
#pragma once
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "../point_cloud.h"
#include "../point_types.h"
namespace pcl { namespace io {
namespace detail {
inline void put_intensity(PointXYZI& p, float v) { p.intensity = v; }
inline void put_intensity(PointXYZINormal& p, float v) { p.intensity = v; }
template <typename P> void put_intensity(P&, float) {}
}  // namespace detail
template <typename P> int loadPCDFile(const std::string& path, pcl::PointCloud<P>& out) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return -1;
  std::vector<std::string> fields; std::vector<int> size, count; std::vector<char> type;
  long long points = -1; std::string data, line;
  while (std::getline(f, line)) {
    std::istringstream is(line); std::string key; is >> key;
    if (key == "FIELDS") { std::string v; while (is >> v) fields.push_back(v); }
    else if (key == "SIZE") { int v; while (is >> v) size.push_back(v); }
    else if (key == "TYPE") { char v; while (is >> v) type.push_back(v); }
    else if (key == "COUNT") { int v; while (is >> v) count.push_back(v); }
    else if (key == "POINTS") is >> points;
    else if (key == "DATA") { is >> data; break; }
  }
  const size_t nf = fields.size();
  if (nf == 0 || size.size() != nf || type.size() != nf || points < 0) return -1;
  if (count.empty()) count.assign(nf, 1);
  std::vector<size_t> off(nf); size_t rec = 0;
  for (size_t k = 0; k < nf; ++k) { off[k] = rec; rec += (size_t)size[k] * count[k]; }
  int ix = -1, iy = -1, iz = -1, ii = -1;
  for (size_t k = 0; k < nf; ++k) { if (fields[k] == "x") ix = (int)k; if (fields[k] == "y") iy = (int)k; if (fields[k] == "z") iz = (int)k; if (fields[k] == "intensity") ii = (int)k; }
  if (ix < 0 || iy < 0 || iz < 0) return -1;
  out.clear();
  if (data == "ascii") {
    for (long long i = 0; i < points && std::getline(f, line); ++i) {
      std::istringstream is(line); std::vector<double> v; double x;
      while (is >> x) v.push_back(x);
      std::vector<size_t> col(nf); size_t c = 0; for (size_t k = 0; k < nf; ++k) { col[k] = c; c += count[k]; }
      if (v.size() < c) return -1;
      P p; p.x = (float)v[col[ix]]; p.y = (float)v[col[iy]]; p.z = (float)v[col[iz]];
      if (ii >= 0) detail::put_intensity(p, (float)v[col[ii]]);
      out.push_back(p);
    }
    return 0;
  }
  if (data != "binary") return -1;
  if (size[ix] != 4 || size[iy] != 4 || size[iz] != 4) return -1;
  std::vector<char> buf(rec);
  for (long long i = 0; i < points; ++i) {
    if (!f.read(buf.data(), (std::streamsize)rec)) return -1;
    P p; float v;
    std::memcpy(&v, buf.data() + off[ix], 4); p.x = v; std::memcpy(&v, buf.data() + off[iy], 4); p.y = v; std::memcpy(&v, buf.data() + off[iz], 4); p.z = v;
    if (ii >= 0 && size[ii] == 4) { std::memcpy(&v, buf.data() + off[ii], 4); detail::put_intensity(p, v); }
    out.push_back(p);
  }
  return 0;
}
template <typename P> int savePCDFileBinary(const std::string&, const pcl::PointCloud<P>&) { return 0; }
template <typename P> int savePCDFile(const std::string&, const pcl::PointCloud<P>&) { return 0; }
template <typename P> int savePLYFileBinary(const std::string&, const pcl::PointCloud<P>&) { return 0; }
} }
