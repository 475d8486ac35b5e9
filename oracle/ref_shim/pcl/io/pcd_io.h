// Stand-in (NOT PCL; test infrastructure).  loadPCDFile reads the three encodings of the PCD format (DATA ascii / binary / binary_compressed = LZF,
// float32 fields x y z [intensity] at the offsets the header declares) — enough for src/dataset_io.cpp to run on the files oracle/dataset_writer.py writes.  PCL's own parser is NOT what runs here.  The save functions
// (visualisation / export of the reference) do nothing.
#pragma once
#include <cstdint>
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
  if (data == "binary_compressed") {
    // PCD's third encoding: uint32 compressed size, uint32 uncompressed size, LZF stream; the uncompressed block is field by field (all x, all y, ...)
    uint32_t csz = 0, usz = 0;
    if (!f.read(reinterpret_cast<char*>(&csz), 4) || !f.read(reinterpret_cast<char*>(&usz), 4)) return -1;
    std::vector<unsigned char> in(csz), raw(usz);
    if (csz && !f.read(reinterpret_cast<char*>(in.data()), csz)) return -1;
    size_t ip = 0, op = 0;
    while (ip < in.size()) {                                 // liblzf's published format: literal runs (ctrl < 32) and back references
      const unsigned ctrl = in[ip++];
      if (ctrl < 32) {
        const size_t len = ctrl + 1;
        if (ip + len > in.size() || op + len > raw.size()) return -1;
        std::memcpy(&raw[op], &in[ip], len); ip += len; op += len;
      } else {
        size_t len = ctrl >> 5;
        if (len == 7) { if (ip >= in.size()) return -1; len += in[ip++]; }
        if (ip >= in.size()) return -1;
        const size_t back = ((size_t)(ctrl & 0x1f) << 8) + in[ip++] + 1;
        len += 2;
        if (back > op || op + len > raw.size()) return -1;
        for (size_t k = 0; k < len; ++k, ++op) raw[op] = raw[op - back];
      }
    }
    if (op != raw.size() || (size_t)points * rec != raw.size()) return -1;
    if (size[ix] != 4 || size[iy] != 4 || size[iz] != 4) return -1;
    std::vector<size_t> col(nf); size_t acc = 0;
    for (size_t k = 0; k < nf; ++k) { col[k] = acc; acc += (size_t)size[k] * count[k] * (size_t)points; }
    for (long long i = 0; i < points; ++i) {
      P p; float v;
      std::memcpy(&v, &raw[col[ix] + 4 * (size_t)i], 4); p.x = v; std::memcpy(&v, &raw[col[iy] + 4 * (size_t)i], 4); p.y = v; std::memcpy(&v, &raw[col[iz] + 4 * (size_t)i], 4); p.z = v;
      if (ii >= 0 && size[ii] == 4) { std::memcpy(&v, &raw[col[ii] + 4 * (size_t)i], 4); detail::put_intensity(p, v); }
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
