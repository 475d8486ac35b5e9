// lvba_dataset.hpp — reader / writer for the reference's on-disk dataset layout, without ROS, PCL, Sophus or OpenCV
// (SURVEY.md §8f N4).  Header only, C++17.  Mirrors src/dataset_io.cpp of xuankuzcr/Global-LVBA:
//
//   <data>/all_pcd_body/<timestamp>.pcd      body-frame scans (pcl::PointXYZI), sorted by the number in the file name
//                                            (handleBodyPoints :199-260, parseTimestampFromName include/utils.hpp:462-477)
//   <data>/all_pcd_body/lidar_poses.txt      TUM lines `t tx ty tz qx qy qz qw`, one per scan      (handleLidarPoses :182-191)
//   <data>/all_image/image_poses.txt         TUM, every image_stride-th valid line                  (handleCamPoses :193-197)
//   loadPosesTUM :129-180                    '#' and empty lines skipped, unparsable lines skipped with a warning, the stride
//                                            counts VALID lines, quaternion normalised before use
//
// PCD: the three DATA encodings pcl::io::loadPCDFile accepts (ascii, binary, binary_compressed = LZF over the field-major
// buffer); only x, y, z (float32) are kept — the only members of PointType the BA path reads.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <regex>
#include <sstream>
#include <limits>
#include <string>
#include <vector>

namespace lvba_b200 {
namespace dataset {

// IMUST subset with the accessors the templates of lvba_shim.hpp use (R(r, c), p(r), t)
struct Mat3 { double m[9]; double& operator()(int r, int c) { return m[3 * r + c]; } double operator()(int r, int c) const { return m[3 * r + c]; } };
struct Vec3 { double v[3]; double& operator()(int r) { return v[r]; } double operator()(int r) const { return v[r]; } };
struct Pose {
  Mat3 R;
  Vec3 p;
  double t = 0.0;
};

struct Point { float x, y, z; };
struct Cloud { std::vector<Point> points; };

// first "digits[.digits]" in the name, as std::stod reads it   (include/utils.hpp:462-477)
inline bool parse_timestamp_from_name(const std::string& name, double& ts) {
  static const std::regex re(R"(([0-9]+(?:\.[0-9]+)?))");
  std::smatch m;
  if (!std::regex_search(name, m, re)) return false;
  try { ts = std::stod(m[1].str()); } catch (...) { return false; }
  return true;
}

// Eigen::Quaterniond(qw, qx, qy, qz).normalize() -> rotation matrix (what Sophus::SE3(q, t).rotation_matrix() returns)
inline void quat_to_R(double qw, double qx, double qy, double qz, double R[9]) {
  const double n = std::sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  qw /= n; qx /= n; qy /= n; qz /= n;
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// rotation matrix -> unit quaternion (w >= 0), for the TUM writer
inline void R_to_quat(const double R[9], double q[4] /* w x y z */) {
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) {
    const double s = std::sqrt(tr + 1.0) * 2;
    q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s;
  } else if (R[0] > R[4] && R[0] > R[8]) {
    const double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
    q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s;
  } else if (R[4] > R[8]) {
    const double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
    q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s;
  } else {
    const double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
    q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s;
  }
  if (q[0] < 0) for (int k = 0; k < 4; ++k) q[k] = -q[k];
}

// loadPosesTUM (src/dataset_io.cpp:129-180).  Returns false when the file cannot be opened or holds no pose.
inline bool load_poses_tum(const std::string& file, size_t stride, std::vector<Pose>& out, std::string* err = nullptr,
                           size_t* parsed = nullptr, size_t* skipped = nullptr) {
  out.clear();
  std::ifstream fin(file);
  if (!fin.is_open()) { if (err) *err = "cannot open " + file; return false; }
  if (stride == 0) { if (err) *err = "stride 0"; return false; }
  std::string line;
  size_t valid = 0, bad = 0;
  while (std::getline(fin, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream iss(line);
    double ts, tx, ty, tz, qx, qy, qz, qw;
    if (!(iss >> ts >> tx >> ty >> tz >> qx >> qy >> qz >> qw)) { ++bad; continue; }
    if (valid % stride == 0) {
      Pose p;
      quat_to_R(qw, qx, qy, qz, p.R.m);
      p.p.v[0] = tx; p.p.v[1] = ty; p.p.v[2] = tz; p.t = ts;
      out.push_back(p);
    }
    ++valid;
  }
  if (parsed) *parsed = valid;
  if (skipped) *skipped = bad;
  if (out.empty()) { if (err) *err = "no poses in " + file; return false; }
  return true;
}

inline bool save_poses_tum(const std::string& file, const std::vector<Pose>& poses) {
  FILE* f = std::fopen(file.c_str(), "w");
  if (!f) return false;
  for (const Pose& p : poses) {
    double q[4];
    R_to_quat(p.R.m, q);
    std::fprintf(f, "%.9f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n", p.t, p.p.v[0], p.p.v[1], p.p.v[2], q[1], q[2], q[3], q[0]);
  }
  return std::fclose(f) == 0;
}

// ---------------------------------------------------------------- LZF (the codec of PCD binary_compressed)
// Format (liblzf): control byte c < 32 -> c + 1 literal bytes follow; otherwise a back reference of length (c >> 5) + 2
// (a length field of 7 is extended by the next byte) at distance ((c & 31) << 8 | next) + 1.
inline bool lzf_decompress(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len) {
  size_t ip = 0, op = 0;
  while (ip < in_len) {
    unsigned ctrl = in[ip++];
    if (ctrl < 32) {
      const size_t n = ctrl + 1;
      if (ip + n > in_len || op + n > out_len) return false;
      std::memcpy(out + op, in + ip, n);
      ip += n; op += n;
    } else {
      size_t len = ctrl >> 5;
      if (len == 7) { if (ip >= in_len) return false; len += in[ip++]; }
      if (ip >= in_len) return false;
      const size_t dist = ((size_t)(ctrl & 31) << 8 | in[ip++]) + 1;
      len += 2;
      if (dist > op || op + len > out_len) return false;
      for (size_t k = 0; k < len; ++k, ++op) out[op] = out[op - dist];          // may overlap: byte by byte
    }
  }
  return op == out_len;
}

// ---------------------------------------------------------------- PCD
struct PcdField { std::string name; int size = 4; char type = 'F'; int count = 1; size_t offset = 0; };

inline bool load_pcd_xyz(const std::string& file, Cloud& cloud, std::string* err = nullptr) {
  cloud.points.clear();
  auto fail = [&](const std::string& m) { if (err) *err = file + ": " + m; return false; };
  std::ifstream fin(file, std::ios::binary);
  if (!fin.is_open()) return fail("cannot open");
  std::vector<PcdField> fields;
  size_t n_points = 0, width = 0, height = 1;
  std::string data;
  std::string line;
  while (std::getline(fin, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream iss(line);
    std::string key;
    iss >> key;
    if (key == "FIELDS") { std::string f; while (iss >> f) { PcdField pf; pf.name = f; fields.push_back(pf); } }
    else if (key == "SIZE") { for (auto& f : fields) iss >> f.size; }
    else if (key == "TYPE") { for (auto& f : fields) iss >> f.type; }
    else if (key == "COUNT") { for (auto& f : fields) iss >> f.count; }
    else if (key == "WIDTH") iss >> width;
    else if (key == "HEIGHT") iss >> height;
    else if (key == "POINTS") iss >> n_points;
    else if (key == "DATA") { iss >> data; break; }
  }
  if (data.empty() || fields.empty()) return fail("not a PCD header");
  // the header is untrusted input: sizes / counts / point counts are checked before they size a buffer or an offset
  for (const PcdField& f : fields) {
    if (f.size != 1 && f.size != 2 && f.size != 4 && f.size != 8) return fail("field '" + f.name + "': SIZE must be 1, 2, 4 or 8");
    if (f.count < 1 || f.count > (1 << 20)) return fail("field '" + f.name + "': COUNT out of range");
  }
  if (n_points == 0) {
    if (height != 0 && width > std::numeric_limits<size_t>::max() / height) return fail("WIDTH x HEIGHT overflows");
    n_points = width * height;
  }
  size_t rec = 0;
  int ix = -1, iy = -1, iz = -1;
  for (size_t k = 0; k < fields.size(); ++k) {
    fields[k].offset = rec;
    rec += (size_t)fields[k].size * (size_t)fields[k].count;
    if (rec > (size_t(1) << 30)) return fail("record size out of range");
    if (fields[k].name == "x") ix = (int)k; else if (fields[k].name == "y") iy = (int)k; else if (fields[k].name == "z") iz = (int)k;
  }
  if (ix < 0 || iy < 0 || iz < 0) return fail("no x / y / z fields");
  for (int k : {ix, iy, iz})
    if (fields[k].size != 4 || fields[k].type != 'F' || fields[k].count != 1 || fields[k].offset + 4 > rec) return fail("x / y / z must be float32");
  // plausibility against what is left of the file: an ascii point needs >= 2 bytes per field, a binary one `rec` bytes, a
  // compressed stream cannot expand more than ~256x (LZF back references)
  {
    const std::streampos here = fin.tellg();
    fin.seekg(0, std::ios::end);
    const std::streampos end = fin.tellg();
    fin.seekg(here);
    const size_t left = (here >= 0 && end >= here) ? (size_t)(end - here) : 0;
    if (rec != 0 && n_points > std::numeric_limits<size_t>::max() / rec) return fail("POINTS x record size overflows");
    const size_t need = data == "ascii" ? n_points * 2 : data == "binary" ? n_points * rec : (n_points * rec) / 256;
    if (need > left) return fail("POINTS (" + std::to_string(n_points) + ") does not fit the file");
  }
  try {
    cloud.points.resize(n_points);
  } catch (const std::exception&) {
    cloud.points.clear();
    return fail("cannot allocate " + std::to_string(n_points) + " points");
  }
  if (data == "ascii") {
    for (size_t i = 0; i < n_points; ++i) {
      if (!std::getline(fin, line)) return fail("truncated ascii data");
      std::istringstream iss(line);
      size_t col = 0;
      for (size_t k = 0; k < fields.size(); ++k)
        for (int c = 0; c < fields[k].count; ++c, ++col) {
          double v;
          if (!(iss >> v)) return fail("bad ascii record");
          if ((int)k == ix) cloud.points[i].x = (float)v; else if ((int)k == iy) cloud.points[i].y = (float)v; else if ((int)k == iz) cloud.points[i].z = (float)v;
        }
    }
    return true;
  }
  if (data == "binary") {
    std::vector<uint8_t> buf;
    try { buf.resize(rec * n_points); } catch (const std::exception&) { return fail("cannot allocate the binary records"); }
    fin.read((char*)buf.data(), (std::streamsize)buf.size());
    if ((size_t)fin.gcount() != buf.size()) return fail("truncated binary data");
    for (size_t i = 0; i < n_points; ++i) {
      const uint8_t* r = buf.data() + i * rec;
      std::memcpy(&cloud.points[i].x, r + fields[ix].offset, 4);
      std::memcpy(&cloud.points[i].y, r + fields[iy].offset, 4);
      std::memcpy(&cloud.points[i].z, r + fields[iz].offset, 4);
    }
    return true;
  }
  if (data == "binary_compressed") {
    uint32_t csize = 0, usize = 0;
    fin.read((char*)&csize, 4); fin.read((char*)&usize, 4);
    if (!fin || (size_t)usize != rec * n_points) return fail("bad compressed header");
    std::vector<uint8_t> comp, buf;
    try { comp.resize(csize); buf.resize(usize); } catch (const std::exception&) { return fail("cannot allocate the compressed stream"); }
    fin.read((char*)comp.data(), csize);
    if ((size_t)fin.gcount() != csize) return fail("truncated compressed data");
    if (!lzf_decompress(comp.data(), csize, buf.data(), usize)) return fail("LZF stream corrupt");
    // field-major: all values of field 0, then field 1, ...
    size_t base = 0;
    for (size_t k = 0; k < fields.size(); ++k) {
      const size_t stride = (size_t)fields[k].size * fields[k].count;
      if ((int)k == ix || (int)k == iy || (int)k == iz)
        for (size_t i = 0; i < n_points; ++i) {
          float v;
          std::memcpy(&v, buf.data() + base + i * stride, 4);
          if ((int)k == ix) cloud.points[i].x = v; else if ((int)k == iy) cloud.points[i].y = v; else cloud.points[i].z = v;
        }
      base += stride * n_points;
    }
    return true;
  }
  return fail("unknown DATA encoding '" + data + "'");
}

// ---------------------------------------------------------------- the dataset (handleLidarPoses + handleBodyPoints)
struct LidarDataset {
  std::vector<Pose> x_buf;            // pose of every scan with its timestamp (dataset_io_->x_buf_)
  std::vector<Cloud> clouds;          // dataset_io_->pl_fulls_
  std::vector<Cloud*> pl_fulls;       // pointer view (what lvba_shim.hpp's templates take)
  std::vector<std::string> warnings;
  size_t size() const { return std::min(x_buf.size(), clouds.size()); }
};

inline bool load_lidar_dataset(const std::string& root, LidarDataset& ds, std::string* err = nullptr) {
  namespace fs = std::filesystem;
  ds = LidarDataset();
  const fs::path dir = fs::path(root) / "all_pcd_body";
  if (!fs::exists(dir) || !fs::is_directory(dir)) { if (err) *err = "pcd dir missing: " + dir.string(); return false; }
  std::vector<Pose> lidar_poses;
  size_t bad = 0;
  if (!load_poses_tum((dir / "lidar_poses.txt").string(), 1, lidar_poses, err, nullptr, &bad)) return false;
  if (bad) ds.warnings.push_back(std::to_string(bad) + " unparsable line(s) in lidar_poses.txt");
  std::vector<std::pair<double, std::string>> pcds;
  for (const auto& e : fs::directory_iterator(dir)) {
    if (!e.is_regular_file() || e.path().extension() != ".pcd") continue;
    double ts = 0.0;
    if (!parse_timestamp_from_name(e.path().filename().string(), ts)) { ds.warnings.push_back("bad pcd name: " + e.path().string()); continue; }
    pcds.emplace_back(ts, e.path().string());
  }
  if (pcds.empty()) { if (err) *err = "no pcd files in " + dir.string(); return false; }
  std::sort(pcds.begin(), pcds.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
  if (lidar_poses.size() > pcds.size()) { if (err) *err = "more poses than pcd files"; return false; }   // the reference would index past pcd_ts
  for (size_t m = 0; m < lidar_poses.size(); ++m) { Pose p = lidar_poses[m]; p.t = pcds[m].first; ds.x_buf.push_back(p); }   // :228-233
  for (const auto& kv : pcds) {
    Cloud c;
    std::string e2;
    if (!load_pcd_xyz(kv.second, c, &e2)) { ds.warnings.push_back(e2); continue; }                      // :243-246: failed loads are skipped
    ds.clouds.push_back(std::move(c));
  }
  for (auto& c : ds.clouds) ds.pl_fulls.push_back(&c);
  return true;
}

}  // namespace dataset
}  // namespace lvba_b200
