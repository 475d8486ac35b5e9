// runtime.cuh — host-side plumbing shared by both paths: error reporting, device buffers,
// CUDA-event phase timers, the block-envelope structure builder and the LDL^T solve driver.
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lvba_b200.h"
#include "envelope.cuh"

namespace lvba {

// ---------------------------------------------------------------- errors
inline std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}
#define LVBA_CUDA(call)                                                                        \
  do {                                                                                         \
    cudaError_t err__ = (call);                                                                \
    if (err__ != cudaSuccess)                                                                  \
      return ::lvba::fail(LVBA_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,  \
                          cudaGetErrorString(err__));                                          \
  } while (0)
#define LVBA_TRY(call)             \
  do {                             \
    int rc__ = (call);             \
    if (rc__ != LVBA_OK) return rc__; \
  } while (0)

inline int device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
inline int select_device(int device) {
  const int n = device_count();
  if (n <= 0) return fail(LVBA_ERR_NO_DEVICE, "no CUDA device available: the LVBA hot path has no CPU fallback");
  if (device >= n) return fail(LVBA_ERR_INVALID_ARG, "device %d out of range (have %d)", device, n);
  if (device >= 0) LVBA_CUDA(cudaSetDevice(device));
  return LVBA_OK;
}

// ---------------------------------------------------------------- device memory pool
// cudaMalloc / cudaFree cost milliseconds each (cudaFree also synchronises the device); the one-shot ABI calls
// create and destroy ~40 buffers per call.  Freed buffers are therefore parked in a per-device, size-bucketed
// pool and reused by later calls of the same process (identical problem sizes hit exactly).  The pool is
// capped; lvba_release_cached_memory() (include/lvba_b200.h) empties it.
struct DevicePool {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> free_list;     // (device, bytes) -> pointer
  size_t cached_bytes = 0;
  static constexpr size_t kMaxCached = 16ull << 30;
  static size_t bucket(size_t bytes) { return (bytes + 511) & ~size_t(511); }
  void* take(int dev, size_t bytes) {
    std::lock_guard<std::mutex> g(mu);
    auto it = free_list.find({dev, bytes});
    if (it == free_list.end()) return nullptr;
    void* p = it->second;
    free_list.erase(it);
    cached_bytes -= bytes;
    return p;
  }
  void give(int dev, size_t bytes, void* p) {
    std::lock_guard<std::mutex> g(mu);
    if (cached_bytes + bytes > kMaxCached) { cudaFree(p); return; }
    free_list.insert({{dev, bytes}, p});
    cached_bytes += bytes;
  }
  void clear() {
    std::lock_guard<std::mutex> g(mu);
    for (auto& kv : free_list) cudaFree(kv.second);
    free_list.clear();
    cached_bytes = 0;
  }
};
inline DevicePool& device_pool() {
  static DevicePool* p = new DevicePool();     // intentionally leaked: no CUDA calls in static destructors
  return *p;
}

// ---------------------------------------------------------------- device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  size_t bytes_ = 0;
  int dev_ = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) device_pool().give(dev_, bytes_, p);
    p = nullptr; n = 0; bytes_ = 0;
  }
  int alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return LVBA_OK;
    bytes_ = DevicePool::bucket(count * sizeof(T));
    cudaGetDevice(&dev_);
    p = (T*)device_pool().take(dev_, bytes_);
    if (p) return LVBA_OK;
    cudaError_t e = cudaMalloc((void**)&p, bytes_);
    if (e != cudaSuccess) {
      cudaGetLastError();
      device_pool().clear();                                   // give cached memory back and retry once
      e = cudaMalloc((void**)&p, bytes_);
    }
    if (e != cudaSuccess) { p = nullptr; n = 0; bytes_ = 0; return fail(LVBA_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e)); }
    return LVBA_OK;
  }
  int upload(const T* h, size_t count, cudaStream_t s, int64_t* bytes = nullptr) {
    if (count > n) LVBA_TRY(alloc(count));
    if (count) LVBA_CUDA(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
    if (bytes) *bytes += (int64_t)(count * sizeof(T));
    return LVBA_OK;
  }
  int upload(const std::vector<T>& h, cudaStream_t s, int64_t* bytes = nullptr) { return upload(h.data(), h.size(), s, bytes); }
  int zero(cudaStream_t s) {
    if (n) LVBA_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s));
    return LVBA_OK;
  }
};

// ---------------------------------------------------------------- phase timers (CUDA events on the launch stream)
enum Phase { PH_BUILD = 0, PH_SOLVE = 1, PH_RESID = 2, PH_COUNT = 3 };
struct PhaseTimers {
  struct Span { cudaEvent_t a, b; int ph; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> pool;
  cudaStream_t stream = nullptr;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void begin(int ph) { Span s{get(), get(), ph}; cudaEventRecord(s.a, stream); spans.push_back(s); }
  void end() { cudaEventRecord(spans.back().b, stream); }
  // must be called after a stream synchronize
  void collect(double ms[PH_COUNT]) {
    for (auto& s : spans) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess) ms[s.ph] += t;
      pool.push_back(s.a); pool.push_back(s.b);
    }
    spans.clear();
  }
  ~PhaseTimers() {
    for (auto& s : spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : pool) cudaEventDestroy(e);
  }
};
inline double wall_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- block envelope (host build + device copy)
struct Envelope {
  int n = 0;
  std::vector<int> first, last;
  std::vector<long long> row_start;
  long long nblocks = 0;
  int max_col = 0;
  DevBuf<int> d_first, d_last;
  DevBuf<long long> d_row_start;

  // first_raw[r] = smallest column coupled to row r (<= r).  Made monotone so that the rows below a
  // pivot column form a contiguous range (see envelope.cuh header).
  int build(const std::vector<int>& first_raw, cudaStream_t s, int64_t* bytes) {
    n = (int)first_raw.size();
    first = first_raw;
    for (int r = 0; r < n; ++r) first[r] = std::min(first[r], r);
    for (int r = n - 2; r >= 0; --r) first[r] = std::min(first[r], first[r + 1]);
    row_start.assign(n + 1, 0);
    for (int r = 0; r < n; ++r) row_start[r + 1] = row_start[r] + (r - first[r] + 1);
    nblocks = row_start[n];
    last.assign(n, 0);
    // last[k] = max i with first[i] <= k ; first monotone => two-pointer sweep
    int i = 0;
    max_col = 0;
    for (int k = 0; k < n; ++k) {
      if (i < k) i = k;
      while (i + 1 < n && first[i + 1] <= k) ++i;
      last[k] = i;
      max_col = std::max(max_col, i - k);
    }
    LVBA_TRY(d_first.upload(first, s, bytes));
    LVBA_TRY(d_last.upload(last, s, bytes));
    LVBA_TRY(d_row_start.upload(row_start, s, bytes));
    return LVBA_OK;
  }
  EnvView view() const { return EnvView{n, d_first.p, d_row_start.p, d_last.p, nblocks}; }
};

// Thread -> slot-pair map of the register-window kernel.  Pairs {a,b}, a >= b, are grouped by 8x8 super
// blocks of the slot triangle so that the 32 lanes of a warp touch <= 8 distinct slots per operand: the
// shared-memory operand loads of a warp then need 1-2 wavefronts instead of 4-5.
inline std::vector<unsigned short> build_pair_map(int P, int n_threads) {
  std::vector<std::vector<unsigned short>> chunks;          // full 32-lane chunks first, leftovers after
  std::vector<std::vector<unsigned short>> left;
  const int nb = (P + 7) / 8;
  for (int A = 0; A < nb; ++A)
    for (int B = 0; B <= A; ++B) {
      std::vector<unsigned short> cur;
      for (int a = 8 * A; a < std::min(P, 8 * A + 8); ++a)
        for (int b = 8 * B; b < std::min(P, 8 * B + 8) && b <= a; ++b) {
          cur.push_back((unsigned short)(a | (b << 8)));
          if ((int)cur.size() == 32) { chunks.push_back(cur); cur.clear(); }
        }
      if (!cur.empty()) left.push_back(cur);
    }
  std::sort(left.begin(), left.end(), [](const auto& x, const auto& y) { return x.size() > y.size(); });
  const int n_warps = n_threads / 32;
  std::vector<std::vector<unsigned short>> warps = chunks;
  for (auto& l : left) {                                     // first-fit into a warp with room, else a new warp, else split
    bool placed = false;
    for (size_t w = chunks.size(); w < warps.size() && !placed; ++w)
      if (warps[w].size() + l.size() <= 32) { warps[w].insert(warps[w].end(), l.begin(), l.end()); placed = true; }
    if (!placed && (int)warps.size() < n_warps) { warps.push_back(l); placed = true; }
    if (!placed) {
      for (auto& w : warps)
        while (w.size() < 32 && !l.empty()) { w.push_back(l.back()); l.pop_back(); }
    }
  }
  std::vector<unsigned short> map((size_t)n_threads, (unsigned short)0xffff);
  for (size_t w = 0; w < warps.size() && (int)w < n_warps; ++w)
    for (size_t i = 0; i < warps[w].size(); ++i) map[w * 32 + i] = warps[w][i];
  return map;
}

// ---------------------------------------------------------------- LDL^T solve driver
struct EnvSolver {
  DevBuf<double> L, dinv, z;
  DevBuf<int> status;
  DevBuf<unsigned short> pair_map;   // thread -> slot pair of the register-window kernel (depends on P only)
  int map_P = 0;
  DevBuf<long long> dbg;        // LVBA_FACTOR_TIMING=1: per-step phase clocks of the register-window kernel
  int dbg_dumped = 0;
  bool configured = false;
  bool force_generic = false;   // tests: exercise the wide-envelope kernel on narrow problems
  int prepare(const Envelope& env) {
    if (env.max_col > kEnvMaxCol)
      return fail(LVBA_ERR_UNSUPPORTED, "envelope column height %d exceeds the factor kernel limit %d", env.max_col, kEnvMaxCol);
    {
      const char* fg = getenv("LVBA_FORCE_GENERIC_SOLVER");      // tests: run the wide-envelope kernel on narrow problems
      force_generic = fg && fg[0] == '1';
    }
    LVBA_TRY(L.alloc((size_t)env.nblocks * 36));
    LVBA_TRY(dinv.alloc((size_t)env.n * 36));
    LVBA_TRY(z.alloc((size_t)env.n * 6));
    LVBA_TRY(status.alloc(1));
    {
      const char* ft = getenv("LVBA_FACTOR_TIMING");
      if (ft && ft[0] == '1') { LVBA_TRY(dbg.alloc((size_t)env.n * 32)); }
    }
    if (!configured) {
      LVBA_CUDA(cudaFuncSetAttribute(env_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)factor_smem()));
      LVBA_CUDA(cudaFuncSetAttribute(env_factor_reg_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RegCfg<8>::kSmem));
      LVBA_CUDA(cudaFuncSetAttribute(env_factor_reg_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RegCfg<16>::kSmem));
      LVBA_CUDA(cudaFuncSetAttribute(env_factor_reg_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RegCfg<24>::kSmem));
      LVBA_CUDA(cudaFuncSetAttribute(env_factor_reg_kernel<31>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RegCfg<31>::kSmem));
      configured = true;
    }
    return LVBA_OK;
  }
  static size_t factor_smem() { return sizeof(double) * (2 * kEnvMaxCol * 36 + 36 + 8); }
  // Solves (H + diag(dadd)) x = z_in where z already holds the right-hand side.  x may alias nothing.
  int solve(const Envelope& env, const double* H, const double* dadd, double* x, cudaStream_t s, int64_t* launches) {
    const EnvView v = env.view();
    LVBA_CUDA(cudaMemcpyAsync(L.p, H, (size_t)env.nblocks * 36 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    LVBA_CUDA(cudaMemsetAsync(status.p, 0, sizeof(int), s));
    const int n6 = 6 * env.n;
    env_add_diag_kernel<<<(n6 + 255) / 256, 256, 0, s>>>(v, dadd, L.p);
    ++*launches;
    // column height < P: register-resident sliding-window kernel (envelope.cuh v2); wider: global-memory kernel
    const int mc = env.max_col;
    const bool reg_path = mc <= 30 && env.n >= 3 && !force_generic;
    if (reg_path) {
      const int P = mc <= 7 ? 8 : mc <= 15 ? 16 : mc <= 23 ? 24 : 31;
      if (map_P != P) {
        const int nthr = P == 8 ? RegCfg<8>::kPairThreads : P == 16 ? RegCfg<16>::kPairThreads : P == 24 ? RegCfg<24>::kPairThreads : RegCfg<31>::kPairThreads;
        std::vector<unsigned short> m = build_pair_map(P, nthr);
        size_t cnt = 0; for (auto v2 : m) cnt += v2 != 0xffff;
        if ((int)cnt != P * (P + 1) / 2) return fail(LVBA_ERR_UNSUPPORTED, "pair map for P=%d covers %zu of %d pairs", P, cnt, P * (P + 1) / 2);
        LVBA_TRY(pair_map.upload(m, s));
        map_P = P;
      }
      if (mc <= 7) env_factor_reg_kernel<8><<<1, RegCfg<8>::kThreads, RegCfg<8>::kSmem, s>>>(v, pair_map.p, L.p, dinv.p, z.p, status.p, dbg.p);
      else if (mc <= 15) env_factor_reg_kernel<16><<<1, RegCfg<16>::kThreads, RegCfg<16>::kSmem, s>>>(v, pair_map.p, L.p, dinv.p, z.p, status.p, dbg.p);
      else if (mc <= 23) env_factor_reg_kernel<24><<<1, RegCfg<24>::kThreads, RegCfg<24>::kSmem, s>>>(v, pair_map.p, L.p, dinv.p, z.p, status.p, dbg.p);
      else env_factor_reg_kernel<31><<<1, RegCfg<31>::kThreads, RegCfg<31>::kSmem, s>>>(v, pair_map.p, L.p, dinv.p, z.p, status.p, dbg.p);
      if (dbg.p && dbg_dumped < 2) {
        cudaStreamSynchronize(s);
        std::vector<long long> h((size_t)env.n * 32);
        cudaMemcpy(h.data(), dbg.p, h.size() * 8, cudaMemcpyDeviceToHost);
        // average per role: scale (1-0), barrier-1 wait (2-1), main phase (3-2), barrier-2 wait (next 0 - 3), step (next 0 - 0)
        const int k0 = 64, k1 = env.n - 64;
        fprintf(stderr, "[factor timing] n=%d max_col=%d   role: scale | bar1 wait | main | bar2 wait | step   (cycles, avg over k=%d..%d)\n", env.n, mc, k0, k1);
        for (int role = 0; role < 8; ++role) {
          double acc[5] = {0, 0, 0, 0, 0};
          for (int k = k0; k < k1; ++k) {
            const long long* a = &h[((size_t)k * 8 + role) * 4];
            const long long* b = &h[((size_t)(k + 1) * 8 + role) * 4];
            acc[0] += a[1] - a[0]; acc[1] += a[2] - a[1]; acc[2] += a[3] - a[2]; acc[3] += b[0] - a[3]; acc[4] += b[0] - a[0];
          }
          fprintf(stderr, "  role %d (%s): %8.0f %8.0f %8.0f %8.0f %8.0f\n", role, role < 4 ? "pair warp" : (role == 4 ? "inverse  " : role == 5 ? "fwd subst" : "prefetch "),
                  acc[0] / (k1 - k0), acc[1] / (k1 - k0), acc[2] / (k1 - k0), acc[3] / (k1 - k0), acc[4] / (k1 - k0));
        }
        ++dbg_dumped;
      }
      env_ldl_apply_kernel<<<(env.n + 127) / 128, 128, 0, s>>>(env.n, dinv.p, z.p, x);
      env_backsolve_ring_kernel<<<1, kBsThreads, 0, s>>>(v, L.p, x);
      *launches += 3;
    } else {
      env_factor_kernel<<<1, kFactorThreads, factor_smem(), s>>>(v, L.p, dinv.p, z.p, status.p);
      env_backsolve_kernel<<<1, 32, 0, s>>>(v, L.p, dinv.p, z.p, x);
      *launches += 2;
    }
    LVBA_CUDA(cudaGetLastError());
    return LVBA_OK;
  }
};

}  // namespace lvba
