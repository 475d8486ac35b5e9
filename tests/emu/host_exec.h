// host_exec.h — TEST INFRASTRUCTURE: the sequential host execution policy the no-GPU tests instantiate the device
// pipelines with (voxel_pipeline.h, depth_pipeline.h).  Never part of liblvba_b200.so.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <random>
#include <vector>

struct HostExec {
  template <class T>
  struct Buf {
    T* p = nullptr;
    size_t n = 0;
    std::vector<T> v;
    // device allocations come back uninitialised (and recycled through the pool): poison the host copy the same way, so that a
    // pass relying on zero-filled memory cannot pass here
    int alloc(size_t count) { v.resize(count); if (count) std::memset((void*)v.data(), 0xA5, count * sizeof(T)); p = v.data(); n = count; return 0; }
    void swap(Buf& o) { v.swap(o.v); std::swap(p, o.p); std::swap(n, o.n); }
  };
  // On the device the items of a pass run concurrently in no particular order.  With LVBA_EMU_SHUFFLE=<seed> the items are
  // visited in a pseudo-random order instead of 0..n-1, so a pass that silently relied on sequential execution (reading what
  // an earlier item of the SAME pass wrote) gives different results and the tests that compare against the oracle catch it.
  template <class F>
  int for_each(int64_t n, const F& f) {
    static const char* env = std::getenv("LVBA_EMU_SHUFFLE");
    if (!env || n < 2) { for (int64_t i = 0; i < n; ++i) f(i); return 0; }
    std::vector<int64_t> order((size_t)n);
    std::iota(order.begin(), order.end(), 0);
    std::mt19937_64 rng((uint64_t)std::atoll(env) * 0x9E3779B97F4A7C15ull + (uint64_t)n);
    std::shuffle(order.begin(), order.end(), rng);
    for (int64_t i : order) f(i);
    return 0;
  }
  template <class T>
  int fill_zero(T* p, size_t n) { std::memset(p, 0, n * sizeof(T)); return 0; }
  template <class T>
  int put(T* dev, const T* host, size_t n) { std::memcpy(dev, host, n * sizeof(T)); return 0; }
  template <class T>
  int fetch(T* host, const T* dev, size_t n) { std::memcpy(host, dev, n * sizeof(T)); return 0; }
  int min_max(const int32_t* p, int64_t n, int32_t* mn, int32_t* mx) {
    *mn = *std::min_element(p, p + n); *mx = *std::max_element(p, p + n); return 0;
  }
  // stable, on key bits [0, end_bit) — what an LSD radix sort gives
  int sort_pairs(const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout, int64_t n, int end_bit) {
    const uint64_t mask = end_bit >= 64 ? ~0ull : ((1ull << end_bit) - 1);
    std::vector<int64_t> perm((size_t)n);
    std::iota(perm.begin(), perm.end(), 0);
    std::stable_sort(perm.begin(), perm.end(), [&](int64_t a, int64_t b) { return (kin[a] & mask) < (kin[b] & mask); });
    for (int64_t i = 0; i < n; ++i) { kout[i] = kin[perm[i]]; vout[i] = vin[perm[i]]; }
    return 0;
  }
  template <class T>
  int exclusive_scan(const T* in, T* out, int64_t n) { T acc = 0; for (int64_t i = 0; i < n; ++i) { const T v = in[i]; out[i] = acc; acc += v; } return 0; }
  int sync() { return 0; }
};
