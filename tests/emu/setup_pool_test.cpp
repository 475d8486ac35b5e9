// Host test of lvba::parallel_chunks / SetupPool (global-lvba_b200/csrc/setup_pool.h): every index visited exactly once, worker ids
// inside [0, 8), nested regions and concurrent callers fall back to inline loops, a forked child starts from a fresh pool.
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <cstdio>
#include <numeric>

#include "../../global-lvba_b200/csrc/setup_pool.h"

static int check(int64_t n, int64_t min_chunk) {
  std::vector<int> hits((size_t)n, 0);
  std::atomic<int> bad{0};
  lvba::parallel_chunks(n, min_chunk, [&](int64_t a, int64_t b, int w) {
    if (w < 0 || w >= 8 || a > b || a < 0 || b > n) bad = 1;
    for (int64_t i = a; i < b; ++i) ++hits[(size_t)i];
  });
  for (int64_t i = 0; i < n; ++i) if (hits[(size_t)i] != 1) return 1;
  return bad.load();
}

extern "C" int setup_pool_selftest() {
  for (int rep = 0; rep < 200; ++rep)
    for (int64_t n : {0LL, 1LL, 7LL, 1000LL, 65536LL, 200000LL})
      if (check(n, rep % 3 == 0 ? 1 : 1 << 10)) return 1;
  // nested region: the inner one runs inline on whichever thread asks
  std::atomic<long long> sum{0};
  lvba::parallel_chunks(64, 1, [&](int64_t a, int64_t b, int) {
    for (int64_t i = a; i < b; ++i) lvba::parallel_chunks(100, 1, [&](int64_t c, int64_t d, int) { sum += d - c; });
  });
  if (sum.load() != 6400) return 2;
  // two callers at once: one of them runs inline, both are complete
  std::atomic<int> fails{0};
  std::thread other([&] { for (int r = 0; r < 100; ++r) if (check(50000, 1 << 10)) fails = 1; });
  for (int r = 0; r < 100; ++r) if (check(50000, 1 << 10)) fails = 1;
  other.join();
  if (fails.load()) return 3;
  // a forked child (the parent's helper threads do not exist there) still gets its regions done
  const pid_t pid = fork();
  if (pid == 0) _exit(check(100000, 1 << 10) ? 1 : 0);
  int st = 0;
  if (waitpid(pid, &st, 0) != pid || !WIFEXITED(st) || WEXITSTATUS(st) != 0) return 4;
  return check(100000, 1 << 10) ? 5 : 0;
}

#ifdef SETUP_POOL_MAIN
#include <chrono>
int main() {
  const int rc = setup_pool_selftest();
  std::vector<int64_t> a(200000);
  const auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < 200; ++r) lvba::parallel_chunks(200000, 1 << 14, [&](int64_t b, int64_t e, int) { for (int64_t i = b; i < e; ++i) a[(size_t)i] = i + r; });
  const auto t1 = std::chrono::steady_clock::now();
  std::printf("selftest rc=%d, %.1f us per region of 200k trivial items\n", rc, std::chrono::duration<double, std::micro>(t1 - t0).count() / 200);
  return rc;
}
#endif
