// wide_emu.cpp — TEST INFRASTRUCTURE: the pass functors of global-lvba_b200/csrc/envelope_wide.h (the any-width LDL^T the
// device runs for loop-closure envelopes) executed with a plain loop, so that tests/test_wide_solver_emu.py can check the
// arithmetic against a dense numpy solve without a GPU.  Never part of the product.
#include <vector>

#include "../../global-lvba_b200/csrc/envelope_wide.h"
#include "host_exec.h"

extern "C" int emu_wide_solve(int n, const int* first, const int* last, const long long* row_start, double* L, double* z, double* x,
                              double* dinv) {
  lvba::wide::View e{n, first, row_start};
  int max_col = 0;
  for (int k = 0; k < n; ++k) if (last[k] - k > max_col) max_col = last[k] - k;
  std::vector<double> colT((size_t)(max_col > 0 ? max_col : 1) * 36);
  int status = 0;
  HostExec ex;
  auto launch = [&](int64_t items, const auto& f) { ex.for_each(items, f); };
  lvba::wide::factor_and_solve(launch, e, first, last, L, dinv, z, colT.data(), x, &status);
  return status;
}
