"""The any-width block LDL^T (global-lvba_b200/csrc/envelope_wide.h — what the device runs when a loop closure makes the
envelope wider than the in-SM solvers hold) checked without a GPU: the same pass functors, run by a plain loop
(tests/emu/wide_emu.cpp), against a dense numpy solve of the same symmetric INDEFINITE block system (the BALM2 Newton
Hessian is indefinite, SURVEY.md Q5: LDL^T without pivoting, as SimplicialLDLT)."""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("emu") / "libwide_emu.so"
    cmd = ["g++", "-std=c++17", "-O2", *(["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-g"] if os.environ.get("LVBA_EMU_SANITIZE") else []), "-Wall", "-fPIC", "-shared", str(ROOT / "tests" / "emu" / "wide_emu.cpp"), "-o", str(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return ctypes.CDLL(str(so))


def envelope(first_raw):
    """Envelope::build (csrc/runtime.cuh): first made monotone, row offsets, last[k] = max row coupled to column k."""
    n = len(first_raw)
    first = np.minimum(np.asarray(first_raw), np.arange(n))
    for r in range(n - 2, -1, -1):
        first[r] = min(first[r], first[r + 1])
    row_start = np.zeros(n + 1, np.int64)
    row_start[1:] = np.cumsum(np.arange(n) - first + 1)
    last = np.array([max(i for i in range(k, n) if first[i] <= k) for k in range(n)], np.int32)
    return first.astype(np.int32), last, row_start


def random_system(n, first, row_start, rng, indefinite):
    """Dense symmetric matrix whose lower blocks live inside the envelope + the envelope storage of its lower triangle."""
    M = np.zeros((6 * n, 6 * n))
    for r in range(n):
        for c in range(first[r], r):
            if rng.random() < 0.7:
                M[6 * r:6 * r + 6, 6 * c:6 * c + 6] = rng.normal(0, 1, (6, 6))
    M = M + M.T
    for r in range(n):
        D = rng.normal(0, 1, (6, 6)); D = D @ D.T + (12.0 * (r - first[r] + 2)) * np.eye(6)      # block diagonally dominant
        if indefinite and r % 3 == 1:
            D = -D                                                                                # negative pivots: still LDL^T-able
        M[6 * r:6 * r + 6, 6 * r:6 * r + 6] = D
    L = np.zeros((row_start[-1], 36))
    for r in range(n):
        for c in range(first[r], r + 1):
            blk = M[6 * r:6 * r + 6, 6 * c:6 * c + 6].copy()
            if c == r:
                blk = np.tril(blk) + np.triu(rng.normal(0, 99, (6, 6)), 1)                        # the upper triangle of a diagonal block is never read
            L[row_start[r] + c - first[r]] = blk.ravel()
    return M, L


@pytest.mark.parametrize("n,seed,indefinite", [(1, 0, False), (2, 1, False), (9, 2, False), (30, 3, True), (60, 4, True)])
def test_solve_equals_dense_numpy(emu, n, seed, indefinite):
    rng = np.random.default_rng(seed)
    raw = np.maximum(np.arange(n) - rng.integers(0, 6, n), 0)
    if n > 20:
        raw[n - 3] = 1                                   # a loop closure: row n-3 couples to column 1 -> every row between reaches back
    first, last, row_start = envelope(raw)
    M, L = random_system(n, first, row_start, rng, indefinite)
    b = rng.normal(0, 1, 6 * n)
    z = b.copy(); x = np.zeros(6 * n); dinv = np.zeros((n, 36))
    P = ctypes.POINTER
    rc = emu.emu_wide_solve(ctypes.c_int(n), first.ctypes.data_as(P(ctypes.c_int)), last.ctypes.data_as(P(ctypes.c_int)),
                            row_start.ctypes.data_as(P(ctypes.c_longlong)), L.ctypes.data_as(P(ctypes.c_double)),
                            z.ctypes.data_as(P(ctypes.c_double)), x.ctypes.data_as(P(ctypes.c_double)), dinv.ctypes.data_as(P(ctypes.c_double)))
    assert rc == 0
    ref = np.linalg.solve(M, b)
    assert np.abs(x - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    assert np.abs(M @ x - b).max() <= 1e-9 * np.abs(b).max()
    if n > 20:
        assert (last - np.arange(n)).max() >= n - 5      # the closure did widen the envelope to (almost) the whole system


def test_singular_pivot_is_flagged(emu):
    n = 3
    first, last, row_start = envelope([0, 0, 1])
    L = np.zeros((row_start[-1], 36)); L[0] = np.zeros(36)          # a zero pivot block
    z = np.ones(6 * n); x = np.zeros(6 * n); dinv = np.zeros((n, 36))
    P = ctypes.POINTER
    rc = emu.emu_wide_solve(ctypes.c_int(n), first.ctypes.data_as(P(ctypes.c_int)), last.ctypes.data_as(P(ctypes.c_int)),
                            row_start.ctypes.data_as(P(ctypes.c_longlong)), L.ctypes.data_as(P(ctypes.c_double)),
                            z.ctypes.data_as(P(ctypes.c_double)), x.ctypes.data_as(P(ctypes.c_double)), dinv.ctypes.data_as(P(ctypes.c_double)))
    assert rc == 1
