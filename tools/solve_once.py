#!/usr/bin/env python
"""One lvba_env_solve call (for ncu): python tools/solve_once.py n b path chunks reps"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as graft  # noqa: E402
import solver_systems as ss  # noqa: E402
pkg = graft.load_package(); pkg.load_library()
n, b, path, chunks, reps = (int(v) for v in sys.argv[1:6])
first, blocks, dadd, rhs, A = ss.make([max(0, r - b) for r in range(n)], seed=n + b)
x, ms, info = pkg.env_solve(first, blocks, dadd, rhs, path=path, chunks=chunks, reps=reps)
print("ms", ms, info, "resid", float(np.abs(A @ x - rhs).max()))
