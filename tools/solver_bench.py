#!/usr/bin/env python
"""Device time of the block LDL^T paths (lvba_env_solve) on banded systems shaped like the pose / camera systems of the
BASELINE configs: n block rows, half-bandwidth b blocks.  Prints one JSON line per system."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as graft  # noqa: E402
import solver_systems as ss  # noqa: E402

pkg = graft.load_package(); pkg.load_library()
cases = [(2000, 30), (2000, 20), (5000, 30), (5000, 20)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for n, b in cases:
    first, blocks, dadd, rhs, A = ss.make([max(0, r - b) for r in range(n)], seed=n + b)
    xr = ss.reference_solve(A, rhs)
    out = {"n": n, "b": b}
    x, ms, info = pkg.env_solve(first, blocks, dadd, rhs, path=pkg.SOLVE_TWISTED, reps=5)
    out["twisted_ms"] = round(ms, 4); out["twisted_err"] = float(np.abs(x - xr).max())
    x, ms, info = pkg.env_solve(first, blocks, dadd, rhs, path=pkg.SOLVE_AUTO, reps=5)
    out["auto"] = {"ms": round(ms, 4), **info, "err": float(np.abs(x - xr).max())}
    for p in (4, 8, 16, 24, 32, 48, 64, 96):
        try:
            x, ms, info = pkg.env_solve(first, blocks, dadd, rhs, path=pkg.SOLVE_CHUNKED, chunks=p, reps=5)
            out[f"chunks_{p}"] = {"ms": round(ms, 4), "levels": info["levels"], "launches": info["launches"], "err": float(np.abs(x - xr).max())}
        except Exception as e:  # noqa: BLE001
            out[f"chunks_{p}"] = {"error": str(e)[:200]}
    print(json.dumps(out), flush=True)
