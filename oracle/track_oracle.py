"""CPU restatement of the per-track numerics of the reference's track fusion — TEST INFRASTRUCTURE:
  TriangulateTrackDLT   src/lvba_system.cpp:52-111   ;   ComputeMeanReproj   src/lvba_system.cpp:8-50
(undistortPixelToNormalized / projectWorldToPixel: include/utils.hpp:183-233, restated in oracle/depth_oracle.py).
PARITY: pinned against the reference's own file-scope functions, called from src/lvba_system.cpp compiled where it lies (tests/test_ref_system_pin.py:
directly on 400 tracks to 1e-7 — LAPACK eigh here, Jacobi under the reference's lines — and through BuildTracksAndFuse3D to 1e-12)."""
from __future__ import annotations

import numpy as np

from oracle import depth_oracle as dep


def mean_reproj(Xw, cams_sel, uv_sel, intr, min_count):
    """Returns (ok, mean, count) over the selected observations (cams_sel (k, 12), uv_sel (k, 2))."""
    s = 0.0; cnt = 0
    for cam, uv in zip(cams_sel, uv_sel):
        R = cam[:9].reshape(3, 3); t = cam[9:]
        ok, uu, vv = dep.project((R @ Xw + t)[None, :], intr)
        if not ok[0]:
            continue
        s += float(np.hypot(uu[0] - float(uv[0]), vv[0] - float(uv[1]))); cnt += 1
    if cnt < min_count:
        return False, 0.0, cnt
    m = s / cnt
    return bool(np.isfinite(m)), m, cnt


def triangulate_dlt(cams_sel, uv_sel, intr):
    """TriangulateTrackDLT.  Returns (ok, Xw, mean_reproj, count)."""
    if len(cams_sel) < 4:
        return False, np.zeros(3), 0.0, 0
    AtA = np.zeros((4, 4)); rows = 0
    for cam, uv in zip(cams_sel, uv_sel):
        ok, x, y = dep.undistort_pixel(intr, float(uv[0]), float(uv[1]))
        if not ok:
            continue
        P = np.column_stack([cam[:9].reshape(3, 3), cam[9:]])
        ru = x * P[2] - P[0]; rv = y * P[2] - P[1]
        AtA += np.outer(ru, ru); AtA += np.outer(rv, rv); rows += 2
    if rows < 8:
        return False, np.zeros(3), 0.0, 0
    w, V = np.linalg.eigh(AtA)
    Xh = V[:, 0]
    if abs(Xh[3]) < 1e-12:
        return False, np.zeros(3), 0.0, 0
    X = Xh[:3] / Xh[3]
    if not np.all(np.isfinite(X)):
        return False, np.zeros(3), 0.0, 0
    ok, m, cnt = mean_reproj(X, cams_sel, uv_sel, intr, 4)
    return ok, X, (m if ok else 0.0), cnt
