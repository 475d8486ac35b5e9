"""The per-track numerics of the track fusion (global-lvba_b200/csrc/track_pipeline.h: DLT triangulation and mean
reprojection error, src/lvba_system.cpp:8-111) without a GPU: the device functors run by plain loops (tests/emu/track_emu.cpp)
against the numpy restatement oracle/track_oracle.py, on the tracks of a synthetic visual problem."""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import dataset_writer as dw
from oracle import synth
from oracle import track_oracle as tro

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("emu") / "libtrack_emu.so"
    cmd = ["g++", "-std=c++17", "-O2", *(["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-g"] if os.environ.get("LVBA_EMU_SANITIZE") else []), "-ffp-contract=off", "-Wall", "-fPIC", "-shared", str(ROOT / "tests" / "emu" / "track_emu.cpp"), "-o", str(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return ctypes.CDLL(str(so))


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _problem():
    p = synth.make_problem(40, 0, 400, seed=9, lidar=False)
    cams = np.zeros((40, 12))
    for k in range(40):
        cams[k, :9] = dw.quat_to_R(p["q_gt"][k]).ravel(); cams[k, 9:] = p["t_gt"][k]
    op = np.ascontiguousarray(p["obs_ptr"], np.int64); oc = np.ascontiguousarray(p["obs_cam"], np.int32)
    uv = np.ascontiguousarray(p["obs_uv"], np.float32); intr = np.ascontiguousarray(p["intr"], np.float64)
    return p, cams, op, oc, uv, intr


def test_triangulation_equals_numpy_restatement(emu):
    p, cams, op, oc, uv, intr = _problem()
    oc = oc.copy(); oc[op[7]] = 99; oc[op[8] + 1] = -3                      # out-of-range camera ids are skipped (:74-78)
    T = len(op) - 1
    Xw = np.full((T, 3), 7.0); mean = np.full(T, 7.0); cnt = np.full(T, 7, np.int32); ok = np.full(T, 7, np.uint8)
    emu.emu_tracks_triangulate(ctypes.c_int64(T), _ptr(op, ctypes.c_int64), _ptr(oc, ctypes.c_int32), _ptr(uv, ctypes.c_float), ctypes.c_int32(40),
                               _ptr(cams, ctypes.c_double), _ptr(intr, ctypes.c_double), _ptr(Xw, ctypes.c_double), _ptr(mean, ctypes.c_double),
                               _ptr(cnt, ctypes.c_int32), ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    n_ok = 0
    for t in range(T):
        sel = [q for q in range(op[t], op[t + 1]) if 0 <= oc[q] < 40]
        r_ok, X, m, c = tro.triangulate_dlt(cams[oc[sel]], uv[sel], intr) if op[t + 1] - op[t] >= 4 else (False, np.zeros(3), 0.0, 0)
        assert bool(ok[t]) == r_ok, t
        if r_ok:
            n_ok += 1
            assert np.abs(Xw[t] - X).max() <= 1e-7 * max(1.0, np.abs(X).max()) and abs(mean[t] - m) <= 1e-7 and cnt[t] == c
            assert np.linalg.norm(Xw[t] - p["X_gt"][t]) < 3.0 and mean[t] < 3.0   # half-pixel noise, short baselines: still near the truth
    short = np.diff(op) < 4
    assert n_ok > 100 and short.any() and not ok[short].any()                     # tracks with < 4 views never triangulate


def test_mean_reprojection_equals_numpy_restatement(emu):
    p, cams, op, oc, uv, intr = _problem()
    T = len(op) - 1
    X = np.ascontiguousarray(p["X_gt"] + 0.01)
    mean = np.zeros(T); cnt = np.zeros(T, np.int32); ok = np.zeros(T, np.uint8)
    emu.emu_tracks_mean_reproj(ctypes.c_int64(T), _ptr(op, ctypes.c_int64), _ptr(oc, ctypes.c_int32), _ptr(uv, ctypes.c_float), ctypes.c_int32(40),
                               _ptr(cams, ctypes.c_double), _ptr(intr, ctypes.c_double), _ptr(X, ctypes.c_double), ctypes.c_int32(5),
                               _ptr(mean, ctypes.c_double), _ptr(cnt, ctypes.c_int32), ok.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)))
    for t in range(0, T, 7):
        sel = list(range(op[t], op[t + 1]))
        r_ok, m, c = tro.mean_reproj(X[t], cams[oc[sel]], uv[sel], intr, 5)
        assert bool(ok[t]) == r_ok and cnt[t] == c and (not r_ok or abs(mean[t] - m) <= 1e-9)
    assert ok.any() and not ok.all()                                              # min_count = 5 splits the tracks
