"""Path A for voxels seen from more poses than one batch CTA holds (global-lvba_b200/csrc/lidar_big.h) checked without a
GPU: the three passes (voxel parameters, slots, slot pairs) run by plain loops over every voxel of a problem
(tests/emu/big_emu.cpp) against oracle/lidar_oracle.acc_evaluate2 — residual, gradient and every Hessian block, including
a voxel seen from 300 poses."""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import lidar_oracle as lo
from oracle import synth

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("emu") / "libbig_emu.so"
    cmd = ["g++", "-std=c++17", "-O2", *(["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-g"] if os.environ.get("LVBA_EMU_SANITIZE") else []), "-Wall", "-fPIC", "-shared", str(ROOT / "tests" / "emu" / "big_emu.cpp"), "-o", str(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = ctypes.CDLL(str(so))
    lib.emu_big_accumulate.restype = ctypes.c_double
    return lib


def envelope_of(vox_ptr, pose_idx, W):
    first = np.arange(W)
    for a in range(len(vox_ptr) - 1):
        s = pose_idx[vox_ptr[a]:vox_ptr[a + 1]]
        first[s] = np.minimum(first[s], s.min())
    for r in range(W - 2, -1, -1):
        first[r] = min(first[r], first[r + 1])
    row_start = np.zeros(W + 1, np.int64)
    row_start[1:] = np.cumsum(np.arange(W) - first + 1)
    return first.astype(np.int32), row_start


def run(emu, p, W, residual_only=False):
    vp = np.ascontiguousarray(p["vox_ptr"], np.int64); pi = np.ascontiguousarray(p["pose_idx"], np.int32)
    cl = np.ascontiguousarray(p["clusters"], np.float64); ps = np.ascontiguousarray(p["poses"], np.float64)
    first, row_start = envelope_of(vp, pi, W)
    H = np.zeros((row_start[-1], 36)); g = np.zeros((W, 6))
    P = ctypes.POINTER
    r = emu.emu_big_accumulate(ctypes.c_int64(len(vp) - 1), vp.ctypes.data_as(P(ctypes.c_int64)), pi.ctypes.data_as(P(ctypes.c_int32)),
                               cl.ctypes.data_as(P(ctypes.c_double)), ps.ctypes.data_as(P(ctypes.c_double)), first.ctypes.data_as(P(ctypes.c_int)),
                               row_start.ctypes.data_as(P(ctypes.c_longlong)), H.ctypes.data_as(P(ctypes.c_double)), g.ctypes.data_as(P(ctypes.c_double)),
                               ctypes.c_int(int(residual_only)))
    dense = np.zeros((6 * W, 6 * W))
    for row in range(W):
        for c in range(first[row], row + 1):
            dense[6 * row:6 * row + 6, 6 * c:6 * c + 6] = H[row_start[row] + c - first[row]].reshape(6, 6)
    return r, g, dense


def check(emu, p, W):
    r, g, lower = run(emu, p, W)
    r_ref, g_ref, blocks = lo.acc_evaluate2(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"], W)
    H_ref = lo.assemble_dense(blocks, W)
    assert abs(r - r_ref) <= 1e-8 * abs(r_ref)                       # lambda_0 is a difference of O(1e4) terms
    assert np.abs(g - g_ref).max() <= 1e-7 * np.abs(g_ref).max()
    mask = np.kron(np.tril(np.ones((W, W))), np.ones((6, 6))) > 0   # block lower triangle incl. full diagonal blocks
    assert np.abs(lower - H_ref * mask).max() <= 1e-7 * np.abs(H_ref).max()
    assert np.abs(lower[~mask]).max() == 0.0
    assert abs(run(emu, p, W, residual_only=True)[0] - r_ref) <= 1e-8 * abs(r_ref)


def test_passes_equal_oracle_on_ordinary_voxels(emu):
    check(emu, synth.make_problem(30, 600, 0, seed=11, visual=False), 30)
    check(emu, synth.make_problem(12, 80, 0, seed=4242, visual=False), 12)


def test_voxels_seen_from_hundreds_of_poses(emu):
    """k_lo = k_hi = 300 distinct poses per voxel out of 400: the case the 128-thread batch kernels cannot hold."""
    rng = np.random.Generator(np.random.Philox(key=77))
    R_gt, p_gt = synth.make_trajectory(400, rng)
    vp, pi, cl = synth.make_lidar(400, 6, R_gt, p_gt, rng, k_lo=300, k_hi=300, half=399)
    assert np.diff(vp).max() >= 290
    R0 = R_gt @ synth.so3_exp(rng.normal(0, 0.003, (400, 3)))
    poses = np.concatenate([R0.reshape(400, 9), p_gt + rng.normal(0, 0.02, (400, 3))], 1)
    check(emu, dict(vox_ptr=vp, pose_idx=pi, clusters=cl, poses=poses), 400)
