"""Anchor clouds of runWindowBA (global-lvba_b200/csrc/anchor_pipeline.h — boundary B6: pl_transform of every scan into its
anchor frame + down_sampling_voxel2) without a GPU: the device passes through the sequential host policy against
oracle/anchor_oracle.py — the surviving points EXACTLY (float32), ordered by voxel key."""
import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import anchor_oracle as ao
from oracle import synth

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("emu") / "libanchor_emu.so"
    cmd = ["g++", "-std=c++17", "-O2", *(["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-g"] if os.environ.get("LVBA_EMU_SANITIZE") else []), "-ffp-contract=off", "-Wall", "-fPIC", "-shared", str(ROOT / "tests" / "emu" / "anchor_emu.cpp"), "-o", str(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return ctypes.CDLL(str(so))


def _ptr(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def run(emu, scans, rel, win_ptr, leaf):
    sp = np.zeros(len(scans) + 1, np.int64); sp[1:] = np.cumsum([len(s) for s in scans])
    xyz = np.ascontiguousarray(np.concatenate(scans) if len(scans) else np.zeros((0, 3)), np.float32)
    wp = np.ascontiguousarray(win_ptr, np.int32); rl = np.ascontiguousarray(rel, np.float64)
    h = ctypes.c_void_p(); n = ctypes.c_int64()
    rc = emu.emu_anchor_create(ctypes.c_int32(len(wp) - 1), _ptr(wp, ctypes.c_int32), _ptr(sp, ctypes.c_int64), _ptr(xyz, ctypes.c_float),
                               _ptr(rl, ctypes.c_double), ctypes.c_double(leaf), ctypes.byref(h), ctypes.byref(n))
    if rc != 0:
        return rc, None
    cp = np.zeros(len(wp), np.int64); out = np.zeros((n.value, 3), np.float32)
    emu.emu_anchor_export(h, _ptr(cp, ctypes.c_int64), _ptr(out, ctypes.c_float)); emu.emu_anchor_destroy(h)
    return 0, [out[cp[w]:cp[w + 1]] for w in range(len(wp) - 1)]


@pytest.mark.parametrize("leaf", [0.1, 0.25, 1.0])
def test_anchor_clouds_equal_oracle_exactly(emu, leaf):
    sizes = [4, 1, 5, 3]
    scans, poses = synth.make_scan_scene(15, W=sum(sizes), n_per_scan=1500)
    scans[6] = scans[6][:0]                                         # an empty scan inside a window
    win_ptr = np.concatenate([[0], np.cumsum(sizes)])
    rel = ao.rel_poses(poses, win_ptr)
    rc, got = run(emu, scans, rel, win_ptr, leaf)
    assert rc == 0
    ref = ao.anchor_clouds(scans, rel, win_ptr, leaf)
    lit = ao.anchor_clouds_literal(scans, rel, win_ptr, leaf)
    for w in range(len(sizes)):
        assert np.array_equal(ref[w], lit[w])
        assert np.array_equal(got[w], ref[w]), w
        n_in = sum(len(scans[j]) for j in range(win_ptr[w], win_ptr[w + 1]))
        assert 0 < len(got[w]) <= n_in and (leaf < 0.2 or len(got[w]) < n_in)


def test_ties_pass_through_and_bad_points(emu):
    # two points at the same distance from the voxel centre: the first in cloud order survives (strict '<', tools.hpp:342)
    scans = [np.array([[0.25, 0.5, 0.5], [0.75, 0.5, 0.5], [0.75, 0.5, 0.5001]], np.float32), np.array([[0.75, 0.5, 0.5]], np.float32)]
    rel = np.tile(np.concatenate([np.eye(3).ravel(), np.zeros(3)]), (2, 1))
    rc, got = run(emu, scans, rel, [0, 2], 1.0)
    assert rc == 0 and got[0].tolist() == [[0.25, 0.5, 0.5]]
    # leaf < 0.001: no down-sampling, transformed points in cloud order (:303)
    s2, poses = synth.make_scan_scene(16, W=3, n_per_scan=300)
    rel = ao.rel_poses(poses, np.array([0, 3]))
    rc, got = run(emu, s2, rel, [0, 3], 0.0005)
    assert rc == 0 and np.array_equal(got[0], np.concatenate([ao.transform(s2[j], rel[j]) for j in range(3)]))
    bad = [x.copy() for x in s2]; bad[1][5, 2] = np.nan
    assert run(emu, bad, rel, [0, 3], 0.1)[0] == -1
    rc, got = run(emu, [np.zeros((0, 3), np.float32)] * 3, rel, [0, 1, 3], 0.1)
    assert rc == 0 and all(len(g) == 0 for g in got)
