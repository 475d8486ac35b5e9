"""The off-ROS dataset loader (global-lvba_b200/host/lvba_dataset.hpp) and the offline tool's --check
mode, without a GPU: PCD in all three encodings (LZF with back references), TUM poses with comment / empty / unparsable
lines and un-normalised quaternions, timestamps from file names, scans sorted by timestamp."""
import json
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import dataset_writer as dw
from oracle import synth

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def tool(pkg, tmp_path_factory):
    exe = tmp_path_factory.mktemp("offline") / "lvba_offline"
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tools" / "lvba_offline.cpp"), "-o", str(exe),
           str(pkg.LIB_PATH), f"-Wl,-rpath,{pkg.LIB_PATH.parent}", "-L/usr/local/cuda/lib64", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_lzf_round_trip_uses_back_references():
    rng = np.random.default_rng(0)
    raw = (np.repeat(rng.integers(0, 255, 400, dtype=np.uint8), 9).tobytes() + bytes(5000) + rng.integers(0, 255, 3000, dtype=np.uint8).tobytes())
    comp = dw.lzf_compress(raw)
    assert len(comp) < len(raw) // 2 and any(c >= 32 for c in comp[:64])
    assert dw.lzf_decompress(comp, len(raw)) == raw


def test_check_mode_reads_what_was_written(tool, tmp_path, pkg):
    scans, poses = synth.make_scan_scene(3, W=7, n_per_scan=900)
    scans[4] = scans[4][:0]                                             # an empty scan file
    ts = dw.write_lidar_dataset(tmp_path, scans, poses)
    (tmp_path / "all_pcd_body" / "notes.txt").write_text("ignored")
    (tmp_path / "all_pcd_body" / "no_number_here.pcd").write_text("junk")           # bad name: skipped with a warning
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["scans"] == 7 and info["poses"] == 7 and info["points"] == sum(len(s) for s in scans)
    want = sum(float(np.asarray(s, np.float64).sum()) for s in scans)
    assert abs(info["coordinate_sum"] - want) <= 1e-6 * max(1.0, abs(want))        # float32 survives all three encodings exactly
    w = np.concatenate([np.arange(1, 10), np.arange(10, 13)]).astype(np.float64)
    pose_sum = 0.0
    for i in range(7):                                                   # written un-normalised, 15 digits: R to ~1e-14
        q = dw.R_to_quat(poses[i, :9].reshape(3, 3))
        pose_sum += float(w[:9] @ dw.quat_to_R(q).ravel() + w[9:] @ poses[i, 9:])
    assert abs(info["pose_sum"] - pose_sum) <= 1e-8
    assert abs(info["first_ts"] - ts[0]) < 1e-9 and abs(info["last_ts"] - ts[-1]) < 1e-9   # from the file names, not the TUM column
    assert "bad pcd name" in r.stderr and "unparsable" in r.stderr
    if pkg.device_count() == 0:                                          # the full run needs the GPU library: loud refusal
        r = subprocess.run([str(tool), "--data", str(tmp_path)], capture_output=True, text=True)
        assert r.returncode == 2 and "no CUDA device" in r.stderr


def test_loader_failures(tool, tmp_path):
    r = subprocess.run([str(tool), "--data", str(tmp_path / "nowhere"), "--check"], capture_output=True, text=True)
    assert r.returncode == 1 and "pcd dir missing" in r.stderr
    d = tmp_path / "all_pcd_body"; d.mkdir()
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--check"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr
    (d / "lidar_poses.txt").write_text("0 0 0 0 0 0 0 1\n")
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--check"], capture_output=True, text=True)
    assert r.returncode == 1 and "no pcd files" in r.stderr
    (d / "1.5.pcd").write_bytes(b"VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary\n" + bytes(10))
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--check"], capture_output=True, text=True)
    assert r.returncode == 0 and "does not fit the file" in r.stderr and json.loads(r.stdout.strip().splitlines()[-1])["scans"] == 0


@pytest.mark.parametrize("header,msg", [
    (b"FIELDS x y z\nSIZE -4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary\n", "SIZE must be"),
    (b"FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 0 1\nWIDTH 2\nHEIGHT 1\nPOINTS 2\nDATA binary\n", "COUNT out of range"),
    (b"FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 2\nHEIGHT 1\nPOINTS 18446744073709551615\nDATA binary\n", "overflows"),
    (b"FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 40000000\nHEIGHT 4000\nDATA binary_compressed\n", "does not fit"),
    (b"FIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 99999999999\nHEIGHT 1\nPOINTS 99999999999\nDATA ascii\n1 2 3\n", "does not fit"),
])
def test_hostile_pcd_headers_are_refused_not_trusted(tool, tmp_path, header, msg):
    """SIZE / COUNT / POINTS come from the file; a bad one must fail that file, never index outside a buffer or throw."""
    d = tmp_path / "all_pcd_body"; d.mkdir()
    (d / "lidar_poses.txt").write_text("0 0 0 0 0 0 0 1\n")
    (d / "1.5.pcd").write_bytes(b"VERSION 0.7\n" + header + bytes(64))
    r = subprocess.run([str(tool), "--data", str(tmp_path), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert msg in r.stderr, r.stderr
    assert json.loads(r.stdout.strip().splitlines()[-1])["scans"] == 0
