"""The helper threads behind the host set-up loops of the one-shot calls (global-lvba_b200/csrc/setup_pool.h): a persistent pool
replaces per-region std::thread spawns.  Plain C++, so it is tested here without a GPU: every index exactly once, nested and
concurrent regions, a forked child; once more under ThreadSanitizer when the toolchain has it."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "emu" / "setup_pool_test.cpp"


def _build_and_run(tmp_path, extra, env=None):
    exe = tmp_path / "pooltest"
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-pthread", "-DSETUP_POOL_MAIN", *extra, str(SRC), "-o", str(exe)], capture_output=True, text=True)
    if r.returncode != 0:
        return None, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300, env=env)
    return r.returncode, r.stdout + r.stderr


def test_parallel_chunks_visits_every_index_once(tmp_path):
    rc, out = _build_and_run(tmp_path, ["-O2"])
    assert rc == 0, out
    assert "selftest rc=0" in out


def test_no_data_race_under_thread_sanitizer(tmp_path):
    import os
    rc, out = _build_and_run(tmp_path, ["-O1", "-g", "-fsanitize=thread"], env=dict(os.environ, TSAN_OPTIONS="die_after_fork=0"))
    if rc is None:
        pytest.skip("ThreadSanitizer runtime not available: " + out[-200:])
    assert rc == 0 and "WARNING: ThreadSanitizer" not in out, out[-3000:]
