"""Order independence of the device passes: on the GPU the items of one pass run concurrently in no particular order, the
sequential host policy of tests/emu/ visits them 0..n-1.  Re-running the host-policy suites with LVBA_EMU_SHUFFLE set
(items of every pass visited in a pseudo-random order) must give the same oracle-equal results — a pass that depended on an
earlier item of the same pass would not."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


SUITES = ["tests/test_voxel_emu.py", "tests/test_depth_emu.py", "tests/test_anchor_emu.py", "tests/test_wide_solver_emu.py",
          "tests/test_big_voxel_emu.py", "tests/test_track_emu.py", "tests/test_nd_solver_emu.py", "tests/test_fuse_emu.py"]
_runs = {}


def _asan():
    lib = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    return lib if lib and Path(lib).exists() else None


def _launch_all():
    """The three re-runs are independent child processes: start them together, each test then waits for its own (the CPU suite
    has to stay within a few minutes).  LVBA_EMU_RERUN makes the suites drop their largest systems — order independence and
    memory safety do not depend on the size."""
    if _runs:
        return
    base = dict(os.environ, LVBA_EMU_RERUN="1")
    envs = {"1": dict(base, LVBA_EMU_SHUFFLE="1"), "20260923": dict(base, LVBA_EMU_SHUFFLE="20260923")}
    asan = _asan()
    if asan:
        envs["asan"] = dict(base, LVBA_EMU_SANITIZE="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", LVBA_EMU_SHUFFLE="7")
    for key, env in envs.items():
        _runs[key] = subprocess.Popen([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", *SUITES],
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(ROOT), env=env)


def _wait(key, timeout):
    _launch_all()
    proc = _runs[key]
    try:
        out, err = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        proc.kill()
        out, err = proc.communicate()
        pytest.fail(f"re-run {key} timed out\n" + out[-2000:] + err[-2000:])
    return proc.returncode, out, err


@pytest.mark.parametrize("seed", ["1", "20260923"])
def test_host_policy_suites_pass_with_shuffled_items(seed):
    rc, out, err = _wait(seed, 900)
    assert rc == 0, out[-3000:] + err[-2000:]
    assert "passed" in out


def test_host_policy_suites_pass_under_address_and_ub_sanitizers():
    """The same functors compiled with -fsanitize=address,undefined: an out-of-bounds index in a pass (an illegal address on the
    device) or signed overflow / bad shifts in the key arithmetic aborts the run."""
    if _asan() is None:
        pytest.skip("libasan not available")
    rc, out, err = _wait("asan", 1800)
    assert rc == 0, out[-3000:] + err[-3000:]
    assert "passed" in out
