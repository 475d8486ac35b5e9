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


@pytest.mark.parametrize("seed", ["1", "20260923"])
def test_host_policy_suites_pass_with_shuffled_items(seed):
    env = dict(os.environ, LVBA_EMU_SHUFFLE=seed)
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        "tests/test_voxel_emu.py", "tests/test_depth_emu.py", "tests/test_anchor_emu.py",
                        "tests/test_wide_solver_emu.py", "tests/test_big_voxel_emu.py", "tests/test_track_emu.py", "tests/test_nd_solver_emu.py"],
                       capture_output=True, text=True, cwd=str(ROOT), env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "passed" in r.stdout


def test_host_policy_suites_pass_under_address_and_ub_sanitizers():
    """The same functors compiled with -fsanitize=address,undefined: an out-of-bounds index in a pass (an illegal address on the
    device) or signed overflow / bad shifts in the key arithmetic aborts the run."""
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not Path(asan).exists():
        pytest.skip("libasan not available")
    env = dict(os.environ, LVBA_EMU_SANITIZE="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1", LVBA_EMU_SHUFFLE="7")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        "tests/test_voxel_emu.py", "tests/test_depth_emu.py", "tests/test_anchor_emu.py",
                        "tests/test_wide_solver_emu.py", "tests/test_big_voxel_emu.py", "tests/test_track_emu.py", "tests/test_nd_solver_emu.py"],
                       capture_output=True, text=True, cwd=str(ROOT), env=env, timeout=1800)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
