import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import __graft_entry__ as graft  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real CUDA device (run on the B200 box)")


def pytest_collection_finish(session):
    """tests/test_emu_shuffled.py re-runs the host-policy suites in three child processes; when it is part of the run, start
    them now so that they work beside the rest of the (mostly single-threaded) CPU suite instead of after it."""
    for item in session.items:
        if item.nodeid.startswith("tests/test_emu_shuffled.py") or "/test_emu_shuffled.py" in item.nodeid:
            if not session.config.option.collectonly:
                item.module._launch_all()
            break


@pytest.fixture(scope="session")
def pkg():
    """The ctypes binding of liblvba_b200.so (built on demand; never a CPU fallback)."""
    p = graft.load_package()
    p.build_library()
    p.load_library()
    return p


@pytest.fixture(scope="session")
def gpu_pkg(pkg):
    if pkg.device_count() < 1:
        pytest.fail("test marked gpu but no CUDA device is visible")
    return pkg


_cache = {}


@pytest.fixture(scope="session")
def problem_small():
    from oracle import synth
    if "small" not in _cache:
        _cache["small"] = synth.make_problem(30, 600, 300, seed=11)
    return _cache["small"]


@pytest.fixture(scope="session")
def problem_A():
    from oracle import synth
    if "A" not in _cache:
        _cache["A"] = synth.make_config("A")
    return _cache["A"]
