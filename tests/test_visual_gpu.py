"""GPU parity tests for boundary B2 (visual LM) — CUDA path vs the numpy oracle, through the C ABI.

Tolerances (float64): cost rel 1e-10 per evaluation; reduced camera system / rhs 1e-8 of their max;
first LM step 1e-6 of max|step| (the normal equations are solved by different eliminations: sparse LU on
the full system in the oracle, Schur + block LDL^T on the GPU); final cost rel 1e-6 (north star).
"""
import numpy as np
import pytest

from oracle import synth
from oracle import visual_oracle as vo

pytestmark = pytest.mark.gpu

KEYS = ("q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr", "sigma_px", "sigma_plane")


def _args(p):
    return [p[k] for k in KEYS]


def _oracle(p, **kw):
    return vo.VisualProblem(*_args(p), **kw)


def test_cost_matches(gpu_pkg, problem_A):
    p = problem_A
    P = gpu_pkg.VisualProblem(*_args(p))
    pr = _oracle(p)
    assert abs(P.cost() - pr.cost()) <= 1e-10 * pr.cost()
    P.set_state(p["q_gt"], p["t_gt"], p["X_gt"])
    c_gt = pr.cost(p["q_gt"], p["t_gt"], p["X_gt"])
    assert abs(P.cost() - c_gt) <= 1e-10 * c_gt
    P.close()


@pytest.mark.parametrize("scaling", [True, False])
@pytest.mark.parametrize("radius", [1e4, 3.0])
def test_single_step_matches(gpu_pkg, problem_small, scaling, radius):
    p = problem_small
    P = gpu_pkg.VisualProblem(*_args(p))
    cs, ps, model, cost = P.step(radius, jacobi_scaling=scaling)
    ref = vo.single_step(_oracle(p), radius, scaling)
    assert abs(cost - ref["cost"]) <= 1e-10 * ref["cost"]
    assert abs(model - ref["model"]) <= 1e-7 * abs(ref["model"])
    assert np.abs(cs - ref["cam_step"]).max() <= 1e-6 * np.abs(ref["cam_step"]).max()
    assert np.abs(ps - ref["pt_step"]).max() <= 1e-6 * np.abs(ref["pt_step"]).max()
    # reduced camera system (before the camera LM diagonal is added) and its right-hand side
    cam, rhs, br, bc, bl = P.get_system()
    S = gpu_pkg.env_blocks_to_dense(br, bc, bl, len(cam))
    assert np.array_equal(cam, np.nonzero(_oracle(p).cam_active)[0])
    assert np.abs(S - ref["S_nodamp"]).max() <= 1e-8 * np.abs(ref["S_nodamp"]).max()
    assert np.abs(rhs.ravel() - ref["rhs"]).max() <= 1e-8 * np.abs(ref["rhs"]).max()
    P.close()


def test_lm_matches_oracle(gpu_pkg, problem_A):
    p = problem_A
    q, t, X, s = gpu_pkg.visual_lm(*_args(p))
    pr, info = vo.ceres_lm(_oracle(p))
    assert s["iterations"] == info["iters"]
    assert s["accepted"] == info["accepted"]
    assert abs(s["cost_first"] - info["cost0"]) <= 1e-10 * info["cost0"]
    assert abs(s["cost_last"] - info["cost"]) <= 1e-6 * info["cost"]
    assert np.abs(q - pr.q).max() <= 1e-6 and np.abs(t - pr.t).max() <= 1e-6 and np.abs(X - pr.X).max() <= 1e-6
    # camera 0 is constant (src/lvba_system.cpp:1582-1583)
    assert np.array_equal(q[0], p["q"][0]) and np.array_equal(t[0], p["t"][0])


def test_edge_cases(gpu_pkg, problem_small):
    """Landmarks without a valid plane are skipped and left untouched; a landmark behind a camera gives a
    zero residual (utils.hpp:78); a camera observed twice by one landmark; a camera with no residuals."""
    p = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in problem_small.items()}
    T = p["X"].shape[0]
    p["plane_nd"][::7, :3] = 0.0                       # no valid plane
    p["plane_nd"][3, 0] = np.nan
    # duplicate camera inside landmark 5
    s5 = p["obs_ptr"][5]
    p["obs_cam"][s5 + 1] = p["obs_cam"][s5]
    # put landmark 8 behind its first camera
    c = p["obs_cam"][p["obs_ptr"][8]]
    R = vo.quat_to_rot(p["q"][c][None])[0]
    p["X"][8] = R.T @ (np.array([0.1, 0.1, -2.0]) - p["t"][c])
    pr = _oracle(p)
    P = gpu_pkg.VisualProblem(*_args(p))
    assert abs(P.cost() - pr.cost()) <= 1e-10 * pr.cost()
    cs, ps, model, cost = P.step(1e4)
    ref = vo.single_step(pr, 1e4, True)
    assert np.abs(cs - ref["cam_step"]).max() <= 1e-6 * np.abs(ref["cam_step"]).max()
    assert np.abs(ps - ref["pt_step"]).max() <= 1e-6 * np.abs(ref["pt_step"]).max()
    P.close()
    q, t, X, s = gpu_pkg.visual_lm(*_args(p))
    pr2, info = vo.ceres_lm(_oracle(p))
    assert abs(s["cost_last"] - info["cost"]) <= 1e-6 * info["cost"]
    skipped = ~vo.valid_tracks(p["plane_nd"])
    assert skipped.sum() > 0 and np.array_equal(X[skipped], p["X"][skipped])


def test_invalid_arguments(gpu_pkg, problem_small):
    p = dict(problem_small)
    bad = p["obs_cam"].copy(); bad[0] = 9999
    p["obs_cam"] = bad
    with pytest.raises(gpu_pkg.LvbaError):
        gpu_pkg.VisualProblem(*_args(p))


def test_config_B_single_iteration_cost_match(gpu_pkg):
    """BASELINE config[1]: 500 cameras / 20k tracks — single LM iteration on 1 B200, cost match."""
    p = synth.make_config("B", lidar=False)
    P = gpu_pkg.VisualProblem(*_args(p))
    pr = _oracle(p)
    cs, ps, model, cost = P.step(1e4)
    ref = vo.single_step(pr, 1e4, True)
    assert abs(cost - ref["cost"]) <= 1e-10 * ref["cost"]
    assert abs(model - ref["model"]) <= 1e-7 * abs(ref["model"])
    assert np.abs(cs - ref["cam_step"]).max() <= 1e-6 * np.abs(ref["cam_step"]).max()
    assert np.abs(ps - ref["pt_step"]).max() <= 1e-6 * np.abs(ref["pt_step"]).max()
    qn, tn, Xn = pr.plus(np.concatenate([ref["cam_step"][pr.cam_active].ravel(), ref["pt_step"][pr.tv].ravel()]))
    P.set_state(qn, tn, Xn)
    c2 = pr.cost(qn, tn, Xn)
    assert abs(P.cost() - c2) <= 1e-10 * c2
    P.close()
