"""Pins the path-B oracle (no GPU): analytic Jacobians vs finite differences along the Ceres manifold,
scipy.optimize.least_squares reaching the same optimum, C++ restatement vs numpy, golden regression."""
from pathlib import Path

import numpy as np
import scipy.optimize

from oracle import cpu_ref, synth, visual_oracle as vo

GOLD = np.load(Path(__file__).parent / "golden" / "small_problem.npz")
KEYS = ("q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr", "sigma_px", "sigma_plane")


def _prob(p):
    return vo.VisualProblem(*[p[k] for k in KEYS])


def test_jacobian_vs_finite_differences(problem_small):
    pr = _prob(problem_small)
    res, J = pr.residuals(jac=True)
    rng = np.random.default_rng(2)
    d = rng.normal(size=pr.ncols) * 1e-3

    def R(dd):
        q, t, X = pr.plus(dd)
        return pr.residuals(q, t, X)[0]
    h = 1e-4
    fd = (R(h * d) - R(-h * d)) / (2 * h)
    assert np.abs(fd - J @ d).max() <= 1e-7 * np.abs(J @ d).max()


def test_plus_jacobian_is_orthonormal_tangent_basis(problem_small):
    q = problem_small["q"]
    PJ = vo.plus_jacobian(q)
    assert np.abs(np.einsum("nij,nik->njk", PJ, PJ) - np.eye(3)).max() <= 1e-12
    assert np.abs(np.einsum("nij,ni->nj", PJ, q)).max() <= 1e-12
    # Plus(q, d) - q ~ PJ d  to first order, and stays unit norm
    d = np.full((q.shape[0], 3), 1e-6)
    qp = vo.manifold_plus(q, d)
    assert np.abs(qp - q - np.einsum("nij,nj->ni", PJ, d)).max() <= 1e-11
    assert np.abs(np.linalg.norm(qp, axis=1) - 1).max() <= 1e-14


def test_behind_camera_residual_is_zero():
    """utils.hpp:78: Xc.z <= 1e-8 -> residual (0,0) and zero Jacobian."""
    q = np.array([[1.0, 0, 0, 0]]); t = np.zeros((1, 3)); X = np.array([[0.1, 0.2, -1.0]])
    r, Jq, Jt, JX = vo.reproj_eval(q, t, X, np.zeros((1, 2)), synth.INTR, 0.5)
    assert np.all(r == 0) and np.all(Jq == 0) and np.all(Jt == 0) and np.all(JX == 0)


def test_lm_reaches_least_squares_optimum():
    p = synth.make_problem(8, 0, 40, seed=9, lidar=False)
    pr = _prob(p)
    pr, info = vo.ceres_lm(pr, max_iter=50)

    pr0 = _prob(p)

    def fun(x):
        q, t, X = pr0.plus(x)
        return pr0.residuals(q, t, X)[0]
    sol = scipy.optimize.least_squares(fun, np.zeros(pr0.ncols), method="trf", x_scale="jac", xtol=1e-14, ftol=1e-14, gtol=1e-12)
    assert abs(info["cost"] - sol.cost) <= 1e-5 * sol.cost      # Ceres stops at function tolerance 1e-6
    assert info["cost"] >= sol.cost * (1 - 1e-12)


def test_cpp_restatement_matches_numpy(problem_small):
    p = problem_small
    q, t, X, info = cpu_ref.visual_lm(*[p[k] for k in KEYS], threads=3)
    pr, inf = vo.ceres_lm(_prob(p))
    assert int(info["iterations"]) == inf["iters"] and int(info["accepted"]) == inf["accepted"]
    assert abs(info["cost_first"] - inf["cost0"]) <= 1e-10 * inf["cost0"]
    assert abs(info["cost_last"] - inf["cost"]) <= 1e-9 * inf["cost"]
    assert np.abs(q - pr.q).max() <= 1e-8 and np.abs(X - pr.X).max() <= 1e-8


def test_schur_equals_full_normal_equations(problem_small):
    pr = _prob(problem_small)
    st = vo.single_step(pr, 1e4, True)
    y_c = np.linalg.solve(st["S"], st["rhs"])
    assert np.abs(y_c - st["y"][:6 * pr.nc]).max() <= 1e-8 * np.abs(y_c).max()


def test_golden_fixture_regression():
    g = GOLD
    pr = vo.VisualProblem(*[g[k] if k not in ("sigma_px", "sigma_plane") else float(g[k]) for k in KEYS])
    st = vo.single_step(pr, 1e4, True)
    assert abs(st["cost"] - float(g["B_cost0"])) <= 1e-12 * st["cost"]
    assert np.abs(st["cam_step"] - g["B_cam_step"]).max() <= 1e-9 * np.abs(g["B_cam_step"]).max()
    q, t, X, info = cpu_ref.visual_lm(*[g[k] if k not in ("sigma_px", "sigma_plane") else float(g[k]) for k in KEYS], threads=2)
    assert abs(info["cost_last"] - float(g["B_lm_cost"])) <= 1e-9 * float(g["B_lm_cost"])
    assert int(info["iterations"]) == int(g["B_lm_iters"])
