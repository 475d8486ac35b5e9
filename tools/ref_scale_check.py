"""The reference's own source against the CPU arm of bench.py at BASELINE's sizes (run here, where /root/reference exists).

tests/test_ref_pin.py pins the oracles against oracle/_ref/libbalm_ref.so (the reference's BALM headers compiled where they
lie, oracle/ref_driver.cpp) at sizes a test suite can afford.  This tool repeats the comparison at the configurations the
bench quotes, once, and its output is committed (profiles/r02_ref_pin_scale_B.txt, _C.txt):

  config B (500 poses / 50 000 voxels):   BALM2::divide_thread (H, g, residual) and the whole BALM2::damping_iter
                                           vs oracle/cpu_ref.cpp (`--impl reference` / cpu_baseline of bench.py)
  config C (2000 poses / 200 000 voxels): VOX_HESS::acc_evaluate2 (H, g, residual) vs oracle/cpu_ref.cpp.  The reference
                                           keeps vector<PointCluster>(win_size) per voxel (42 GB at this size) and a dense
                                           12000 x 12000 Hessian per thread, so the voxels go through acc_evaluate2 in slices
                                           (what divide_thread does with its 16 parts) in a few child processes and the partial
                                           Hessians are added here.  damping_iter itself (19 dense 1.15 GB matrices and a
                                           144 M-entry triplet loop per pass) is not run at this size.

bench.py's `parity_C` gate then ties the CUDA path to that same CPU arm at config C on the GPU box.
  V: path B at config C — cost and Jacobians from the reference's own cost functors vs the port / the numpy oracle (profiles/r02_ref_pin_scale_V.txt)
  R: the whole damping_iter on 24 random small problems, reference source vs numpy oracle vs port (profiles/r02_ref_pin_scale_R.txt)
TEST INFRASTRUCTURE: reads oracle/, never the product.    python tools/ref_scale_check.py [B] [C] [R]
"""
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import balm_ref, cpu_ref, synth  # noqa: E402


def blocks_vs_dense(br_, bc, bl, H, W):
    """max |port block - reference block| over the port's block list, and the largest reference entry outside it."""
    err = 0.0
    seen = np.zeros((W, W), bool)
    for k in range(len(br_)):
        i, j = int(br_[k]), int(bc[k])
        seen[i, j] = seen[j, i] = True
        err = max(err, float(np.abs(H[6 * i:6 * i + 6, 6 * j:6 * j + 6] - bl[k]).max()))
    out = 0.0
    Hb = np.abs(H).reshape(W, 6, W, 6).max(axis=(1, 3))
    if (~seen).any():
        out = float(Hb[~seen].max())
    return err, out


def config_b(log):
    p = synth.make_config("B", visual=False)
    a = (p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"])
    W, V = 500, len(p["vox_ptr"]) - 1
    t = time.time(); res, g, H, kept = balm_ref.lidar_hessian(*a, threads=True); t_ref = time.time() - t
    t = time.time(); r, g_p, br_, bc, bl = cpu_ref.lidar_build(*a, threads=8); t_port = time.time() - t
    err, out = blocks_vs_dense(br_, bc, bl, H, W)
    sc = float(np.abs(H).max())
    log(f"config B build: divide_thread (reference source, 16 threads) {t_ref:.1f} s, port (8 threads) {t_port:.2f} s; kept {kept} of {V}")
    log(f"  residual: reference sum/kept {res:.15e}  port {r:.15e} (or x kept)  rel {min(abs(r - res), abs(r - res * kept) / kept) / abs(res):.2e}")
    log(f"  g   max|diff| / max|g| = {np.abs(g - g_p).max() / np.abs(g).max():.2e}")
    log(f"  H   max|diff| / max|H| = {err / sc:.2e} over {len(br_)} blocks; largest reference entry outside the port's block list {out:.1e}")
    t = time.time(); lm_ref = balm_ref.lidar_damping_iter(*a); t_ref = time.time() - t
    t = time.time(); lm_port, info = cpu_ref.lidar_lm(*a, threads=8); t_port = time.time() - t
    log(f"config B damping_iter: reference source {t_ref:.1f} s, port {t_port:.2f} s ({int(info['iterations'])} passes, {int(info['accepted'])} accepted)")
    log(f"  end poses max|diff| = {np.abs(lm_ref - lm_port).max():.2e}   (moved by {np.abs(lm_ref - p['poses']).max():.3f})")
    r_ref = balm_ref.lidar_residual(p["vox_ptr"], p["pose_idx"], p["clusters"], lm_ref) / kept
    log(f"  end cost: reference {r_ref:.15e}  port {info['cost_last']:.15e}  rel {abs(r_ref - info['cost_last']) / r_ref:.2e}")


def _slice(args):
    lo_, hi, seed_cfg = args
    p = synth.make_config(seed_cfg, visual=False)
    vp = p["vox_ptr"]
    sel = slice(int(vp[lo_]), int(vp[hi]))
    res, g, H, kept = balm_ref.lidar_hessian((vp[lo_:hi + 1] - vp[lo_]).astype(np.int64), p["pose_idx"][sel], p["clusters"][sel], p["poses"])
    # return the block-max map and the blocks themselves lazily: the dense H is 1.15 GB, keep it in a shared file
    path = f"/tmp/ref_scale_H_{lo_}.npy"
    np.save(path, H)
    return res, g, kept, path


def config_c(log, procs=4, parts=16):
    p = synth.make_config("C", visual=False)
    W, V = 2000, len(p["vox_ptr"]) - 1
    edges = [int(round(V * k / parts)) for k in range(parts + 1)]        # divide_thread's split: part * i .. part * (i + 1)
    t = time.time()
    with mp.get_context("fork").Pool(procs) as pool:
        outs = pool.map(_slice, [(edges[k], edges[k + 1], "C") for k in range(parts)], chunksize=1)
    t_ref = time.time() - t
    res = sum(o[0] for o in outs); g = sum(o[1] for o in outs); kept = sum(o[2] for o in outs)
    H = None
    for o in outs:
        part = np.load(o[3]); Path(o[3]).unlink()
        H = part if H is None else H + part
    t = time.time(); r, g_p, br_, bc, bl = cpu_ref.lidar_build(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"], threads=8); t_port = time.time() - t
    err, out = blocks_vs_dense(br_, bc, bl, H, W)
    sc = float(np.abs(H).max())
    log(f"config C build: acc_evaluate2 (reference source) over {parts} slices in {procs} processes {t_ref:.0f} s, port (8 threads) {t_port:.2f} s; kept {kept} of {V}")
    log(f"  residual: reference sum {res:.15e}  port {r:.15e} (or x kept)  rel {min(abs(r - res), abs(r * kept - res)) / abs(res):.2e}")
    log(f"  g   max|diff| / max|g| = {np.abs(g - g_p).max() / np.abs(g).max():.2e}")
    log(f"  H   max|diff| / max|H| = {err / sc:.2e} over {len(br_)} blocks; largest reference entry outside the port's block list {out:.1e}")
    # the first damped step (u = 0.01, D = diag H: bavoxel.hpp:692-710) solved from the REFERENCE SOURCE's H and g (sparse LU) vs the port's own step
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rows, cols, vals = [], [], []
    for k in range(len(br_)):
        i, j = int(br_[k]), int(bc[k])
        for (a, b_) in ((i, j), (j, i)) if i != j else ((i, j),):
            blk = H[6 * a:6 * a + 6, 6 * b_:6 * b_ + 6]
            r_, c_ = np.meshgrid(np.arange(6 * a, 6 * a + 6), np.arange(6 * b_, 6 * b_ + 6), indexing="ij")
            rows.append(r_.ravel()); cols.append(c_.ravel()); vals.append(blk.ravel())
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=H.shape)
    A = A + 0.01 * sp.diags(A.diagonal())
    dx_ref = spla.splu(A).solve(-g.ravel())
    dx_port, _ = cpu_ref.lidar_step(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"], u=0.01, threads=8)
    log(f"  first damped step: |dx(reference-source H, g) - dx(port)| / max|dx| = {np.abs(dx_ref - dx_port.ravel()).max() / np.abs(dx_ref).max():.2e}"
        f"   (north star: 1e-8 per pose update; backward error of the reference-side solve {np.abs(A @ dx_ref + g.ravel()).max() / np.abs(g).max():.1e})")


def config_c_visual(log):
    """Path B at config C: the cost Ceres would evaluate, from the reference's own cost functors (include/utils.hpp compiled where it lies), against the
    cost the CPU arm starts its LM from; and the functor Jacobians (Jets) against the numpy oracle's analytic ones on every observation."""
    from oracle import visual_oracle as vis
    p = synth.make_config("C")                                   # the bench's problem (LiDAR part generated too: the visual part follows it in the seed's stream)
    trk = np.repeat(np.arange(len(p["obs_ptr"]) - 1), np.diff(p["obs_ptr"]))
    t = time.time()
    r, J = balm_ref.reproj(p["q"][p["obs_cam"]], p["t"][p["obs_cam"]], p["X"][trk], p["obs_uv"], p["intr"], p["sigma_px"], p["sigma_px"])
    rp, Jp = balm_ref.point_plane(p["X"], p["plane_nd"], p["sigma_plane"])
    t_ref = time.time() - t
    tv = vis.valid_tracks(p["plane_nd"])
    cost_ref = 0.5 * (float((r[tv[trk]] ** 2).sum()) + float((rp[tv] ** 2).sum()))
    _, _, _, info = cpu_ref.visual_lm(p["q"], p["t"], p["X"], p["plane_nd"], p["obs_ptr"], p["obs_cam"], p["obs_uv"], p["intr"], p["sigma_px"], p["sigma_plane"],
                                      max_iter=1, threads=8)
    log(f"config C path B: {len(trk)} observations / {len(p['X'])} tracks through the reference's functors (T = double and T = Jet) in {t_ref:.1f} s")
    log(f"  cost 1/2 sum r^2: reference functors {cost_ref:.15e}  port {info['cost_first']:.15e}  rel {abs(cost_ref - info['cost_first']) / cost_ref:.2e}")
    ro, Jq, Jt, JX = vis.reproj_eval(p["q"][p["obs_cam"]], p["t"][p["obs_cam"]], p["X"][trk], np.asarray(p["obs_uv"], np.float32).astype(np.float64), p["intr"], p["sigma_px"])
    qn = p["q"][p["obs_cam"]]; qn = qn / np.linalg.norm(qn, axis=1, keepdims=True)
    Jq_ref = J[:, :, :4] @ vis.plus_jacobian(qn)
    sc = max(np.abs(J).max(), 1.0)
    log(f"  residuals vs numpy {np.abs(ro - r).max() / np.abs(r).max():.2e}; Jacobians (Jet vs analytic) d/dq {np.abs(Jq - Jq_ref).max() / sc:.2e}  d/dt {np.abs(Jt - J[:, :, 4:7]).max() / sc:.2e}  d/dX {np.abs(JX - J[:, :, 7:]).max() / sc:.2e}")


def random_sweep(log, n=24, seed=123):
    """Whole damping_iter on random problems (4..70 poses, 3..30 voxels per pose): reference source vs numpy oracle vs C++ port."""
    from oracle import lidar_oracle as lo
    rng = np.random.default_rng(seed)
    worst = 0.0
    for _ in range(n):
        W = int(rng.integers(4, 70)); V = int(rng.integers(3 * W, 30 * W)); sd = int(rng.integers(1, 10 ** 6))
        p = synth.make_problem(W, V, 0, seed=sd, visual=False)
        a = (p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"])
        ref = balm_ref.lidar_damping_iter(*a)
        ora, info = lo.damping_iter(*a)
        port, _ = cpu_ref.lidar_lm(*a, threads=4)
        d1, d2 = float(np.abs(ref - ora).max()), float(np.abs(ref - port).max())
        worst = max(worst, d1, d2)
        log(f"  W {W:3d} V {V:5d} seed {sd:6d}: {info['iters']:2d} passes, {info['accepted']} accepted; end poses |ref - numpy| {d1:.1e}  |ref - port| {d2:.1e}  (moved {np.abs(ref - p['poses']).max():.3f})")
    log(f"random sweep: worst end-pose difference {worst:.1e} over {n} problems (north star: final cost 1e-6)")


if __name__ == "__main__":
    assert balm_ref.available(), "needs oracle/_ref/libbalm_ref.so (make -C oracle ref, where /root/reference exists)"
    which = [a.upper() for a in sys.argv[1:]] or ["B", "C"]
    lines = []

    def log(s):
        print(s, flush=True); lines.append(s)

    log("reference source (oracle/_ref/libbalm_ref.so: /root/reference/include/BALM/*.hpp on the stand-in headers of oracle/ref_shim) vs oracle/cpu_ref.cpp")
    if "B" in which:
        config_b(log)
    if "C" in which:
        config_c(log)
    if "V" in which:
        config_c_visual(log)
    if "R" in which:
        random_sweep(log)
    out = ROOT / "profiles" / ("r02_ref_pin_scale_" + "".join(which) + ".txt")
    out.write_text("\n".join(lines) + "\n")
    print("wrote", out)
