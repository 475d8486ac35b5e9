"""BASELINE.json configs[3]: Hessian-build sweep on the 2000-pose problem — achieved algorithmic GB/s of
lidar_build_kernel / lidar_residual_kernel against the measured HBM peak as the number of poses per voxel (K)
changes at constant slot count (nnz ~ 1.4 M, the config-C volume).  GPU box; writes a markdown table to stdout.
    python tools/build_sweep.py [--out gpurun_out/build_sweep.md]"""
import argparse, json, statistics, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import __graft_entry__ as graft
from oracle import synth

ap = argparse.ArgumentParser(); ap.add_argument("--out", default=""); ap.add_argument("--ks", default="2,3,4,6,8,12,16"); args = ap.parse_args()
KS = tuple(int(k) for k in args.ks.split(","))
W, NNZ = 2000, 1_400_000


def generate(K):
    """One sweep point: the config-C generator (oracle/synth.py) at a fixed number of poses per voxel."""
    V = NNZ // K
    rng = np.random.Generator(np.random.Philox(key=20260923 + K))
    R_gt, p_gt = synth.make_trajectory(W, rng)
    R0 = R_gt @ synth.so3_exp(rng.normal(0, 0.005, (W, 3))); p0 = p_gt + rng.normal(0, 0.03, (W, 3))
    vp, pi, cl = synth.make_lidar(W, V, R_gt, p_gt, rng, k_lo=K, k_hi=K)
    return K, V, vp, pi, cl, np.concatenate([R0.reshape(W, 9), p0], 1)


# the generator simulates every point of every cluster (20-40 s per sweep point on one core): all points at once, in child
# processes forked BEFORE the CUDA library is loaded
import multiprocessing as mp
with mp.get_context("fork").Pool(len(KS)) as pool:
    problems = pool.map(generate, KS)
pkg = graft.load_package(); pkg.load_library()
try:
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]); src = "measured"
except Exception:
    peak, src = 6650.0, "fallback"
rows = []
for K, V, vp, pi, cl, poses in problems:
    L = pkg.LidarProblem(vp, pi, cl, poses)
    o = pkg.lidar_default_opts(); o.rel_tol = -1.0; o.max_iter = 1 << 30
    tb, tr = [], []
    for i in range(8):
        L.reset_lm(o); L.reset_state(); s = L.iterate(1)
        if i >= 3: tb.append(s["ms_build"]); tr.append(s["ms_residual"])
    c = L.counts(nonzero=True)
    nnz, nbH = c["nnz"], c["n_blocks_nonzero"]
    bytes_build = 84 * nnz + 96 * W + 288 * nbH + 48 * W
    bytes_res = 84 * nnz + 96 * W
    mb, mr = statistics.median(tb), statistics.median(tr)
    rows.append((K, V, nnz, c["n_pairs"], nbH, mb, bytes_build / mb / 1e6, bytes_build / mb / 1e6 / peak, mr, bytes_res / mr / 1e6, bytes_res / mr / 1e6 / peak))
    L.close()
out = [f"# Hessian-build sweep, {W} poses, ~{NNZ/1e6:.1f} M slots, 1 B200 (HBM peak {peak:.0f} GB/s, {src}); times = CUDA events on the library stream, median of 5",
       "# build = memset(H,g) + lidar_build_kernel + partial-sum reduce ; residual = retraction + lidar_residual_kernel + reduce",
       "", "| K (poses / voxel) | voxels | slots | pose pairs | non-zero H blocks | build ms | build GB/s | of HBM peak | residual ms | residual GB/s | of HBM peak |",
       "|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
for r in rows:
    out.append(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]} | {r[4]} | {r[5]:.3f} | {r[6]:.0f} | {100*r[7]:.1f} % | {r[8]:.3f} | {r[9]:.0f} | {100*r[10]:.1f} % |")
txt = "\n".join(out) + "\n"
print(txt)
if args.out: Path(args.out).write_text(txt)
