#!/usr/bin/env python
"""bench.py — LM iterations/s of the LiDAR-visual BA hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C]

A *step* is one LM iteration of the whole hot path on one synthetic problem:
    t_A  one BALM2 pass  = Hessian/gradient build + damped block-LDL^T solve + retraction
                           + residual-only pass + accept/reject          (bavoxel.hpp:686-766)
    t_B  one Ceres-style pass = Jacobians + Schur elimination + reduced solve + back-substitution
                           + candidate cost + accept/reject              (src/lvba_system.cpp:1643)
    metric value = 1 / (t_A + t_B)            [LM iterations / s]
Every timed step starts from the initial state (device-to-device restore) so that it always contains
the Hessian build; inputs are resident in HBM (`value`).  `e2e` is the same metric through the one-shot
C-ABI calls lvba_lidar_lm / lvba_visual_lm with HOST (pinned) buffers: symbolic set-up, H2D upload of the
whole problem, the reference's own iteration caps (10 / 50) and the D2H result copy are inside the timed
region; t_A, t_B there are call time / iterations executed.

Workload at every N: BASELINE.json configs[2] (2000 poses / 200k plane voxels / 100k tracks) unless
--config says otherwise.  N > 1 shards voxels / tracks by contiguous pose-block rows across ranks (strong
scaling of ONE problem, NCCL all-reduce of H, g, S, rhs and the scalar costs).

`--impl reference` times the reference-restated CPU path (oracle/cpu_ref.cpp — the reference itself needs
Eigen/Ceres/PCL/ROS and cannot be built here) on the box's host cores, same config / metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "LM iterations/sec (Jacobian+Hessian build + Schur solve)"
UNIT = "LM iterations/s"
VKEYS = ("q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr", "sigma_px", "sigma_plane")


def workload_name(cfg, p):
    return (f"config {cfg}: {p['n_poses']} poses / {p['n_vox']} plane voxels / {p['n_tracks']} visual tracks, "
            f"synthetic (oracle/synth.py, seed {p['seed']})")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def algorithmic_bytes(p, lid_counts, vis_counts, n_active):
    """SURVEY.md §8(d) formulas evaluated exactly on the generated instance."""
    W = p["n_poses"]
    nnzA, nbH = lid_counts["nnz"], lid_counts["n_blocks_nonzero"]
    build_A = 84 * nnzA + 96 * W + 288 * nbH + 48 * W            # Hessian pass: read slots + poses, write H, g
    resid_A = 84 * nnzA + 96 * W
    nnzB, Tv, nbS = vis_counts["nnz_valid"], vis_counts["n_valid_tracks"], vis_counts["n_blocks_env"]
    bytes_B = 2 * (16 * nnzB + 56 * p["n_poses"] + 56 * Tv) + 288 * nbS + 48 * n_active + 24 * Tv
    return dict(lidar_build=build_A, lidar_residual=resid_A, bytes_A=build_A + resid_A, bytes_B=bytes_B)


def cpu_reference_iteration(p, threads, iters_A=2, iters_B=2):
    """Bounded sample of the CPU restatement: a few LM passes of each path on the full problem."""
    from oracle import cpu_ref
    _, a = cpu_ref.lidar_lm(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"], max_iter=iters_A, rel_tol=-1.0, threads=threads)
    _, _, _, b = cpu_ref.visual_lm(*[p[k] for k in VKEYS], max_iter=iters_B, threads=threads, function_tolerance=-1.0)
    tA = a["ms_build"] / max(a["builds"], 1) + (a["ms_solve"] + a["ms_residual"]) / max(a["iterations"], 1)
    tB = (b["ms_build"] + b["ms_solve"] + b["ms_residual"]) / max(b["iterations"], 1)
    return tA, tB, a, b


def run_reference(args, p, cfg):
    """--impl reference: the CPU path, all host threads, same metric / config."""
    from oracle import cpu_ref
    threads = cpu_ref.hardware_threads()
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        tA, tB, a, b = cpu_reference_iteration(p, threads, 1, 1)
        wall = time.perf_counter() - t0
        if i >= args.warmup:
            times.append((tA + tB) / 1e3)
    mean = sum(times) / len(times)
    val = 1.0 / mean
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": mean * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, p), "l2": "inputs larger than L2 (CPU run)"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "each step = 1 LM pass of path A + 1 of path B on the full problem (oracle/cpu_ref.cpp)"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C", choices=["A", "B", "C", "E"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-voxel-map", action="store_true", help="skip the boundary-B3 (voxel map) side measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from oracle import synth
    if args.impl == "reference":
        if rank != 0:
            return 0
        p = synth.make_config(args.config)
        run_reference(args, p, args.config)
        return 0

    import torch
    import torch.distributed as dist
    import __graft_entry__ as graft
    pkg = graft.load_package()
    pkg.load_library()
    if pkg.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the LVBA hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = [pkg.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pkg.comm_init(world, rank, uid[0], local_rank)

    p = synth.make_config(args.config)                    # identical on every rank (counter-based RNG)
    L = pkg.LidarProblem(p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"], device=local_rank)
    Vz = pkg.VisualProblem(*[p[k] for k in VKEYS], device=local_rank)
    lo_opts = pkg.lidar_default_opts(); lo_opts.rel_tol = -1.0; lo_opts.max_iter = 1 << 30
    vo_opts = pkg.visual_default_opts(); vo_opts.function_tolerance = -1.0; vo_opts.parameter_tolerance = -1.0
    vo_opts.gradient_tolerance = -1.0; vo_opts.max_iter = 1 << 30

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")      # > 126 MB L2

    def one_step():
        """returns (host seconds, summaryA, summaryB) for one LM pass of each path from the initial state"""
        flush.fill_(rank + 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        L.reset_lm(lo_opts); L.reset_state(); sa = L.iterate(1)
        Vz.reset_lm(vo_opts); Vz.reset_state(); sb = Vz.iterate(1)
        return time.perf_counter() - t0, sa, sb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    wall0 = time.perf_counter()
    host_s, dev_ms, launches = 0.0, {"build_A": 0.0, "solve_A": 0.0, "resid_A": 0.0, "build_B": 0.0, "solve_B": 0.0, "resid_B": 0.0}, 0
    for _ in range(args.steps):
        dt, sa, sb = one_step()
        host_s += dt
        dev_ms["build_A"] += sa["ms_build"]; dev_ms["solve_A"] += sa["ms_solve"]; dev_ms["resid_A"] += sa["ms_residual"]
        dev_ms["build_B"] += sb["ms_build"]; dev_ms["solve_B"] += sb["ms_solve"]; dev_ms["resid_B"] += sb["ms_residual"]
        launches += sa["kernel_launches"] + sb["kernel_launches"]
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    # max over ranks of the timed seconds (host clock around synchronous steps) and of the device time
    dev_total_ms = sum(dev_ms.values())
    tt = torch.tensor([host_s, dev_total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    host_s_max, dev_ms_max = tt[0].item(), tt[1].item()
    ms_per_step = host_s_max * 1e3 / args.steps
    value = args.steps / host_s_max

    # ---- full LM to convergence (BASELINE configs[2]) — informational, device resident
    L.reset_lm(None); L.reset_state(); fa = L.iterate(10)
    Vz.reset_lm(None); Vz.reset_state(); fb = Vz.iterate(50)

    # ---- e2e through the one-shot C-ABI calls, host (pinned) buffers
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy()
    hp = {k: pinned(p[k]) for k in ("vox_ptr", "pose_idx", "clusters", "poses", "q", "t", "X", "plane_nd", "obs_ptr", "obs_cam", "obs_uv", "intr")}
    e2e_tA = e2e_tB = 0.0
    h2d = d2h = 0
    n_e2e = max(1, args.e2e_steps)
    for i in range(n_e2e + 1):                                   # first call is warm-up
        barrier()
        t0 = time.perf_counter()
        _, sa = pkg.lidar_lm(hp["vox_ptr"], hp["pose_idx"], hp["clusters"], hp["poses"])
        t1 = time.perf_counter()
        _, _, _, sb = pkg.visual_lm(hp["q"], hp["t"], hp["X"], hp["plane_nd"], hp["obs_ptr"], hp["obs_cam"], hp["obs_uv"], hp["intr"], p["sigma_px"], p["sigma_plane"])
        t2 = time.perf_counter()
        if i > 0:
            e2e_tA += (t1 - t0) / max(sa["iterations"], 1)
            e2e_tB += (t2 - t1) / max(sb["iterations"], 1)
            h2d += sa["h2d_bytes"] + sb["h2d_bytes"]; d2h += sa["d2h_bytes"] + sb["d2h_bytes"]
            e2e_iters = (sa["iterations"], sb["iterations"])
    te = torch.tensor([e2e_tA / n_e2e + e2e_tB / n_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = 1.0 / te.item()

    if rank == 0:
        lc = L.counts(nonzero=True); vc = Vz.counts(); n_active = len(Vz.structure()[0])
        ab = algorithmic_bytes(p, lc, vc, n_active)
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()); peak = float(peaks["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        traffic = None
        try:     # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (same config)
            prof = json.loads((ROOT / "profiles" / "r01_ncu_full_summary.json").read_text())["lidar_build_kernel"]
            if args.config == "C" and world == 1:
                traffic = (float(prof["dram__bytes_read.sum"]) + float(prof["dram__bytes_write.sum"])) * 1e6
        except Exception:
            pass
        build_ms = dev_ms["build_A"] / args.steps
        achieved = ab["lidar_build"] / (build_ms * 1e-3) / 1e9
        roofline = {"kernel": "lidar_build_kernel (+ memset of H, + partial-sum reduce)", "bound": "hbm", "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ab["lidar_build"], "avg_launch_ms": build_ms,
                    "note": "timed with CUDA events on the library's launch stream; the kernel is atomic/FP64 bound, not HBM bound (DESIGN.md)"}
        # the kernel that dominates the step is the block LDL^T of the pose / camera system: a chain of W sequential
        # pivot columns on two SMs (twisted), bound by ONE SM's FP64 pipe, shared-memory wavefronts and the latency of
        # the look-ahead chain -- neither HBM nor tensor cores.  Its useful FP64 work is reported
        # against the FP64 peak of the SMs it can use; the time is the whole solve phase (factorisation + substitutions).
        def ldl_flops(brow, bcol, n):
            below = np.bincount(bcol[brow > bcol], minlength=n).astype(np.float64)     # blocks under every pivot
            return float((below * (below + 1) / 2 * 432 + below * 72 * 2 + below * 72 * 2).sum())   # trailing update + scale + two substitutions
        brA, bcA = L.structure()
        sv = Vz.structure()
        flA = ldl_flops(np.asarray(brA), np.asarray(bcA), p["n_poses"])
        flB = ldl_flops(np.asarray(sv[1]), np.asarray(sv[2]), n_active)
        solve_s = (dev_ms["solve_A"] + dev_ms["solve_B"]) / args.steps * 1e-3
        fp64_two_sm = 2 * 64 * 2 * 1.965e9 / 1e12                                      # 2 SMs x 64 FMA/clk x 2 flop x 1.965 GHz
        dominant = {"kernel": "env_factor_la_kernel<P> (register-window block LDL^T, two CTAs: twisted halves) + env_backsolve_warp_kernel",
                    "share_of_step": (dev_ms["solve_A"] + dev_ms["solve_B"]) / max(dev_total_ms, 1e-9),
                    "bound": "FP64 pipe + shared-memory wavefronts + look-ahead chain latency of ONE SM per half; sequential over pivot columns",
                    "fp64_flops_per_step": flA + flB, "achieved_tflops": (flA + flB) / solve_s / 1e12,
                    "peak_tflops_of_the_two_sms_it_runs_on": fp64_two_sm, "frac": (flA + flB) / solve_s / 1e12 / fp64_two_sm,
                    "note": "FP64 vector peak measured 63.7 FMA/clk/SM (tools/ubench/fp64_rate.cu); time = solve phase (factor + substitutions) from CUDA events"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.config, p), "parallelism": f"pose-block-row shards x{world}",
                           "l2": "256 MiB buffer written between timed steps (L2 flush)",
                           "step": "1 LM pass of path A + 1 of path B from the initial state, inputs resident in HBM"},
                "device_ms_per_step": {k: v / args.steps for k, v in dev_ms.items()},
                "dominant_kernel_by_time": dominant,
                "device_ms_per_step_total_max_over_ranks": dev_ms_max / args.steps,
                "timed_region_wall_s": wall,
                "roofline": roofline, "algorithmic_bytes": ab,
                "full_lm": {"A": {k: fa[k] for k in ("iterations", "accepted", "cost_first", "cost_last", "ms_total")},
                            "B": {k: fb[k] for k in ("iterations", "accepted", "cost_first", "cost_last", "ms_total")}},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d // n_e2e, "d2h_bytes_per_step": d2h // n_e2e,
                        "step": f"one lvba_lidar_lm (<=10 passes, ran {e2e_iters[0]}) + one lvba_visual_lm (<=50 passes, ran {e2e_iters[1]}) from pinned host buffers; t = call time / passes"},
                "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_ref
            threads = cpu_ref.hardware_threads()
            tA, tB, a, b = cpu_reference_iteration(p, threads, 2, 2)
            line["cpu_baseline"] = {"value": 1e3 / (tA + tB), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "2 LM passes of path A + 2 of path B on the full problem (oracle/cpu_ref.cpp), per-pass time with Hessian build",
                                    "ms_A": tA, "ms_B": tB}
        if world == 1 and not args.no_voxel_map:
            # Boundary B3 (set-up stage) measured OUTSIDE the timed region, in a child process so that
            # nothing it does can disturb the numbers above; reported next to them, not part of `value` / `e2e`.
            import subprocess
            try:
                r = subprocess.run([sys.executable, str(ROOT / "tools" / "bench_voxel_map.py")], capture_output=True, text=True, timeout=300)
                last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                line["voxel_map"] = json.loads(last[-1]) if last else {"error": (r.stderr or r.stdout)[-400:], "rc": r.returncode}
            except Exception as e:          # noqa: BLE001 - a side measurement must never take the bench line down
                line["voxel_map"] = {"error": repr(e)[:400]}
        print(json.dumps(line), flush=True)
    L.close(); Vz.close()
    if world > 1:
        pkg.comm_destroy()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
