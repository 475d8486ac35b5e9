"""world_size-2 gloo rehearsal (CPU) of the round-2 multi-GPU data flow of one LM pass of path A (SURVEY.md
section 8(e)), with REAL inter-process collectives where the device code calls NCCL:

  rows owned per rank (nd_plan.h, what lvba_lidar_owned_rows reports)  ->  a voxel belongs to the owner of its LOWEST pose row
  ->  every rank builds H only from its own voxels  ->  the rows that spill over into the right neighbour's range (at most
  max_col of them) travel by send / recv and are added there (Solver::exchange_rows, csrc/runtime.cuh)  ->  g, the damping
  diagonal and the cost by small all-reduces  ->  every rank eliminates its own subtree from ITS rows (all foreign rows are NaN),
  ONE all-gather of the fixed-size slots  ->  replicated top of the tree, downward sweep  ->  x from the owned rows by an
  all-reduce.  No all-reduce of H anywhere.

The plan builder, the job tables and every layout pass are the library's own code (tests/emu/nd_emu.cpp instantiates them with
the host policy); the per-voxel arithmetic is the numpy oracle.  Rank 0 compares the step with a dense solve of the full system."""
import ctypes
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
W, V, P_WANT = 420, 3200, 4


def _envelope_storage(M, first, row_start):
    n = len(first)
    L = np.zeros((row_start[-1], 36))
    for r in range(n):
        for c in range(first[r], r + 1):
            L[row_start[r] + c - first[r]] = M[6 * r:6 * r + 6, 6 * c:6 * c + 6].ravel()
    return L


def _worker(rank, world, port, so_path, q):
    try:
        _rank_pass(rank, world, port, so_path, q)
    except BaseException as e:      # noqa: BLE001 — the parent must not wait for a result that will never come
        import traceback
        q.put({f"error_rank{rank}": traceback.format_exc()[-1500:]})
        raise e


def _rank_pass(rank, world, port, so_path, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import lidar_oracle as lo, synth
    from test_nd_solver_emu import envelope
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emu = ctypes.CDLL(so_path)
    c = lambda a, t: a.ctypes.data_as(ctypes.POINTER(t))

    p = synth.make_problem(W, V, 0, seed=33, visual=False)
    vp, pi, cl, poses = p["vox_ptr"], p["pose_idx"], p["clusters"], p["poses"]
    lowest = np.minimum.reduceat(pi, vp[:-1])
    first_raw = np.arange(W)
    for a in range(V):                                   # Envelope::build: a row reaches back to the lowest pose of its voxels
        sl = pi[vp[a]:vp[a + 1]]
        first_raw[sl] = np.minimum(first_raw[sl], lowest[a])
    first, rs = envelope(first_raw)
    max_col = int(np.max(np.arange(W) - first))
    rb, re = np.zeros(world, np.int32), np.zeros(world, np.int32)
    p_used = emu.nd_emu_plan_rows(W, c(first, ctypes.c_int), P_WANT, world, c(rb, ctypes.c_int), c(re, ctypes.c_int))
    assert p_used >= world, p_used
    assert rb[0] == 0 and re[-1] == W and all(re[r] == rb[r + 1] for r in range(world - 1))

    # ---- build: own voxels only
    mine = np.nonzero((lowest >= rb[rank]) & (lowest < re[rank]))[0]
    sl = np.concatenate([np.arange(vp[a], vp[a + 1]) for a in mine])
    lvp = np.concatenate([[0], np.cumsum(vp[mine + 1] - vp[mine])]).astype(np.int64)
    cost, g, blocks = lo.acc_evaluate2(lvp, pi[sl], cl[sl], poses, W)
    Hloc = _envelope_storage(lo.assemble_dense(blocks, W), first, rs)
    touched = np.nonzero([np.any(Hloc[rs[r]:rs[r + 1]] != 0) for r in range(W)])[0]
    spill_end = min(W, re[rank] + max_col)
    assert touched.min() >= rb[rank] and touched.max() < spill_end, (touched.min(), touched.max(), rb[rank], spill_end)

    # ---- exchange_rows: spill-over rows to the right neighbour, added there
    reqs = []
    if rank + 1 < world:
        send = torch.from_numpy(Hloc[rs[re[rank]]:rs[spill_end]].copy())
        reqs.append(dist.isend(send, rank + 1))
    if rank > 0:
        lo_r, hi_r = rb[rank], min(W, rb[rank] + max_col)
        recv = torch.zeros((rs[hi_r] - rs[lo_r], 36), dtype=torch.float64)
        dist.recv(recv, rank - 1)
        Hloc[rs[lo_r]:rs[hi_r]] += recv.numpy()
    for rq in reqs:
        rq.wait()
    payload = (rs[spill_end] - rs[re[rank]]) * 288 if rank + 1 < world else 0

    # ---- small all-reduces: g, cost, damping diagonal (owned rows only)
    gt = torch.from_numpy(np.asarray(g, np.float64).ravel().copy()); ct = torch.tensor([cost, float(len(mine))], dtype=torch.float64)
    diag = np.zeros(6 * W)
    for r in range(rb[rank], re[rank]):
        diag[6 * r:6 * r + 6] = Hloc[rs[r + 1] - 1].reshape(6, 6).diagonal()
    dt = torch.from_numpy(diag)
    dist.all_reduce(gt); dist.all_reduce(ct); dist.all_reduce(dt)
    u = 0.01
    dadd = u * dt.numpy()
    rhs = -gt.numpy()

    # ---- reference on every rank: the whole system from all voxels
    cost0, g0, b0 = lo.acc_evaluate2(vp, pi, cl, poses, W)
    H0 = lo.assemble_dense(b0, W)
    Hfull = _envelope_storage(H0, first, rs)
    own = slice(rs[rb[rank]], rs[re[rank]])
    err_rows = float(np.abs(Hloc[own] - Hfull[own]).max() / np.abs(Hfull).max())
    Hloc[:rs[rb[rank]]] = np.nan; Hloc[rs[re[rank]]:] = np.nan          # what is not owned is not there

    # ---- solve: own subtree, all-gather of the slots, top tree, all-reduce of x
    region = ctypes.POINTER(ctypes.c_double)()
    slot = ctypes.c_longlong(0)
    Hc, dc, rc_ = np.ascontiguousarray(Hloc), np.ascontiguousarray(dadd), np.ascontiguousarray(rhs)
    used = emu.nd_emu_rank_up(W, c(first, ctypes.c_int), c(Hc, ctypes.c_double), c(dc, ctypes.c_double), c(rc_, ctypes.c_double), P_WANT, world, rank,
                              ctypes.byref(region), ctypes.byref(slot))
    assert used == p_used, (used, p_used)
    reg = np.ctypeslib.as_array(region, shape=(world * slot.value,))
    mine_slot = torch.from_numpy(reg[rank * slot.value:(rank + 1) * slot.value].copy())
    gathered = [torch.zeros_like(mine_slot) for _ in range(world)]
    dist.all_gather(gathered, mine_slot)
    for r in range(world):
        if r != rank:
            reg[r * slot.value:(r + 1) * slot.value] = gathered[r].numpy()
    x = np.full(6 * W, np.nan)
    bad = emu.nd_emu_rank_down(c(x, ctypes.c_double))
    assert bad == 0
    assert np.all(x[:6 * rb[rank]] == 0) and np.all(x[6 * re[rank]:] == 0)
    xt = torch.from_numpy(x)
    dist.all_reduce(xt)

    if rank == 0:
        xr = np.linalg.solve(H0 + np.diag(u * H0.diagonal()), -np.asarray(g0).ravel())
        q.put({"err_rows": err_rows, "err_x": float(np.abs(xt.numpy() - xr).max() / np.abs(xr).max()), "err_g": float(np.abs(gt.numpy() - np.asarray(g0).ravel()).max() / np.abs(g0).max()),
               "err_cost": abs(ct[0].item() - cost0) / cost0, "n_owned": ct[1].item(), "payload": int(payload), "slot_bytes": int(slot.value * 8),
               "H_bytes": int(rs[-1] * 288), "max_col": max_col, "rows": [int(rb[0]), int(re[0]), int(rb[1]), int(re[1])]})
    else:
        q.put({"err_rows_rank1": err_rows})
    dist.destroy_process_group()


def test_two_rank_row_owned_pass_matches_dense_solve(tmp_path):
    so = tmp_path / "libnd_emu.so"
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-fPIC", "-shared", str(ROOT / "tests" / "emu" / "nd_emu.cpp"), "-o", str(so)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, str(so), q)) for rk in range(2)]
    for pr in procs:
        pr.start()
    res = {}
    for _ in range(2):
        res.update(q.get(timeout=240))
        assert not any(k.startswith("error") for k in res), res
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert res["n_owned"] == V                                   # every voxel built by exactly one rank
    assert res["err_rows"] <= 1e-12 and res["err_rows_rank1"] <= 1e-12, res     # owned rows of H after the neighbour exchange == the full build
    assert res["err_g"] <= 1e-12 and res["err_cost"] <= 1e-12, res
    assert res["err_x"] <= 1e-9, res                             # the damped step of the sharded pass == dense solve
    # what travels: max_col boundary rows + one slot per rank, a small fraction of H (an all-reduce of H would move all of it)
    assert res["payload"] + res["slot_bytes"] < 0.35 * res["H_bytes"], res
