#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/solve_once.py 600 30 3 4 1 > gpurun_out/k_first.txt 2>&1; echo "first dense-v3 solve rc=$? $(tail -1 gpurun_out/k_first.txt)"
timeout 600 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/k_pytest.txt 2>&1; echo "pytest nd rc=$?"
LVBA_ND_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/k_ncu1.log 2>&1
timeout 600 python tools/solver_bench.py 2000x30 2000x20 5000x30 > gpurun_out/k_solver_bench.txt 2>&1
timeout 900 python -m pytest tests/test_zz_fuse_gpu.py -x -q > gpurun_out/k_pytest_fuse.txt 2>&1; echo "pytest fuse rc=$?"
tail -3 gpurun_out/k_pytest.txt; tail -5 gpurun_out/k_pytest_fuse.txt; cut -c1-330 gpurun_out/k_solver_bench.txt
