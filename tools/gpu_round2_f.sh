#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/f_pytest.txt 2>&1; echo "pytest nd rc=$?"
LVBA_ND_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/f_ncu1.log 2>&1
LVBA_SETUP_TIMING=1 timeout 600 python tools/dev_e2e.py C > gpurun_out/f_e2e.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-voxel-map > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/f_pytest.txt; tail -30 gpurun_out/f_e2e.txt
