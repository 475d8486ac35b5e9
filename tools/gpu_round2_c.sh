#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/c_pytest_nd.txt 2>&1; echo "nd pytest (dense sep) rc=$?"
LVBA_ND_DENSE=0 timeout 900 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/c_pytest_nd_nodense.txt 2>&1; echo "nd pytest (register-window sep) rc=$?"
LVBA_ND_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/c_ncu1.log 2>&1
LVBA_ND_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c_launches_2000_32.csv python tools/solve_once.py 2000 30 3 32 2 > gpurun_out/c_ncu2.log 2>&1
timeout 600 python tools/solver_bench.py > gpurun_out/c_solver_bench.txt 2> gpurun_out/c_solver_bench.err; echo "solver bench rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-voxel-map > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/c_pytest_nd.txt; tail -3 gpurun_out/c_pytest_nd_nodense.txt; cat gpurun_out/c_solver_bench.txt; head -c 1200 gpurun_out/c_bench.json
