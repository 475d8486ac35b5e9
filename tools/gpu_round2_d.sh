#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/d_pytest_nd.txt 2>&1; echo "nd pytest rc=$?"
LVBA_ND_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/d_ncu1.log 2>&1
LVBA_ND_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_2000_32.csv python tools/solve_once.py 2000 30 3 32 2 > gpurun_out/d_ncu2.log 2>&1
LVBA_ND_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nd_spike|nd_dense|nd_syrk" -c 6 -o gpurun_out/d_full python tools/solve_once.py 2000 30 3 16 1 > gpurun_out/d_ncu3.log 2>&1
timeout 600 python tools/solver_bench.py > gpurun_out/d_solver_bench.txt 2> gpurun_out/d_solver_bench.err; echo "solver bench rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-voxel-map > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/d_pytest_nd.txt; cat gpurun_out/d_solver_bench.txt | cut -c1-600; head -c 1000 gpurun_out/d_bench.json
