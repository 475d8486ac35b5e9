#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/b_pytest_nd.txt 2>&1; echo "nd pytest rc=$?"
LVBA_ND_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/b_ncu1.log 2>&1
LVBA_ND_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches_5000_32.csv python tools/solve_once.py 5000 30 3 32 2 > gpurun_out/b_ncu2.log 2>&1
LVBA_ND_GRAPH=0 python tools/solve_once.py 2000 30 3 16 5 > gpurun_out/b_nograph.txt 2>&1
python tools/solve_once.py 2000 30 3 16 5 > gpurun_out/b_graph.txt 2>&1
tail -5 gpurun_out/b_pytest_nd.txt; cat gpurun_out/b_nograph.txt gpurun_out/b_graph.txt; tail -2 gpurun_out/b_ncu1.log
