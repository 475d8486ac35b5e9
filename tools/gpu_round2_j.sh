#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/j_pytest.txt 2>&1; echo "pytest nd rc=$?"
LVBA_ND_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/j_ncu1.log 2>&1
for m in 1 7; do
  LVBA_SPIKE_MODE=$m LVBA_ND_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nd_spike --csv --log-file gpurun_out/j_spike_mode$m.csv python tools/solve_once.py 2000 30 3 16 1 > gpurun_out/j_ncu_m$m.log 2>&1
  echo "mode $m: $(grep nd_spike gpurun_out/j_spike_mode$m.csv | head -2 | awk -F'\",\"' '{print $NF}' | tr '\n' ' ')"
done
timeout 600 python tools/solver_bench.py 2000x30 2000x20 5000x30 > gpurun_out/j_solver_bench.txt 2>&1
tail -3 gpurun_out/j_pytest.txt; cut -c1-330 gpurun_out/j_solver_bench.txt
