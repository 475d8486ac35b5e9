#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/dmma > gpurun_out/i_dmma.txt 2>&1
for m in 0 1 2 4 7; do
  LVBA_SPIKE_MODE=$m LVBA_ND_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nd_spike --csv --log-file gpurun_out/i_spike_mode$m.csv python tools/solve_once.py 2000 30 3 16 1 > gpurun_out/i_ncu_m$m.log 2>&1
  echo "mode $m: $(grep nd_spike gpurun_out/i_spike_mode$m.csv | head -2 | awk -F'\",\"' '{print $NF}' | tr '\n' ' ')"
done
cat gpurun_out/i_dmma.txt
