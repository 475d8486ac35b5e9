#!/bin/bash
# run M: dense factor without spills, syrk split sweep, set-up laps, bench, visual offline test, ncu --set full of the top kernels
mkdir -p gpurun_out
timeout 60 python tools/solve_once.py 600 30 3 4 1 > gpurun_out/m_first.txt 2>&1; rc=$?
echo "first solve rc=$rc $(tail -1 gpurun_out/m_first.txt)"
if [ $rc -ne 0 ]; then export LVBA_ND_DENSE=0; echo "dense kernel disabled for the rest of this run"; fi
timeout 300 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/m_pytest_nd.txt 2>&1; echo "pytest nd rc=$?"
timeout 600 python -m pytest tests/test_lidar_gpu.py tests/test_visual_gpu.py -x -q > gpurun_out/m_pytest_build.txt 2>&1; echo "pytest build rc=$?"
for sp in 4 8 16 32; do
  echo "syrk split $sp: $(LVBA_SYRK_SPLIT=$sp timeout 120 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-60) | $(LVBA_SYRK_SPLIT=$sp timeout 120 python tools/solve_once.py 2000 30 3 32 5 2>&1 | tail -1 | cut -c1-60)"
done
LVBA_ND_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/m_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/m_ncu1.log 2>&1
timeout 300 python tools/solver_bench.py 2000x30 5000x30 > gpurun_out/m_solver_bench.txt 2>&1
LVBA_SETUP_TIMING=1 timeout 300 python tools/dev_e2e.py C > gpurun_out/m_setup_laps.txt 2>&1
timeout 600 python bench.py > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"
timeout 600 python -m pytest tests/test_zz_offline_gpu.py -x -q > gpurun_out/m_pytest_offline.txt 2>&1; echo "pytest offline rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k 'regex:lidar_build_kernel|lidar_residual_kernel|visual_build_kernel|visual_residual_kernel|nd_dense_factor_kernel|nd_spike_kernel|nd_syrk_kernel|env_factor_la_kernel|env_backsolve_warp_kernel' \
  --launch-skip 60 -c 24 -o gpurun_out/m_full python tools/dev_e2e.py C > gpurun_out/m_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/m_pytest_nd.txt; tail -3 gpurun_out/m_pytest_build.txt; tail -12 gpurun_out/m_pytest_offline.txt; cut -c1-300 gpurun_out/m_solver_bench.txt
grep -h "setup\]\|call 2" gpurun_out/m_setup_laps.txt | tail -24
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/m_bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["device_ms_per_step"], d["e2e"]["ms_call_A"], d["e2e"]["ms_call_B"])
except Exception as e:
    print("bench unreadable", e)
PY
