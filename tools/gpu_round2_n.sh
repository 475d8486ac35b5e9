#!/bin/bash
# run N: spike kernels beside the factorisations (PDL), Newton residual, offline visual test, ncu --set full of the build kernels
mkdir -p gpurun_out
timeout 60 python tools/solve_once.py 600 30 3 4 1 > gpurun_out/n_first.txt 2>&1; rc=$?
echo "first pipelined solve rc=$rc $(tail -1 gpurun_out/n_first.txt | cut -c1-150)"
if [ $rc -ne 0 ]; then export LVBA_ND_PIPELINE=0; echo "pipeline disabled for the rest of this run"; fi
for pl in 0 1; do
  echo "pipeline $pl: $(LVBA_ND_PIPELINE=$pl timeout 100 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-110) | p32 $(LVBA_ND_PIPELINE=$pl timeout 100 python tools/solve_once.py 2000 30 3 32 5 2>&1 | tail -1 | cut -c1-40) | n5000 $(LVBA_ND_PIPELINE=$pl timeout 100 python tools/solve_once.py 5000 30 3 32 5 2>&1 | tail -1 | cut -c1-40) | b20 $(LVBA_ND_PIPELINE=$pl timeout 100 python tools/solve_once.py 2000 20 3 16 5 2>&1 | tail -1 | cut -c1-40)"
done
echo "no graph: $(LVBA_ND_GRAPH=0 timeout 100 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-110)"
timeout 300 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/n_pytest_nd.txt 2>&1; echo "pytest nd rc=$?"
timeout 600 python -m pytest tests/test_lidar_gpu.py tests/test_config_c_gpu.py -x -q > gpurun_out/n_pytest_build.txt 2>&1; echo "pytest lidar/configC rc=$?"
timeout 600 python bench.py --no-voxel-map > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"
timeout 600 python -m pytest tests/test_zz_offline_gpu.py -x -q > gpurun_out/n_pytest_offline.txt 2>&1; echo "pytest offline rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on \
  -k 'regex:lidar_build_kernel|lidar_residual_kernel|visual_build_kernel|visual_cost_kernel|visual_backsub_kernel' \
  --launch-skip 12 -c 8 -o gpurun_out/n_full_build python tools/dev_e2e.py C > gpurun_out/n_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/n_pytest_nd.txt; tail -3 gpurun_out/n_pytest_build.txt; tail -12 gpurun_out/n_pytest_offline.txt
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/n_bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["device_ms_per_step"], d["e2e"]["ms_call_A"], d["e2e"]["ms_call_B"], d.get("parity_C", {}).get("ok"))
except Exception as e:
    print("bench unreadable", e)
PY
