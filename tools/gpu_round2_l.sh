#!/bin/bash
# run L: dense factor v3 (register budgets fixed), build kernels with the 4-lanes-per-pair mapping, visual offline tool
mkdir -p gpurun_out
timeout 60 python tools/solve_once.py 600 30 3 4 1 > gpurun_out/l_first.txt 2>&1; rc=$?
echo "first dense-v3 solve rc=$rc $(tail -1 gpurun_out/l_first.txt)"
if [ $rc -ne 0 ]; then export LVBA_ND_DENSE=0; echo "dense kernel disabled for the rest of this run"; fi
timeout 300 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/l_pytest_nd.txt 2>&1; echo "pytest nd rc=$?"
timeout 600 python -m pytest tests/test_lidar_gpu.py tests/test_visual_gpu.py tests/test_config_c_gpu.py -x -q > gpurun_out/l_pytest_build.txt 2>&1; echo "pytest build rc=$?"
LVBA_ND_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/l_ncu1.log 2>&1
timeout 300 python tools/solver_bench.py 2000x30 2000x20 5000x30 > gpurun_out/l_solver_bench.txt 2>&1
timeout 600 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"
LVBA_SORT_VOXELS=1 timeout 600 python bench.py > gpurun_out/l_bench_sorted.json 2> gpurun_out/l_bench_sorted.err; echo "bench sorted rc=$?"
timeout 600 python -m pytest tests/test_zz_fuse_gpu.py tests/test_zz_offline_gpu.py -x -q > gpurun_out/l_pytest_fuse.txt 2>&1; echo "pytest fuse/offline rc=$?"
tail -3 gpurun_out/l_pytest_nd.txt; tail -3 gpurun_out/l_pytest_build.txt; tail -8 gpurun_out/l_pytest_fuse.txt; cut -c1-300 gpurun_out/l_solver_bench.txt
python - <<'PY'
import json
for f in ("gpurun_out/l_bench.json", "gpurun_out/l_bench_sorted.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), {k: d["breakdown_ms"][k] for k in d.get("breakdown_ms", {})})
    except Exception as e:
        print(f, "unreadable", e)
PY
