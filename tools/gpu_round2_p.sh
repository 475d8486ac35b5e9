#!/bin/bash
# run P (1 GPU): the whole GPU suite as the driver runs it, smoke(), the default bench, the reference arm, and the ncu captures that are still missing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/p_pytest_gpu.txt 2>&1; echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/p_pytest_gpu.txt)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/p_smoke.txt 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/p_smoke.txt)"
timeout 900 python bench.py > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p_bench_ref.json 2> gpurun_out/p_bench_ref.err; echo "bench reference rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:lidar_build_kernel|lidar_residual_kernel' --launch-skip 2 -c 2 -o gpurun_out/p_full_lidar python tools/dev_e2e.py C > gpurun_out/p_ncu_full.log 2>&1; echo "ncu lidar rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/p_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-voxel-map > gpurun_out/p_ncu_bench.log 2>&1; echo "launch list rc=$?"
for dm in 0 1 2 4 7; do echo "dense mode $dm: $(LVBA_ND_PIPELINE=0 LVBA_DENSE_MODE=$dm timeout 100 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-60)"; done
LVBA_SETUP_TIMING=1 timeout 300 python tools/bench_voxel_map.py > gpurun_out/p_voxel_laps.json 2> gpurun_out/p_voxel_laps.txt; grep -h "voxel lookup\]" gpurun_out/p_voxel_laps.txt | tail -6
python - <<'PY'
import json
for f in ("gpurun_out/p_bench.json", "gpurun_out/p_bench_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("device_ms_per_step"), d.get("roofline_residual", {}).get("frac"))
    except Exception as e:
        print(f, "unreadable", e)
PY
