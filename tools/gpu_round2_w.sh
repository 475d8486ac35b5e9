#!/bin/bash
# run W (7.6 GPU-minutes left): the voxel-map build with the slot-parallel fill pass (SlotFillF) — timing, launch list, GPU tests of
# everything that goes through the voxel map
mkdir -p gpurun_out
timeout 150 python tools/bench_voxel_map.py --cpu-sample-scans 0 > gpurun_out/w_voxel_map.json 2> gpurun_out/w_voxel_map.err; echo "bench_voxel_map rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/w_voxel_map.json").read().strip().splitlines()[-1])
    print("voxel map:", d["n_points"], "points", d["n_voxels"], "voxels", d["nnz"], "slots; ms_device", d["ms_device"], "ms_call", d["ms_call"], "launches", d["kernel_launches"], "checks", all(d["checks"].values()))
    print("depth:", json.dumps(d.get("depth"))[:600])
except Exception as e:
    print("unreadable", e)
PY
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/w_launches_setup.csv python tools/bench_voxel_map.py --scans 100 --points 50000 --repeats 1 --cpu-sample-scans 0 > gpurun_out/w_ncu_setup.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/w_launches_setup.csv 0 2>/dev/null | grep -n "VoxelFillF\|SlotFillF\|VoxelCountF" | head -8
timeout 200 python -m pytest tests/test_zz_voxel_gpu.py tests/test_window_batch_gpu.py tests/test_zz_golden_gpu.py -m gpu -x -q > gpurun_out/w_pytest.txt 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/w_pytest.txt)"; grep -B2 -A12 '^E  ' gpurun_out/w_pytest.txt | head -40
