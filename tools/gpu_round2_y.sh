#!/bin/bash
# run Y (3.8 GPU-minutes left): the default bench line of the final build (pinned scratch pooled, persistent set-up helpers, slot-parallel
# voxel-map fill), then the batched window BA tests (their per-window read-back block comes from the pinned pool)
mkdir -p gpurun_out
timeout 170 python bench.py > gpurun_out/y_bench.json 2> gpurun_out/y_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/y_bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["device_ms_per_step"], d["e2e"]["ms_call_A"], d["e2e"]["ms_call_B"], "parity", d.get("parity_C", {}).get("ok"), "cpu", d.get("cpu_baseline", {}).get("value"))
    print("clocks", d.get("clocks"), "voxel_map ms_device", d.get("voxel_map", {}).get("ms_device"))
except Exception as e:
    print("bench unreadable", e); print(open("gpurun_out/y_bench.err").read()[-1500:])
PY
timeout 60 python -m pytest tests/test_window_batch_gpu.py -m gpu -x -q > gpurun_out/y_pytest.txt 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/y_pytest.txt)"
