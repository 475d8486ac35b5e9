#!/bin/bash
# run R: warp-cooperative 6x6 inverse in every factorisation kernel — clocks, solve times, the whole GPU suite, bench
mkdir -p gpurun_out
timeout 60 python tools/solve_once.py 600 30 3 4 1 > gpurun_out/r_first.txt 2>&1; echo "first solve rc=$? $(tail -1 gpurun_out/r_first.txt | cut -c1-150)"
echo "clocks: $(LVBA_B200_DEV_LIB=liblvba_b200_clk.so LVBA_ND_PIPELINE=0 LVBA_ND_GRAPH=0 LVBA_DENSE_MODE=16 timeout 100 python tools/solve_once.py 2000 30 3 16 1 2>&1 | grep 'dense clocks' | head -2)"
echo "pipeline off: $(LVBA_ND_PIPELINE=0 timeout 100 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-110)"
echo "default 2000x30: $(timeout 100 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-110) | p32 $(timeout 100 python tools/solve_once.py 2000 30 3 32 5 2>&1 | tail -1 | cut -c1-40) | 2000x20 $(timeout 100 python tools/solve_once.py 2000 20 3 16 5 2>&1 | tail -1 | cut -c1-40) | 5000x30 $(timeout 100 python tools/solve_once.py 5000 30 3 32 5 2>&1 | tail -1 | cut -c1-40) | twisted $(timeout 100 python tools/solve_once.py 2000 30 2 0 5 2>&1 | tail -1 | cut -c1-40)"
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r_pytest_gpu.txt 2>&1; echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/r_pytest_gpu.txt)"
timeout 900 python bench.py > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err; echo "bench rc=$?"
grep -B2 -A14 "^FAILED\|^E  " gpurun_out/r_pytest_gpu.txt | head -60
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r_bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["device_ms_per_step"], d["e2e"]["ms_call_A"], d["e2e"]["ms_call_B"], d.get("parity_C", {}).get("ok"))
    print("voxel_map", json.dumps(d.get("voxel_map", {}))[:900])
except Exception as e:
    print("bench unreadable", e)
PY
