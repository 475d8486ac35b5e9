#!/bin/bash
# run Q: where a pivot of the dense separator factorisation spends its cycles (development build with clocks), the no-overlap
# variant, the tests run P did not reach, bench
mkdir -p gpurun_out
echo "clocks: $(LVBA_B200_DEV_LIB=liblvba_b200_clk.so LVBA_ND_PIPELINE=0 LVBA_ND_GRAPH=0 LVBA_DENSE_MODE=16 timeout 100 python tools/solve_once.py 2000 30 3 16 1 2>&1 | grep 'dense clocks' | head -4)"
echo "clocks, pair threads idle: $(LVBA_B200_DEV_LIB=liblvba_b200_clk.so LVBA_ND_PIPELINE=0 LVBA_ND_GRAPH=0 LVBA_DENSE_MODE=17 timeout 100 python tools/solve_once.py 2000 30 3 16 1 2>&1 | grep 'dense clocks' | head -2)"
for dm in 0 8; do echo "dense mode $dm: $(LVBA_ND_PIPELINE=0 LVBA_DENSE_MODE=$dm timeout 100 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-110)"; done
echo "default: $(timeout 100 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-110)"
timeout 1200 python -m pytest tests/test_window_batch_gpu.py tests/test_nd_solver_gpu.py tests/test_zz_depth_gpu.py tests/test_zz_fuse_gpu.py tests/test_zz_offline_gpu.py tests/test_zz_voxel_gpu.py tests/test_zz_wide_gpu.py -q > gpurun_out/q_pytest.txt 2>&1; echo "pytest rest rc=$? $(tail -1 gpurun_out/q_pytest.txt)"
timeout 900 python bench.py --no-voxel-map > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
LVBA_ND_PIPELINE=0 LVBA_ND_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/q_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-voxel-map --e2e-steps 1 > gpurun_out/q_ncu_bench.log 2>&1; echo "launch list rc=$?"
grep -B2 -A12 "FAILED\|Error" gpurun_out/q_pytest.txt | head -60
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/q_bench.json").read().strip().splitlines()[-1])
    print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["device_ms_per_step"], d["e2e"]["ms_call_A"], d["e2e"]["ms_call_B"], d.get("parity_C", {}).get("ok"))
except Exception as e:
    print("bench unreadable", e)
PY
