#!/bin/bash
# run T: final state — the GPU suite as the driver runs it, smoke(), the default bench, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/t_pytest_gpu.txt 2>&1; echo "pytest -m gpu rc=$? $(tail -1 gpurun_out/t_pytest_gpu.txt)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/t_smoke.txt 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/t_smoke.txt)"
timeout 900 python bench.py > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t_bench_ref.json 2> gpurun_out/t_bench_ref.err; echo "bench reference rc=$?"
timeout 200 python tools/solver_bench.py 2000x30 5000x30 > gpurun_out/t_solver_bench.txt 2>&1
grep -B2 -A14 "^FAILED\|^E  " gpurun_out/t_pytest_gpu.txt | head -40
cut -c1-400 gpurun_out/t_solver_bench.txt
python - <<'PY'
import json
for f in ("gpurun_out/t_bench.json", "gpurun_out/t_bench_ref.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), d.get("e2e", {}).get("value"), d.get("device_ms_per_step"), d.get("e2e", {}).get("ms_call_A"), d.get("e2e", {}).get("ms_call_B"), d.get("parity_C", {}).get("ok"))
    except Exception as e:
        print(f, "unreadable", e)
PY
