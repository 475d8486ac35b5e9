#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nd_solver_gpu.py tests/test_config_c_gpu.py -x -q -s > gpurun_out/e_pytest.txt 2>&1; echo "pytest nd + config C rc=$?"
LVBA_ND_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/e_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/e_ncu1.log 2>&1
timeout 600 python tools/solver_bench.py 2000x30 2000x20 5000x30 > gpurun_out/e_solver_bench.txt 2> gpurun_out/e_solver_bench.err; echo "solver bench rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err; echo "bench ref rc=$?"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/e_pytest_all.txt 2>&1; echo "pytest all gpu rc=$?"
grep -E "config C|passed|failed" gpurun_out/e_pytest.txt | tail; cut -c1-400 gpurun_out/e_solver_bench.txt; tail -3 gpurun_out/e_pytest_all.txt
