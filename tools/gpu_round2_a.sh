#!/bin/bash
# first hardware contact of the substructured solver: parity tests, solver timings, a short bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_nd_solver_gpu.py -x -q > gpurun_out/a_pytest_nd.txt 2>&1; echo "nd pytest rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 600 python tools/solver_bench.py > gpurun_out/a_solver_bench.txt 2> gpurun_out/a_solver_bench.err; echo "solver bench rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 900 python -m pytest tests/test_lidar_gpu.py tests/test_visual_gpu.py -x -q > gpurun_out/a_pytest_lv.txt 2>&1; echo "lidar/visual pytest rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-voxel-map > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" | tee -a gpurun_out/a_summary.txt
tail -5 gpurun_out/a_pytest_nd.txt; cat gpurun_out/a_solver_bench.txt; tail -3 gpurun_out/a_pytest_lv.txt; cat gpurun_out/a_bench.json | head -c 1500
