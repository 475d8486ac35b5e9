#!/bin/bash
# run X: persistent helper threads for the host set-up loops — one-shot call times with laps, GPU tests of both LM paths
mkdir -p gpurun_out
timeout 200 python tools/e2e_laps.py C 6 > gpurun_out/x_e2e.txt 2>&1; echo "e2e rc=$?"; grep -v "^$" gpurun_out/x_e2e.txt | tail -28
timeout 150 python -m pytest tests/test_lidar_gpu.py tests/test_visual_gpu.py -m gpu -x -q > gpurun_out/x_pytest.txt 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/x_pytest.txt)"; grep -B2 -A12 '^E  ' gpurun_out/x_pytest.txt | head -40
