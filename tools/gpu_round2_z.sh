#!/bin/bash
# Run Z: the CUDA path against the fixture the reference's own source wrote (tests/test_zzz_ref_gpu.py) — one short call.
mkdir -p gpurun_out
timeout 70 python -m pytest tests/test_zzz_ref_gpu.py -x -q -m gpu > gpurun_out/z_pytest.txt 2>&1
echo "pytest rc=$?"
tail -15 gpurun_out/z_pytest.txt
