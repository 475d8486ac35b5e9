#!/bin/bash
# Run ZZ: the staged device checks against the reference's pipeline source (tests/test_zzz_ref_gpu.py without the test run Z covered).
mkdir -p gpurun_out
timeout 38 python -m pytest tests/test_zzz_ref_gpu.py -q -m gpu -k "pipeline or fusion or offline" > gpurun_out/zz_pytest.txt 2>&1
echo "pytest rc=$?"
tail -25 gpurun_out/zz_pytest.txt
