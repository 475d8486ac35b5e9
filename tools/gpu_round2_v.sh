#!/bin/bash
# run V (11.8 GPU-minutes left): Hessian-build sweep (BASELINE configs[3]) with the round-2 kernels, then the set-up rows B3 / B4:
# ncu launch list and --set full of their pass kernels.  Every piece under its own timeout, most valuable first.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_zz_golden_gpu.py -m gpu -x -q > gpurun_out/v_pytest_golden.txt 2>&1; echo "golden fixture on the device rc=$? $(tail -1 gpurun_out/v_pytest_golden.txt)"; grep -A12 '^E  ' gpurun_out/v_pytest_golden.txt | head -30
timeout 170 python tools/build_sweep.py --out gpurun_out/v_build_sweep.md > gpurun_out/v_build_sweep.log 2>&1; echo "build sweep rc=$?"; tail -9 gpurun_out/v_build_sweep.md
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/v_launches_setup.csv python tools/bench_voxel_map.py --scans 100 --points 50000 --repeats 1 --cpu-sample-scans 0 > gpurun_out/v_ncu_setup.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/v_launches_setup.csv 2>/dev/null | sed -n '/total us/,$p' | head -30
timeout 200 ncu --set full --clock-control none --import-source on -k 'regex:vox_for_each_kernel|DeviceRadixSort|DeviceScan' --launch-skip 40 -c 30 -o gpurun_out/v_full_setup python tools/bench_voxel_map.py --scans 100 --points 50000 --repeats 1 --cpu-sample-scans 0 > gpurun_out/v_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 200 python -m pytest tests/test_lidar_gpu.py tests/test_visual_gpu.py tests/test_zz_fuse_gpu.py tests/test_zz_depth_gpu.py -m gpu -x -q > gpurun_out/v_pytest_subset.txt 2>&1; echo "pytest subset (library rebuilt with the ABI exception guards) rc=$? $(tail -1 gpurun_out/v_pytest_subset.txt)"
ls -la gpurun_out | head -20
