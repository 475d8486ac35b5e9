#!/bin/bash
# run U: Hessian-build sweep (BASELINE configs[3]) and the profiles of the set-up rows B3 / B4 (launch list + --set full of their top kernels)
mkdir -p gpurun_out
timeout 300 python tools/build_sweep.py --out gpurun_out/u_build_sweep.md > gpurun_out/u_build_sweep.log 2>&1; echo "build sweep rc=$?"; tail -9 gpurun_out/u_build_sweep.md
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/u_launches_setup.csv python tools/bench_voxel_map.py --scans 200 --points 50000 --repeats 1 --cpu-sample-scans 0 > gpurun_out/u_ncu_setup.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/u_launches_setup.csv 2>/dev/null | sed -n '/total us/,$p' | head -30
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:vox_for_each_kernel|DeviceRadixSort|DeviceScan' --launch-skip 40 -c 40 -o gpurun_out/u_full_setup python tools/bench_voxel_map.py --scans 200 --points 50000 --repeats 1 --cpu-sample-scans 0 > gpurun_out/u_ncu_full.log 2>&1; echo "ncu full rc=$?"
