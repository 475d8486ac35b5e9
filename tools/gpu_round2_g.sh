#!/bin/bash
# two GPUs: row-owned sharded solve against the oracle, then the bench at N = 2 (config E)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/mgpu_check.py B > gpurun_out/g_mgpu_B.txt 2>&1; echo "mgpu B rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 tools/mgpu_check.py A > gpurun_out/g_mgpu_A.txt 2>&1; echo "mgpu A rc=$?"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/g_bench2.json 2> gpurun_out/g_bench2.err; echo "bench N=2 rc=$?"
tail -12 gpurun_out/g_mgpu_B.txt; tail -6 gpurun_out/g_mgpu_A.txt; head -c 1500 gpurun_out/g_bench2.json; tail -5 gpurun_out/g_bench2.err
