#!/bin/bash
# run O (8 GPUs): config E at N = 8 and N = 4, the driver's launch line
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/o_bench_E_n$n.json 2> gpurun_out/o_bench_E_n$n.err
  echo "N=$n rc=$? $(tail -1 gpurun_out/o_bench_E_n$n.json | cut -c1-200)"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/mgpu_check.py > gpurun_out/o_mgpu_check.txt 2>&1; echo "mgpu_check rc=$? $(tail -2 gpurun_out/o_mgpu_check.txt | tr '\n' ' ' | cut -c1-300)"
python - <<'PY'
import json
for n in (8, 4):
    try:
        d = json.loads(open(f"gpurun_out/o_bench_E_n{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d.get("device_ms_per_step"), d.get("run", {}).get("n1_same_config"), d.get("run", {}).get("efficiency_vs_n1_same_config"), d.get("run", {}).get("nccl_payload_bytes_per_step_rank0"), d.get("parity_vs_n1"))
    except Exception as e:
        print(n, "unreadable", e)
PY
