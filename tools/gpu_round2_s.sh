#!/bin/bash
# run S: does the warp-cooperative inverse pay off when the pair threads are not competing for issue slots?
mkdir -p gpurun_out
for dm in 16 17 24; do echo "clocks mode $dm: $(LVBA_B200_DEV_LIB=liblvba_b200_clk.so LVBA_ND_PIPELINE=0 LVBA_ND_GRAPH=0 LVBA_DENSE_MODE=$dm timeout 100 python tools/solve_once.py 2000 30 3 16 1 2>&1 | grep 'dense clocks' | head -1)"; done
for dm in 0 8; do echo "mode $dm pipeline off: $(LVBA_ND_PIPELINE=0 LVBA_DENSE_MODE=$dm timeout 100 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-60) | pipeline on: $(LVBA_DENSE_MODE=$dm timeout 100 python tools/solve_once.py 2000 30 3 16 5 2>&1 | tail -1 | cut -c1-60) | 5000x30: $(LVBA_DENSE_MODE=$dm timeout 100 python tools/solve_once.py 5000 30 3 32 5 2>&1 | tail -1 | cut -c1-40) | 2000x20: $(LVBA_DENSE_MODE=$dm timeout 100 python tools/solve_once.py 2000 20 3 16 5 2>&1 | tail -1 | cut -c1-40)"; done
LVBA_ND_PIPELINE=0 LVBA_ND_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s_launches_2000_16.csv python tools/solve_once.py 2000 30 3 16 2 > gpurun_out/s_ncu1.log 2>&1
python tools/launch_summary.py gpurun_out/s_launches_2000_16.csv 2>/dev/null | sed -n '/total us/,$p' | head -12
