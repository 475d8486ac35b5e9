#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/trace_E.py E > gpurun_out/h_trace_E.txt 2>&1; echo "trace E rc=$?"
timeout 900 python tools/trace_E.py B > gpurun_out/h_trace_B.txt 2>&1; echo "trace B rc=$?"
cat gpurun_out/h_trace_E.txt; cat gpurun_out/h_trace_B.txt
