#!/bin/bash
mkdir -p gpurun_out
for b in 32 64; do
  ( B=$b timeout 120 python scripts/rec_trace.py ) > gpurun_out/r02h_trace_b${b}.txt 2>&1
  echo "B=$b: $(tail -n 1 gpurun_out/r02h_trace_b${b}.txt)"
done
( timeout 600 python -m pytest tests/test_reference_gpu.py tests/test_kernels_gpu.py -m gpu -q -s -k "cfg4 or batch_limits" 2>&1 | grep "cfg[24] \|passed\|failed\|Error" ) > gpurun_out/r02h_cfg4.log 2>&1
cat gpurun_out/r02h_cfg4.log
