#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_backward_gpu.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02f_tests.log 2>&1
tail -3 gpurun_out/r02f_tests.log
for b in 32 64 96; do for ch in 1 2; do
  ( B=$b CHAINS=$ch timeout 120 python scripts/rec_trace.py ) > gpurun_out/r02f_trace_b${b}_ch$ch.txt 2>&1
  echo "B=$b chains=$ch: $(tail -n 1 gpurun_out/r02f_trace_b${b}_ch$ch.txt)"
done; done
