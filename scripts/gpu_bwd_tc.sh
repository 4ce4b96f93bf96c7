#!/bin/bash
# tcgen05 BPTT kernel: differential tests against the per-step kernel, then step trace + timing of modes 2 and 1
mkdir -p gpurun_out
TAG=${1:-r02n}
( timeout 900 python -m pytest tests/test_backward_gpu.py -m gpu -x -q -k "persistent_kernels" 2>&1 | tail -n 25 ) > gpurun_out/${TAG}_tests.log 2>&1
cat gpurun_out/${TAG}_tests.log | tail -n 12
for m in 2 1; do
  ( MODE=$m timeout 300 python scripts/bwd_trace.py ) > gpurun_out/${TAG}_bwd_trace_mode${m}.txt 2>&1
  tail -n 14 gpurun_out/${TAG}_bwd_trace_mode${m}.txt
done
