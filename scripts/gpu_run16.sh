#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
ONSSEN_LIB=$PWD/onssen_b200/libonssen_b200_prev.so timeout -k 10 300 python scripts/bwd_trace.py 2>&1 | grep persistent= | sed 's/^/prev: /'
timeout -k 10 300 python scripts/bwd_trace.py > gpurun_out/bwd_trace.log 2>&1; grep persistent= gpurun_out/bwd_trace.log | sed 's/^/new:  /'
done
head -16 gpurun_out/bwd_trace.log
B=64 timeout -k 10 300 python scripts/bwd_trace.py 2>&1 | grep persistent=
timeout -k 10 900 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -i "worst\|passed\|failed\|Error" | tail -12
