#!/bin/bash
mkdir -p gpurun_out
P=$PWD/onssen_b200/libonssen_b200_prev.so
for rep in 1 2; do
ONSSEN_LIB=$P timeout -k 10 200 python scripts/bwd_trace.py 2>&1 | grep persistent= | sed 's/^/bwd prev: /'
timeout -k 10 200 python scripts/bwd_trace.py > gpurun_out/bwd_trace.log 2>&1; grep persistent= gpurun_out/bwd_trace.log | sed 's/^/bwd new:  /'
done
head -9 gpurun_out/bwd_trace.log
B=64 timeout -k 10 200 python scripts/bwd_trace.py 2>&1 | grep persistent=
timeout -k 10 900 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -2
