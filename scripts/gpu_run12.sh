#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python scripts/bwd_trace.py > gpurun_out/bwd_trace.log 2>&1; echo "trace exit=$?"; cat gpurun_out/bwd_trace.log | tail -60
B=64 timeout -k 10 300 python scripts/bwd_trace.py 2>&1 | grep persistent=
timeout -k 10 600 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -s 2>&1 | grep -i "worst\|passed\|failed\|Error" | tail -12
