#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python scripts/rec_trace.py 2>&1 | tail -3
timeout -k 10 600 python -m pytest tests/test_parity_gpu.py tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -s -k "baseline_shape or blstm or fixture" 2>&1 | grep -i "parity\|passed\|failed\|Error" | head
