#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "blstm" > gpurun_out/pytest_blstm.log 2>&1; echo "blstm tests exit=$?"
tail -n 3 gpurun_out/pytest_blstm.log
timeout -k 10 300 python scripts/rec_trace.py > gpurun_out/rec_trace.log 2>&1; echo "trace exit=$?"
tail -n 12 gpurun_out/rec_trace.log
