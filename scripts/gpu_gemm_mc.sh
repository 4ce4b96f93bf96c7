#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02ag}
( timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" 2>&1 | tail -n 8 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -n 4 gpurun_out/${TAG}_tests.log
