#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02ae}
( timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_reference_gpu.py -m gpu -x -q 2>&1 | tail -n 6 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -n 4 gpurun_out/${TAG}_tests.log
BATCHES="32 48 64 96" bash scripts/gpu_ab_rec.sh ${TAG} main | grep "B="
