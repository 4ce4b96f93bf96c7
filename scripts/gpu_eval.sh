#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02ar}
( timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -n 25 ) > gpurun_out/${T}_tests.log 2>&1; tail -n 25 gpurun_out/${T}_tests.log
