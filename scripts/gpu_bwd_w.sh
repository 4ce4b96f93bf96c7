#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02ao}
( timeout 900 python -m pytest tests/test_backward_gpu.py -m gpu -x -q -k "persistent_kernels" 2>&1 | tail -n 5 ) > gpurun_out/${TAG}_tests.log 2>&1
tail -n 3 gpurun_out/${TAG}_tests.log
bash scripts/gpu_ab_bwd.sh ${TAG} main w8 main w8
grep -A11 "step 102" gpurun_out/${TAG}_bwd_main_p0.txt
