#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I onssen_b200/csrc scripts/microbench/mma_cost.cu -o gpurun_out/mma_cost 2>/dev/null
timeout 120 gpurun_out/mma_cost > gpurun_out/mma_cost.txt; cat gpurun_out/mma_cost.txt
bash scripts/gpu_run3.sh
