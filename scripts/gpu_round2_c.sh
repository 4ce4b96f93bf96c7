#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02c_tests.log 2>&1
for v in "" _rec00 _rec01 _rec10 _r01rec; do
  ( ONSSEN_LIB=onssen_b200/libonssen_b200$v.so timeout 120 python scripts/rec_trace.py ) > gpurun_out/r02c_trace$v.txt 2>&1
done
( timeout 120 python scripts/stft_time.py ) > gpurun_out/r02c_stft_new.txt 2>&1
( ONSSEN_LIB=onssen_b200/libonssen_b200_r01all.so timeout 120 python scripts/stft_time.py ) > gpurun_out/r02c_stft_old.txt 2>&1
tail -3 gpurun_out/r02c_tests.log; tail -qn 1 gpurun_out/r02c_trace*.txt; cat gpurun_out/r02c_stft_*.txt
