#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02b_tests.log 2>&1
( timeout 120 python scripts/rec_trace.py ) > gpurun_out/r02b_trace_new.txt 2>&1
( ONSSEN_LIB=onssen_b200/libonssen_b200_r01rec.so timeout 120 python scripts/rec_trace.py ) > gpurun_out/r02b_trace_old.txt 2>&1
( B=64 timeout 120 python scripts/rec_trace.py | tail -1 ) > gpurun_out/r02b_trace_new_b64.txt 2>&1
( B=64 ONSSEN_LIB=onssen_b200/libonssen_b200_r01rec.so timeout 120 python scripts/rec_trace.py | tail -1 ) > gpurun_out/r02b_trace_old_b64.txt 2>&1
( timeout 600 python -m pytest tests/test_reference_gpu.py -m gpu -q -s -k cfg4 2>&1 | grep "cfg4" ) > gpurun_out/r02b_cfg4.log 2>&1
tail -3 gpurun_out/r02b_tests.log; tail -1 gpurun_out/r02b_trace_new.txt gpurun_out/r02b_trace_old.txt
