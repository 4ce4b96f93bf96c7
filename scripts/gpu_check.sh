#!/bin/bash
# First-line GPU validation: each kernel group in its own process under a hard timeout (a hung kernel must
# not take the box with it). Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, pytest args...
  local name=$1; local to=$2; shift 2
  timeout -k 10 $to python -m pytest -q -m gpu -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1
  echo "$name exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$name.log
}
rm -f gpurun_out/summary.txt
run gemm 300 tests/test_kernels_gpu.py -k "gemm"
run misc 300 tests/test_kernels_gpu.py -k "pack or bn or loss or stft"
run blstm_simt 300 tests/test_kernels_gpu.py -k "blstm and False"
run blstm_tc 300 tests/test_kernels_gpu.py -k "blstm and True"
run parity 600 tests/test_parity_gpu.py -s
cat gpurun_out/summary.txt
