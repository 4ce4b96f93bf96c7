#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "blstm or pack" > gpurun_out/pytest_blstm.log 2>&1; echo "blstm tests exit=$?"
tail -n 12 gpurun_out/pytest_blstm.log
timeout -k 10 300 python scripts/rec_trace.py > gpurun_out/rec_trace.log 2>&1; echo "trace exit=$?"
cat gpurun_out/rec_trace.log | tail -22
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
tail -n 6 gpurun_out/pytest_gpu.log
timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
cat gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['us_per_step'], d['roofline']['share_of_step'])"
if [ "$1" == "ncugemm" ]; then
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 10 -c 3 -f -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit=$?"
fi
