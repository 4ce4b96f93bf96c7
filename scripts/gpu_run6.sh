#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 300 python scripts/rec_trace.py > gpurun_out/rec_trace.log 2>&1; echo "trace exit=$?"
tail -n 13 gpurun_out/rec_trace.log
timeout -k 10 600 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
tail -n 4 gpurun_out/pytest_gpu.log
timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"
cat gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['us_per_step'], d['roofline']['share_of_step'])"
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"stft_feat|loss_dc_partial|gemm_tc05" -s 12 -c 6 -f -o gpurun_out/prof_misc python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_misc.log 2>&1; echo "ncu misc exit=$?"
