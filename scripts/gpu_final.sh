#!/bin/bash
# Consolidated evidence run: tests, smoke, bench (ours + reference arm), launch list, full ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
grep -i "parity\|passed\|failed" gpurun_out/pytest_gpu.log | tail -6
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout -k 10 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -2 gpurun_out/bench.err
cat gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['us_per_step'], d['roofline']['frac'], d['clocks'], d.get('train'), d['cpu_baseline'])"
timeout -k 10 300 python scripts/cfg3_time.py 2>&1 | tail -1 | tee gpurun_out/cfg3.txt
timeout -k 10 300 python scripts/rec_trace.py > gpurun_out/rec_trace.log 2>&1; tail -1 gpurun_out/rec_trace.log
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-train > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit=$?"
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"blstm_rec|stft_feat|loss_dc_partial|gemm_tc05" -s 40 -c 10 -f -o gpurun_out/prof_all python bench.py --steps 2 --warmup 3 --no-train > gpurun_out/ncu_all.log 2>&1; echo "ncu full exit=$?"
