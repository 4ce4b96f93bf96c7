#!/bin/bash
# Round-1 evidence on one B200: tests, smoke, bench (both arms), secondary configs, BPTT trace, ncu launch lists
# (forward bench + training step) and full captures of the BPTT / optimiser kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
grep -i "worst\|passed\|failed\|Error" gpurun_out/pytest_gpu.log | tail -12
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout -k 10 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -2 gpurun_out/bench.err
cat gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['us_per_step'], d['roofline']['frac'], d['clocks'], d.get('train'), d['cpu_baseline'])"
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit=$?"; cat gpurun_out/bench_ref.json | cut -c1-200
timeout -k 10 900 python scripts/cfg_time.py > gpurun_out/cfg_times.txt 2>&1; echo "cfg exit=$?"; tail -5 gpurun_out/cfg_times.txt
timeout -k 10 300 python scripts/bwd_trace.py > gpurun_out/bwd_trace.log 2>&1; tail -1 gpurun_out/bwd_trace.log
timeout -k 10 300 python scripts/rec_trace.py > gpurun_out/rec_trace.log 2>&1; tail -1 gpurun_out/rec_trace.log
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-train > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit=$?"
STEPS=2 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/train_launches.csv python scripts/train_step.py > gpurun_out/ncu_train.log 2>&1; echo "ncu train exit=$?"
STEPS=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_bwd_persistent|adam_kernel|grad_sumsq|normalize_bwd|loss_dc_bwd" -c 7 -f -o gpurun_out/prof_train python scripts/train_step.py > gpurun_out/ncu_train_full.log 2>&1; echo "ncu full exit=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
