#!/bin/bash
# Validate enhance training + vectorised normalize_bwd + persistent BPTT; bench; training launch list.
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
grep -i "worst\|passed\|failed\|Error" gpurun_out/pytest_gpu.log | tail -12
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -1 gpurun_out/smoke.log
timeout -k 10 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -2 gpurun_out/bench.err
cat gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['us_per_step'], d['roofline']['frac'], d['clocks'], d.get('train'), d['cpu_baseline'])"
STEPS=2 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/train_launches.csv python scripts/train_step.py > gpurun_out/ncu_train.log 2>&1; echo "ncu train exit=$?"
