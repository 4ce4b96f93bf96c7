#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 1200 python -m pytest tests -q -m gpu -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?"
grep -i "worst\|passed\|failed\|Error" gpurun_out/pytest_gpu.log | tail -14
timeout -k 10 900 python scripts/cfg_time.py > gpurun_out/cfg_times.txt 2>&1; echo "cfg exit=$?"; tail -6 gpurun_out/cfg_times.txt
timeout -k 10 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -2 gpurun_out/bench.err
cat gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['us_per_step'], d['roofline']['frac'], d['clocks'], d.get('train'))"
