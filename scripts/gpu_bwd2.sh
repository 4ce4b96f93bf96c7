#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_bwd.log 2>&1; echo "bwd tests exit=$?"
tail -n 4 gpurun_out/pytest_bwd.log
timeout -k 10 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?"; tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d.get('train'))"
