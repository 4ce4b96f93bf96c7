#!/bin/bash
mkdir -p gpurun_out
T=${1:-r02aq}
( timeout 600 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -n 3 ) > gpurun_out/${T}_tests.log 2>&1; tail -n 2 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-configs --no-gpu-reference > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -n 2 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('train', d['train']['ms_per_step'], 'disk', d['e2e_disk'])
PY
