#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/r02k_tests.log 2>&1
tail -12 gpurun_out/r02k_tests.log
( timeout 900 python bench.py --steps 20 --warmup 5 --no-configs ) > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
tail -3 gpurun_out/r02k_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02k_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d['train']['ms_per_step'])
print('disk', d.get('e2e_disk'))
print('gpu_ref', {k:v for k,v in d['gpu_reference'].items() if k!='what'})
PY
