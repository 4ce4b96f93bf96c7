#!/bin/bash
# full GPU suite + default bench: scripts/gpu_full.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/${tag}_tests.log 2>&1
tail -6 gpurun_out/${tag}_tests.log
( timeout 1200 python bench.py --steps 20 --warmup 5 ${BENCH_FLAGS} ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d.get('train',{}).get('ms_per_step'))
print('roofline', d['roofline']['frac'], d['roofline']['us_per_step'])
print('disk', d.get('e2e_disk'))
print('gpu_ref', {k:v for k,v in d.get('gpu_reference',{}).items() if k!='what'})
print('configs', d.get('configs'))
PY
