#!/bin/bash
# full GPU suite with the parity numbers printed (-s), then the default bench
tag=${1:-run}
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | grep -i "worst\|rel dev\|max abs err\|passed\|failed\|Error" | tail -40 ) > gpurun_out/${tag}_tests.log 2>&1
tail -n 30 gpurun_out/${tag}_tests.log
( timeout 1200 python bench.py --steps 20 --warmup 5 ${BENCH_FLAGS} ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d.get('train',{}).get('ms_per_step'))
print('roofline', d['roofline']['frac'], d['roofline']['us_per_step'])
print('disk', d.get('e2e_disk',{}).get('ms_per_step'), d.get('e2e_disk',{}).get('fraction_of_device_resident_train'))
print('configs', {k:(round(v['fwd_loss_ms'],2), round(v['train_ms'],2)) for k,v in d.get('configs',{}).items()})
PY
