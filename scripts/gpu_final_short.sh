#!/bin/bash
# last check of a build: full -m gpu suite (parity numbers printed), smoke, default bench (both arms)
mkdir -p gpurun_out
T=${1:-r02g}
( timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | grep -i "worst\|rel dev\|max abs err\|passed\|failed\|Error" | tail -40 ) > gpurun_out/${T}_tests.log 2>&1
tail -n 2 gpurun_out/${T}_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit=$?"; tail -n 1 gpurun_out/${T}_smoke.log
timeout -k 10 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit=$?"; tail -n 2 gpurun_out/${T}_bench.err
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref exit=$?"; cut -c1-160 gpurun_out/${T}_bench_ref.json
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d.get('train',{}).get('ms_per_step'))
print('roofline', d['roofline']['frac'], d['roofline']['us_per_step'], 'launches', d['gpu_launches'])
print('disk', d.get('e2e_disk',{}).get('ms_per_step'), d.get('e2e_disk',{}).get('fraction_of_device_resident_train'))
print('gpu_ref', {k:v for k,v in d.get('gpu_reference',{}).items() if k in ('fwd_loss_ms','train_ms','speedup_fwd_loss','speedup_train')})
print('configs', {k:(round(v['fwd_loss_ms'],2), round(v['train_ms'],2)) for k,v in d.get('configs',{}).items()})
PY
