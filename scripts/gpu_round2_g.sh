#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r02g_tests.log 2>&1
tail -3 gpurun_out/r02g_tests.log
for b in 32 64; do
  ( B=$b timeout 120 python scripts/rec_trace.py ) > gpurun_out/r02g_trace_b${b}.txt 2>&1
  echo "B=$b: $(tail -n 1 gpurun_out/r02g_trace_b${b}.txt)"
done
( timeout 600 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02g_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d['train']['ms_per_step'], 'frac', d['roofline']['frac'])
print('gpu_ref', d['gpu_reference'].get('fwd_loss_ms'), d['gpu_reference'].get('train_ms'))
print({k:(v['fwd_loss_ms'], v['train_ms']) for k,v in d['configs'].items()})
PY
