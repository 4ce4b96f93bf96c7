#!/bin/bash
# Final round-2 evidence on one B200 (after the last kernel change): full -m gpu suite with parity numbers, smoke, default
# bench (both arms), ncu launch lists (forward bench; training step without the cluster BPTT kernel, which ncu cannot
# profile) and one ncu --set full capture of the inference recurrence for roofline.traffic.
mkdir -p gpurun_out
T=${1:-r02f}
( timeout 2400 python -m pytest tests -m gpu -q -s 2>&1 | grep -i "worst\|rel dev\|max abs err\|passed\|failed\|Error" | tail -40 ) > gpurun_out/${T}_tests.log 2>&1
tail -n 3 gpurun_out/${T}_tests.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke exit=$?"; tail -n 1 gpurun_out/${T}_smoke.log
timeout -k 10 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench exit=$?"; tail -n 2 gpurun_out/${T}_bench.err
timeout -k 10 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref exit=$?"; cut -c1-160 gpurun_out/${T}_bench_ref.json
timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-train --no-configs --no-gpu-reference > /dev/null 2>&1; echo "ncu fwd launches exit=$?"
STEPS=2 timeout -k 10 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv -k 'regex:^(?!.*lstm_bwd_tc)' --log-file gpurun_out/${T}_launches_train_steps2.csv python scripts/train_step.py > gpurun_out/${T}_ncu_train.log 2>&1; echo "ncu train launches exit=$?"
timeout -k 10 400 ncu --set full --clock-control none --import-source on -k regex:blstm_rec_kernel -s 9 -c 3 -f -o gpurun_out/${T}_prof_rec python bench.py --steps 2 --warmup 3 --no-train --no-configs --no-gpu-reference > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rec exit=$?"
ncu -i gpurun_out/${T}_prof_rec.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_rec_infer.csv 2>/dev/null
timeout -k 10 200 python scripts/rec_trace.py > gpurun_out/${T}_rec_step_trace.txt 2>&1; tail -n 1 gpurun_out/${T}_rec_step_trace.txt
python - <<PY
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'train', d.get('train',{}).get('ms_per_step'))
print('roofline', d['roofline']['frac'], d['roofline']['us_per_step'], 'launches', d['gpu_launches'])
print('disk', d.get('e2e_disk',{}).get('ms_per_step'), d.get('e2e_disk',{}).get('fraction_of_device_resident_train'))
print('gpu_ref', {k:v for k,v in d.get('gpu_reference',{}).items() if k in ('fwd_loss_ms','train_ms','speedup_fwd_loss','speedup_train')})
print('configs', {k:(round(v['fwd_loss_ms'],2), round(v['train_ms'],2)) for k,v in d.get('configs',{}).items()})
PY
