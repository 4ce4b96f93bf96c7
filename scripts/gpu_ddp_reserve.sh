#!/bin/bash
mkdir -p gpurun_out
TAG=${1:-r02y}
for r in 0 48; do
  ( RESERVE=$r MODE=2 timeout 300 python scripts/bwd_trace.py ) > gpurun_out/${TAG}_bwd_reserve${r}.txt 2>&1
  echo "reserve $r: $(tail -n 1 gpurun_out/${TAG}_bwd_reserve${r}.txt)"
done
for r in 48 0 48; do
  export ONSSEN_DDP_SM_RESERVE=$r
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 --no-configs --no-gpu-reference --no-disk > gpurun_out/${TAG}_n2_reserve${r}.json 2> gpurun_out/${TAG}_n2_reserve${r}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_n2_reserve${r}.json').read().strip().splitlines()[-1])
print('reserve=$r', 'fwd ms', round(d['ms_per_step'],3), 'train ms', round(d['train']['ms_per_step'],3))
PY
done
