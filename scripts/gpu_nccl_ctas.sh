#!/bin/bash
# training-step time at N GPUs for several NCCL CTA limits (the cooperative recurrence kernels need 114-120 free SMs)
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02x}
for c in ${CTAS:-default 4 8 16}; do
  if [ "$c" == default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$c; fi
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --no-configs --no-gpu-reference --no-disk > gpurun_out/${TAG}_n${N}_ctas${c}.json 2> gpurun_out/${TAG}_n${N}_ctas${c}.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_n${N}_ctas${c}.json').read().strip().splitlines()[-1])
print('NCCL_MAX_CTAS=$c', 'fwd ms', round(d['ms_per_step'],3), 'train ms', round(d['train']['ms_per_step'],3), {k:v for k,v in d.items() if 'ddp' in k or 'check' in k})
PY
done
