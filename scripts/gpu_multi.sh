#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit=$?"
tail -5 gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['e2e'], d.get('train'), d['clocks'])"
timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N exit=$?"
cat gpurun_out/bench_ref_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('impl','value','n_gpus')}, d['cpu_baseline']['cores'])"
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 8 --warmup 3 --sync-bn > gpurun_out/bench_syncbn_n$N.json 2> gpurun_out/bench_syncbn_n$N.err; echo "bench sync-bn N=$N exit=$?"
tail -3 gpurun_out/bench_syncbn_n$N.err
cat gpurun_out/bench_syncbn_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('train'))"
