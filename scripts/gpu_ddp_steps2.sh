#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02ab}
run() { timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29517 scripts/ddp_train_steps.py > gpurun_out/${TAG}_last.log 2>&1; grep "^world" gpurun_out/${TAG}_last.log | tee -a gpurun_out/${TAG}_ddp_steps.txt; grep -q "^world" gpurun_out/${TAG}_last.log || tail -n 8 gpurun_out/${TAG}_last.log; }
echo "# one rank" | tee -a gpurun_out/${TAG}_ddp_steps.txt
run 1
echo "# $N ranks, no gradient all-reduce (SKIP_SYNC=1)" | tee -a gpurun_out/${TAG}_ddp_steps.txt
SKIP_SYNC=1 run $N
echo "# $N ranks, deferred all-reduce (default)" | tee -a gpurun_out/${TAG}_ddp_steps.txt
run $N
echo "# $N ranks, overlapped all-reduce" | tee -a gpurun_out/${TAG}_ddp_steps.txt
ONSSEN_DDP_OVERLAP=1 run $N
