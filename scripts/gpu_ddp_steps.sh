#!/bin/bash
# per-step training time under torchrun: deferred (default) vs overlapped gradient all-reduce, both BPTT kernels
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r02z}
run() { timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/ddp_train_steps.py > gpurun_out/${TAG}_last.log 2>&1; grep "^world" gpurun_out/${TAG}_last.log | tee -a gpurun_out/${TAG}_ddp_steps.txt; grep -q "^world" gpurun_out/${TAG}_last.log || tail -n 8 gpurun_out/${TAG}_last.log; }
echo "# deferred all-reduce (default)" | tee -a gpurun_out/${TAG}_ddp_steps.txt
ONSSEN_BPTT_MODE=2 run
ONSSEN_BPTT_MODE=2 run
ONSSEN_BPTT_MODE=1 run
echo "# overlapped all-reduce (ONSSEN_DDP_OVERLAP=1, 48 SMs reserved)" | tee -a gpurun_out/${TAG}_ddp_steps.txt
ONSSEN_DDP_OVERLAP=1 ONSSEN_BPTT_MODE=2 run
ONSSEN_DDP_OVERLAP=1 ONSSEN_BPTT_MODE=1 run
