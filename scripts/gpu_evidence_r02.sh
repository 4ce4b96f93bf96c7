#!/bin/bash
# Round-2 evidence on one B200: ncu launch lists (forward bench, training step), one `ncu --set full` capture of the
# dominant kernels (raw pages exported as csv on the box), step traces of both persistent recurrence kernels.
mkdir -p gpurun_out
T=${1:-r02}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,power.limit --format=csv > gpurun_out/${T}_gpu.txt 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-train --no-configs --no-gpu-reference > gpurun_out/${T}_ncu_launch.log 2>&1; echo "ncu fwd launches exit=$?"
STEPS=2 timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${T}_launches_train_steps2.csv python scripts/train_step.py > gpurun_out/${T}_ncu_train.log 2>&1; echo "ncu train launches exit=$?"
STEPS=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"blstm_rec_kernel|lstm_bwd_tc_kernel|stft_feat|loss_dc_partial|loss_dc_bwd" -c 9 -f -o gpurun_out/${T}_prof_rec python scripts/train_step.py > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full (recurrences) exit=$?"
STEPS=1 timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc05_kernel" -c 12 -f -o gpurun_out/${T}_prof_gemm python scripts/train_step.py > gpurun_out/${T}_ncu_full_gemm.log 2>&1; echo "ncu full (gemm) exit=$?"
for f in rec gemm; do
  ncu -i gpurun_out/${T}_prof_${f}.ncu-rep --page raw --csv > gpurun_out/${T}_ncu_full_${f}.csv 2>/dev/null
done
ls -la gpurun_out/${T}_*.ncu-rep gpurun_out/${T}_ncu_full_*.csv 2>/dev/null
timeout -k 10 300 python scripts/rec_trace.py > gpurun_out/${T}_rec_step_trace.txt 2>&1; tail -n 1 gpurun_out/${T}_rec_step_trace.txt
B=64 timeout -k 10 300 python scripts/rec_trace.py > gpurun_out/${T}_rec_step_trace_b64.txt 2>&1; tail -n 1 gpurun_out/${T}_rec_step_trace_b64.txt
MODE=2 timeout -k 10 300 python scripts/bwd_trace.py > gpurun_out/${T}_bwd_step_trace.txt 2>&1; tail -n 1 gpurun_out/${T}_bwd_step_trace.txt
