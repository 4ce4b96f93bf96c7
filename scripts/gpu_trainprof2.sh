#!/bin/bash
mkdir -p gpurun_out
STEPS=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:lstm_bwd_step -s 600 -c 2 -f -o gpurun_out/prof_bwdstep python scripts/train_step.py > gpurun_out/ncu_bwdstep.log 2>&1; echo "ncu exit=$?"; tail -2 gpurun_out/ncu_bwdstep.log
