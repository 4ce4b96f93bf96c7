#!/bin/bash
mkdir -p gpurun_out
STEPS=2 timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_train.csv python scripts/train_step.py > gpurun_out/ncu_train.log 2>&1; echo "ncu exit=$?"; tail -1 gpurun_out/ncu_train.log
