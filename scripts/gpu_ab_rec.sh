#!/bin/bash
# A/B of forward-recurrence builds on one box: scripts/gpu_ab_rec.sh <tag> <variant> [<variant> ...]
# (variant "main" = onssen_b200/libonssen_b200.so, else onssen_b200/libonssen_b200_<variant>.so)
mkdir -p gpurun_out
TAG=$1; shift
for v in "$@"; do
  if [ "$v" == main ]; then L=""; else L=$PWD/onssen_b200/libonssen_b200_$v.so; fi
  for b in ${BATCHES:-32}; do
    ( ONSSEN_LIB=$L B=$b timeout 120 python scripts/rec_trace.py ) > gpurun_out/${TAG}_trace_${v}_b${b}.txt 2>&1
    echo "$v B=$b: $(tail -n 1 gpurun_out/${TAG}_trace_${v}_b${b}.txt)"
    grep -A3 "step 102 (cycles" gpurun_out/${TAG}_trace_${v}_b${b}.txt | tail -n 3
  done
done
