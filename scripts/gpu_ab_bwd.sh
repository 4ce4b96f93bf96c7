#!/bin/bash
# A/B of BPTT builds on one box: scripts/gpu_ab_bwd.sh <tag> <variant> [...]   (variant "main" = the default library)
mkdir -p gpurun_out
TAG=$1; shift
for v in "$@"; do
  if [ "$v" == main ]; then L=""; else L=$PWD/onssen_b200/libonssen_b200_$v.so; fi
  for pr in 0; do
    ( ONSSEN_LIB=$L MODE=2 timeout 300 python scripts/bwd_trace.py ) > gpurun_out/${TAG}_bwd_${v}_p${pr}.txt 2>&1
    echo "$v probe=$pr: $(tail -n 1 gpurun_out/${TAG}_bwd_${v}_p${pr}.txt)"
  done
done
