#!/bin/bash
mkdir -p gpurun_out
for bo in 0 40 100 200 400; do
 for v in "" _rec01; do
  ( BACKOFF=$bo ONSSEN_LIB=onssen_b200/libonssen_b200$v.so timeout 120 python scripts/rec_trace.py ) > gpurun_out/r02e_trace${v}_bo$bo.txt 2>&1
  echo "backoff $bo lib '$v': $(tail -n 1 gpurun_out/r02e_trace${v}_bo$bo.txt)"
 done
done
for pd in 300 600 900; do
  ( POLL_DELAY=$pd timeout 120 python scripts/rec_trace.py ) > gpurun_out/r02e_trace_pd$pd.txt 2>&1
  echo "poll_delay $pd: $(tail -n 1 gpurun_out/r02e_trace_pd$pd.txt)"
done
