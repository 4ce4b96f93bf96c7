#!/bin/bash
# first GPU call of round 2: parity tests (incl. exact-config reference parity), bench with gpu_reference, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt 2>&1
nproc >> gpurun_out/r02a_smi.txt
( time python -m pytest tests -m gpu -x -q -s 2>&1 | tail -150 ) > gpurun_out/r02a_tests.log 2>&1
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02a_ref.json 2> gpurun_out/r02a_ref.err
tail -3 gpurun_out/r02a_tests.log
