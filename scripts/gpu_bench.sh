#!/bin/bash
# Full GPU pass: tests, smoke, bench (ours + reference arm), ncu launch list and one full capture of the
# recurrent kernel. Outputs in gpurun_out/.
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/summary.txt
tail -n 5 gpurun_out/pytest_gpu.log
timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/summary.txt
tail -n 3 gpurun_out/smoke.log
timeout -k 10 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
if [ "$1" != "noncu" ]; then
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit=$?" | tee -a gpurun_out/summary.txt
timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:blstm_rec -s 9 -c 3 -f -o gpurun_out/prof_rec python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_rec.log 2>&1; echo "ncu rec exit=$?" | tee -a gpurun_out/summary.txt
fi
cat gpurun_out/summary.txt
