#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_backward_gpu.py -q -m gpu -p no:cacheprovider -s -x > gpurun_out/pytest_bwd.log 2>&1; echo "bwd tests exit=$?"
tail -n 60 gpurun_out/pytest_bwd.log
