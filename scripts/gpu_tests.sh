#!/bin/bash
# full GPU test suite, log to gpurun_out/<tag>_tests.log
tag=${1:-run}
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/${tag}_tests.log 2>&1
tail -6 gpurun_out/${tag}_tests.log
