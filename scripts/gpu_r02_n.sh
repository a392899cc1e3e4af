#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_guards_gpu.py -m gpu -x -q > gpurun_out/r02n_guards.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n_guards.log; tail -15 gpurun_out/r02n_guards.log
