#!/bin/bash
# usage: gpu_r02_san.sh memcheck|racecheck|synccheck
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
timeout 200 python scripts/sanitize_smoke.py > gpurun_out/r02_san_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r02_san_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $1 --print-limit 30 python scripts/sanitize_smoke.py > gpurun_out/r02_sanitizer_$1.log 2>&1; echo "sanitizer exit $?"
tail -12 gpurun_out/r02_sanitizer_$1.log
