#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
VFP_B200_LIB=build/libvfp_trace.so timeout 300 python scripts/dev_ffn_trace.py > gpurun_out/r02h_trace.txt 2>&1; echo "exit $?"; head -170 gpurun_out/r02h_trace.txt
