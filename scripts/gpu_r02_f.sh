#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
python scripts/dev_forward_small.py 2048 14=1 > $O/r02f_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ffn_kernel -s 4 -c 2 -o $O/r02f_ffn -f python scripts/dev_forward_small.py 2048 14=1 > $O/r02f_ncu.log 2>&1
echo "ncu exit $?"; tail -3 $O/r02f_plain.log; tail -5 $O/r02f_ncu.log; ls -la $O/r02f_ffn.ncu-rep
