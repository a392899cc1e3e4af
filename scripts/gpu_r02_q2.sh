#!/bin/bash
# ncu --set full of the CTA-pair kernels after the remote-arrive change (conv4 and the join)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
python scripts/dev_forward_small.py 2048 > $O/r02q2_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gemm_pair_tcgen05_kernel -s 12 -c 1 -o $O/r02q2_conv4 -f python scripts/dev_forward_small.py 2048 > $O/r02q2_ncu_conv4.log 2>&1; echo "conv4 ncu exit $?"
python scripts/dev_join_small.py 262144 > $O/r02q2_plain_join.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_ares2 -s 2 -c 1 -o $O/r02q2_join -f python scripts/dev_join_small.py 262144 > $O/r02q2_ncu_join.log 2>&1; echo "join ncu exit $?"
ls -la $O/r02q2_*.ncu-rep
