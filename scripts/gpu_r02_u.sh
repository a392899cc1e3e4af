#!/bin/bash
# full GPU suite, then ncu --set full of the tcgen05 attention and the temporal conv kernel
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02u_gputests.log 2>&1; echo "pytest exit $?" >> $O/r02u_gputests.log; tail -4 $O/r02u_gputests.log
python scripts/dev_forward_small.py 2048 > $O/r02u_plain.log 2>&1 || exit 1
for k in attention_tc_kernel temporal_conv_tma_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $O/r02u_$k -f python scripts/dev_forward_small.py 2048 > $O/r02u_ncu_$k.log 2>&1; echo "$k ncu exit $?"
done
ls -la $O/r02u_*.ncu-rep
