#!/bin/bash
# ncu --set full capture of one kernel of the forward: gpu_ncu_one.sh <kernel regex> [launches to skip]
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out; K=$1; S=${2:-2}
python scripts/dev_forward_small.py 2048 > $O/ncu1_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -o $O/ncu1_$K -f python scripts/dev_forward_small.py 2048 > $O/ncu1_$K.log 2>&1; echo "$K ncu exit $?"
ls -la $O/ncu1_$K.ncu-rep
