#!/bin/bash
# ncu --set full captures of the token-path kernels (one launch each, after warm-up launches)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
python scripts/dev_forward_small.py 2048 > $O/r02r_plain.log 2>&1 || exit 1
for k in ffn_pair_kernel attention_fa_kernel layernorm; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o $O/r02r_$k -f python scripts/dev_forward_small.py 2048 > $O/r02r_ncu_$k.log 2>&1; echo "$k ncu exit $?"
done
ls -la $O/r02r_*.ncu-rep
