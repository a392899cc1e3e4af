#!/bin/bash
# ncu --set full captures of the hot kernels (one launch each, after warm-up launches)
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
python scripts/dev_forward_small.py 2048 > $O/r02q_plain.log 2>&1 || exit 1
for k in stem_ts_kernel conv3_ts_kernel gemm_pair_tcgen05_kernel ffn_pair_kernel attention_fa_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 1 -o $O/r02q_$k -f python scripts/dev_forward_small.py 2048 > $O/r02q_ncu_$k.log 2>&1; echo "$k ncu exit $?"
done
python scripts/dev_join_small.py 262144 > $O/r02q_plain_join.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_ares2 -s 2 -c 1 -o $O/r02q_join_pair -f python scripts/dev_join_small.py 262144 > $O/r02q_ncu_join.log 2>&1; echo "join ncu exit $?"
ls -la $O/r02q_*.ncu-rep
