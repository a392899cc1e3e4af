#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/dev_sharded_join_phases.py $2 > gpurun_out/r02s_phases_n$N.log 2>&1; echo "exit $?"
grep "^iter" gpurun_out/r02s_phases_n$N.log | sort
