#!/bin/bash
# 2-GPU session: the NCCL parity test + a short 2-rank bench line
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi topo -m > $O/r02c_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -rs > $O/r02c_multigpu_test.log 2>&1; echo "pytest exit $?" >> $O/r02c_multigpu_test.log; tail -6 $O/r02c_multigpu_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r02c_bench_n2.json 2> $O/r02c_bench_n2.err; echo "bench exit $?"; tail -c 2500 $O/r02c_bench_n2.json; tail -5 $O/r02c_bench_n2.err
