#!/bin/bash
# 8-GPU session: NCCL parity test at world 8, bench at N = 8 and N = 4
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi topo -m > $O/r02w_topo.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q -rs > $O/r02w_multigpu_test.log 2>&1; echo "pytest exit $?" >> $O/r02w_multigpu_test.log; tail -4 $O/r02w_multigpu_test.log
for n in 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 > $O/r02w_bench_n$n.json 2> $O/r02w_bench_n$n.err; echo "bench n=$n exit $?"
python - <<PY
import json
try:
    d=json.loads(open("$O/r02w_bench_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", round(d["value"]), "videos/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"]), "; join", round(d["join"]["value"]), d["join"]["roofline"]["frac"])
    for k in ("cfg3_varlen","cfg4_join","cfg5_topk"):
        x=d[k]; print(" ", k, x.get("value"), x.get("ms"), x.get("roofline",{}).get("frac"), x.get("parity"), x.get("error"))
except Exception as e:
    print("N=$n: no line", e)
PY
done
