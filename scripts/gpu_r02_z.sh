#!/bin/bash
# the weak-scaling join inside bench.py at N ranks, with and without the end-to-end section before it
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out; n=$1
for flags in "--no-cfg3 --no-cfg4 --no-cfg5 --no-cpu" "--no-e2e --no-cfg3 --no-cfg4 --no-cfg5 --no-cpu"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 5 --warmup 3 $flags > $O/r02z.json 2> $O/r02z.err; echo "[$flags] exit $?"
  python - <<PY
import json
d=json.loads(open("$O/r02z.json").read().strip().splitlines()[-1])
print("  join", round(d["join"]["ms"],2), "ms", d["join"].get("ms_each_rank0"), "frac", round(d["join"]["roofline"]["frac"],3))
PY
done
