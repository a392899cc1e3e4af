#!/bin/bash
# N-GPU bench line: gpu_r02_x.sh N
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out; n=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 10 --warmup 3 > $O/r02x_bench_n$n.json 2> $O/r02x_bench_n$n.err; echo "bench n=$n exit $?"
python - <<PY
import json
d=json.loads(open("$O/r02x_bench_n$n.json").read().strip().splitlines()[-1])
print("N=$n", round(d["value"]), "videos/s", round(d["ms_per_step"],2), "ms; e2e", round(d["e2e"]["value"]), "; join", round(d["join"]["value"]), d["join"]["roofline"]["frac"])
for k in ("cfg3_varlen","cfg4_join","cfg5_topk"):
    x=d[k]; print(" ", k, x.get("value"), x.get("ms"), x.get("roofline",{}).get("frac"), x.get("parity"), x.get("error"))
PY
