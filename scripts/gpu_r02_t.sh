#!/bin/bash
# tcgen05 attention: parity against the mma.sync kernel and the oracle, then forward-only timing A/B
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python scripts/dev_tuning_parity.py 17=0 -- 17=1 2>&1 | tail -5; echo "parity exit $?"
for v in 0 1; do
  timeout 600 python bench.py --forward-only --steps 5 --warmup 3 --tuning 17=$v > $O/r02t_ab_$v.json 2> $O/r02t_ab_$v.err; echo "17=$v exit $?"
  python - <<PY
import json
for line in open("$O/r02t_ab_$v.json"):
    if line.startswith("{"):
        d=json.loads(line); st=d.get("stage_ms_per_step",{})
        print("17=$v", round(d["value"]), "videos/s", round(d["ms_per_step"],2), "ms; attention", round(st.get("attention",0),3))
PY
done
