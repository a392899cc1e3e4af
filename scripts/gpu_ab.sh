#!/bin/bash
# A/B of one tuning key: parity of setting 1 against setting 0 and the oracle, then forward-only timing of both
# usage: gpu_ab.sh KEY [stage-name] [value A] [value B]
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out; K=$1; STAGE=${2:-stem_fused}; VA=${3:-0}; VB=${4:-1}
timeout 300 python scripts/dev_tuning_parity.py $K=$VA -- $K=$VB 2>&1 | tail -5; echo "parity exit $?"
for v in $VA $VB $VA $VB; do
  timeout 600 python bench.py --forward-only --steps 5 --warmup 3 --tuning $K=$v > $O/ab_${K}_$v.json 2> $O/ab_${K}_$v.err; echo "$K=$v exit $?"
  python - <<PY
import json
for line in open("$O/ab_${K}_$v.json"):
    if line.startswith("{"):
        d=json.loads(line); st=d.get("stage_ms_per_step",{})
        print("$K=$v", round(d["value"]), "videos/s", round(d["ms_per_step"],2), "ms; $STAGE", round(st.get("$STAGE",0),3))
PY
done
