#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
for t in 13=1 13=2 13=3; do timeout 300 python scripts/dev_tuning_parity.py $t 2>&1 | tail -3; done
ab() { name=$1; shift; timeout 300 python bench.py --forward-only --steps 5 --warmup 3 "$@" > $O/r02d_ab_$name.json 2> $O/r02d_ab_$name.err; echo "$name exit $?"; python - <<PY
import json
try:
    d=json.loads(open("$O/r02d_ab_$name.json").read().strip().splitlines()[-1])
    st=d["stage_ms_per_step"]
    print("$name", round(d["value"]), "videos/s", round(d["ms_per_step"],2), "ms/step; conv3", round(st.get("conv3_igemm",0),2), "conv4", round(st.get("conv4_igemm_pool",0),2), "stem", round(st.get("stem_fused",0),2))
except Exception as e:
    print("$name: no line", e)
PY
}
ab base
ab mc3 --tuning 13=1
ab mc4 --tuning 13=2
ab mc34 --tuning 13=3
