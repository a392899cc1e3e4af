#!/bin/bash
# full GPU suite + forward-only bench + join bench
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02v_gputests.log 2>&1; echo "pytest exit $?" >> $O/r02v_gputests.log; tail -4 $O/r02v_gputests.log
for i in 1 2; do
timeout 600 python bench.py --forward-only --steps 5 --warmup 3 > $O/r02v_fwd.json 2> $O/r02v_fwd.err; echo "fwd exit $?"
python - <<PY
import json
for line in open("$O/r02v_fwd.json"):
    if line.startswith("{"):
        d=json.loads(line); st=d.get("stage_ms_per_step",{})
        print(round(d["value"]), "videos/s", round(d["ms_per_step"],2), "ms;", {k: round(v,2) for k,v in st.items()})
PY
done
timeout 600 python scripts/dev_join_small.py 262144 2>&1 | tail -3
