#!/bin/bash
# Round-2 session M: full parity suite, full bench line, ncu launch list of a bench run
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02m_gputests.log 2>&1; echo "pytest exit $?" >> $O/r02m_gputests.log; tail -4 $O/r02m_gputests.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02m_bench_n1.json 2> $O/r02m_bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/r02m_bench_ref.json 2> $O/r02m_bench_ref.err; echo "ref exit $?"
python bench.py --forward-only --clips 2048 --steps 2 --warmup 3 > $O/r02m_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02m_launches.csv python bench.py --forward-only --clips 2048 --steps 2 --warmup 3 > $O/r02m_ncu.log 2>&1
echo "ncu exit $?"
python - <<PY
import json
d=json.loads(open("$O/r02m_bench_n1.json").read().strip().splitlines()[-1])
print(round(d["value"]), d["ms_per_step"], d["e2e"]["value"], d["roofline"]["whole_step"], d["join"]["value"], d["join"]["roofline"]["frac"])
for k in ("cfg3_varlen","cfg4_join","cfg5_topk"):
    x=d[k]; print(k, x.get("value"), x.get("ms"), x.get("roofline",{}).get("frac"), x.get("parity"), x.get("error"))
print(d["stage_ms_per_step"])
PY
