#!/bin/bash
# stem timing with every UMMA reading half of its weight operand (development build, wrong results) against the product build
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
for lib in video_fingerprint_b200/libvfp_b200.so build/libvfp_halfb.so video_fingerprint_b200/libvfp_b200.so build/libvfp_halfb.so; do
  VFP_B200_LIB=$PWD/$lib timeout 600 python bench.py --forward-only --steps 5 --warmup 3 > $O/r02y.json 2> $O/r02y.err; echo "lib='$lib' exit $?"
  python - <<PY
import json
for line in open("$O/r02y.json"):
    if line.startswith("{"):
        d=json.loads(line); st=d.get("stage_ms_per_step",{})
        print("  ", round(d["ms_per_step"],2), "ms; stem", round(st.get("stem_fused",0),3))
PY
done
