#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python scripts/dev_join_ab.py 3000 262144 1048576 > $O/r02g_join_ab.jsonl 2> $O/r02g_join_ab.err; echo "join_ab exit $?"; cat $O/r02g_join_ab.jsonl; tail -5 $O/r02g_join_ab.err
