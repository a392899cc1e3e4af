#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
timeout 300 python scripts/dev_packing_diff.py 2>&1 | tail -12
