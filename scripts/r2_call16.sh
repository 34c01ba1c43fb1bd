#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== dense timing"
for r in 1 2 3; do timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-130; done
echo "== step"
timeout 300 python scripts/quick_variant_check.py 2>&1 | tail -1
