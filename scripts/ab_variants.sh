#!/bin/bash
# A/B of fused-kernel variants on one box: bench each library twice, interleaved.
# Usage: scripts/ab_variants.sh lib1.so lib2.so ...   (paths relative to synergynet_b200/)
OUT=gpurun_out
mkdir -p $OUT
for round in 1 2; do
  for lib in "$@"; do
    name=$(basename $lib .so)
    SYN_LIB_PATH=$PWD/synergynet_b200/$lib timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline \
        > $OUT/ab_${name}_$round.json 2> $OUT/ab_${name}_$round.err
    python - <<PY
import json
try:
    d = json.loads(open("$OUT/ab_${name}_$round.json").read().strip().splitlines()[-1])
    print("$name", $round, "ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], "kern", {k: round(v, 3) for k, v in sorted(d.get("kernels_ms", {}).items(), key=lambda kv: -kv[1])[:6]})
except Exception as e:
    print("$name", $round, "FAILED", e)
PY
  done
done
