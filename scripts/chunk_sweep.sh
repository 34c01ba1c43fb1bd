#!/bin/bash
# e2e (host-buffer call) against the pipeline chunk size.  Usage: scripts/chunk_sweep.sh 128 256 512 1024
OUT=gpurun_out; mkdir -p $OUT
python scripts/bench_configs.py h2d 2>&1 | tail -3
for c in "$@"; do
  SYN_HOST_CHUNK=$c timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/chunk_$c.json 2> $OUT/chunk_$c.err
  python - <<PY
import json
d = json.loads(open("$OUT/chunk_$c.json").read().strip().splitlines()[-1])
print("chunk $c ms", round(d["ms_per_step"], 3), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e_u8", round(d["e2e_u8"]["value"]))
PY
done
