#!/bin/bash
# Quick GPU check of a kernel change: fused-engine parity tests, one bench line, optional phase trace.
# Usage: scripts/gpu_quick.sh tag [trace blocks...]
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_$TAG.json").read().strip().splitlines()[-1])
print("ms", round(d["ms_per_step"], 4), "faces/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e_u8", round(d["e2e_u8"]["value"]))
print({k.replace("fused_", ""): round(v, 3) for k, v in d["kernels_ms"].items()})
PY
if [ $# -gt 0 ] && [ -f synergynet_b200/libsynergy_b200_trace.so ]; then
  SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200_trace.so timeout 200 python scripts/fused_trace.py "$@" > $OUT/trace_$TAG.txt 2>&1
  cat $OUT/trace_$TAG.txt | head -120
fi
