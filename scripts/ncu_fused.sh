#!/bin/bash
# One ncu --set full capture (with source counters) per fused kernel named on the command line.
# Usage: scripts/ncu_fused.sh tag "regex1" "regex2" ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for k in "$@"; do
  name=$(echo "$k" | tr -c 'A-Za-z0-9' '_' | cut -c1-40)
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$k" -s 3 -c 1 \
      -o $OUT/${TAG}_$name -f python bench.py --steps 1 --warmup 3 --profile > $OUT/${TAG}_$name.log 2>&1
  echo "$k rc=$?"
done
