#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu (band-aligned dense is the default now)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== dense"
for i in 1 2 3; do timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-150; done
SYN_DENSE_VERTEX_MAJOR=1 timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-150
echo "== ncu dense"
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:dense_recon_fm" -s 5 -c 1 -o $OUT/r2_dense4 -f python scripts/bench_configs.py dense > $OUT/r2_ncu_dense4.log 2>&1; echo rc=$?
