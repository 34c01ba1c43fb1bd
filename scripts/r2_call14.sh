#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== ncu dense (source counters)"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:dense_recon_fm" -s 5 -c 1 -o $OUT/r2_dense5 -f python scripts/bench_configs.py dense > $OUT/r2_ncu_dense5.log 2>&1; echo rc=$?
ls -la $OUT/r2_dense5.ncu-rep
