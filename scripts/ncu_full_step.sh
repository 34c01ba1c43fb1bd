#!/bin/bash
# ONE `ncu --set full` capture of every launch of a device-resident step (17 fused blocks, tail, heads, alpha pre-pass,
# sparse reconstruction = 21 launches), after 3 warm-up steps; summarise here with scripts/ncu_report_summary.py.
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -s 63 -c 21 \
    -o $OUT/r2_step_full -f python bench.py --steps 1 --warmup 3 --profile > $OUT/r2_step_full.log 2>&1; echo rc=$?
ls -la $OUT/r2_step_full.ncu-rep
