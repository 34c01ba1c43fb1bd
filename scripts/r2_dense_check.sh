#!/bin/bash
# Dense reconstruction: parity tests, timing, CTA 0's timeline.
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q -k "dense or recon or vertex or ragged or image" 2>&1 | tail -4
for r in 1 2 3; do timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-130; done
SYN_DENSE_TRACE=$OUT/r2_dense_trace.txt timeout 120 python scripts/bench_configs.py dense > /dev/null 2>&1
python scripts/dense_trace.py $OUT/r2_dense_trace.txt 2>/dev/null | head -14
