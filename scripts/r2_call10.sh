#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== config5 (GEMM epilogue v3)"
timeout 600 python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import bench
print(json.dumps(bench.config5_measurement(torch.device('cuda', 0), bench.load_peaks())))
PY
echo "== dense timeline"
SYN_DENSE_TRACE=$OUT/r2_dense_trace.txt timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-120
python scripts/dense_trace.py $OUT/r2_dense_trace.txt | head -30
