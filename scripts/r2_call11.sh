#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
L=$PWD/synergynet_b200
echo "== heads / resnet tests (warp-reduced max-pool)"; timeout 900 python -m pytest tests/test_gpu_heads.py -x -q 2>&1 | tail -3
echo "== config5"
timeout 600 python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import bench
print(json.dumps(bench.config5_measurement(torch.device('cuda', 0), bench.load_peaks())))
PY
echo "== dense: bulk copies per plane 1 / 4 / 8"
for r in 1 2; do for v in "" _var_split4 _var_split8; do
  SYN_LIB_PATH=$L/libsynergy_b200$v.so timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-130
done; done
SYN_LIB_PATH=$L/libsynergy_b200_var_split4.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "dense" 2>&1 | tail -2
