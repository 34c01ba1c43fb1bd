#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu (all; GEMM v2, unpaired dense)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== dense: streaming vs write-back stores"
for i in 1 2; do
  timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-150
  SYN_DENSE_WB_STORES=1 timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-150
done
echo "== config5 with GEMM v2"
timeout 600 python - <<'PY'
import os, sys, json, torch
sys.path.insert(0, os.getcwd())
import bench
print(json.dumps(bench.config5_measurement(torch.device('cuda', 0), bench.load_peaks())))
PY
echo "== ncu dense (write-back stores if faster is decided later; default here)"
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:dense_recon_fm" -s 5 -c 1 -o $OUT/r2_dense3 -f python scripts/bench_configs.py dense > $OUT/r2_ncu_dense3.log 2>&1; echo rc=$?
SYN_DENSE_WB_STORES=1 timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:dense_recon_fm" -s 5 -c 1 -o $OUT/r2_dense3wb -f python scripts/bench_configs.py dense > $OUT/r2_ncu_dense3wb.log 2>&1; echo rc=$?
echo "== ncu gemm family"
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:tc_gemm_kernel|resnet_stem|maxpool3x3|avgpool_kernel|small_k_layer|wing_loss|param_loss|pose_decode" -c 16 -o $OUT/r2_gemm -f python scripts/sanitizer_smoke.py > $OUT/r2_ncu_gemm.log 2>&1; echo rc=$?
ls -la $OUT/*.ncu-rep
