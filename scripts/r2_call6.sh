#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
L=$PWD/synergynet_b200
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== dense: paired face-major vs vertex-major"
for i in 1 2; do timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-160; done
SYN_DENSE_VERTEX_MAJOR=1 timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-160
echo "== eb2 variant"
for v in "" _var_eb2 "" _var_eb2; do SYN_LIB_PATH=$L/libsynergy_b200$v.so timeout 200 python scripts/quick_variant_check.py 2>&1 | tail -1; done
echo "== compute-sanitizer memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python scripts/sanitizer_smoke.py > $OUT/r2_sanitizer.log 2>&1; echo "sanitizer rc=$?"; tail -6 $OUT/r2_sanitizer.log | cut -c1-300
echo "== ncu: dense fm + gemm kernels"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:dense_recon_fm" -s 5 -c 1 -o $OUT/r2_dense2 -f python scripts/bench_configs.py dense > $OUT/r2_ncu_dense2.log 2>&1; echo rc=$?
timeout 900 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:tc_gemm_kernel|resnet_stem|maxpool3x3|avgpool_kernel|small_k_layer|wing_loss|param_loss|pose_decode" -c 40 -o $OUT/r2_gemm -f python scripts/sanitizer_smoke.py > $OUT/r2_ncu_gemm.log 2>&1; echo rc=$?
ls -la $OUT/*.ncu-rep
