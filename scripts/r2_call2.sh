#!/bin/bash
# round-2 GPU call 2: new default library (DW3, padded taps, residual prefetch, PDL, face-major dense, engine 3, flags)
OUT=gpurun_out; mkdir -p $OUT
L=$PWD/synergynet_b200
echo "== pytest gpu (main lib)"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
echo "== variants"
for v in "" _var_dw3off _var_s5 _var_nopdl _var_occ2; do
  SYN_DEBUG_OCC=${v:+1} SYN_LIB_PATH=$L/libsynergy_b200$v.so timeout 200 python scripts/quick_variant_check.py 2>&1 | grep -E "ms/step|syn\] fused|Error|error" | tail -14
done
SYN_FUSED_WARPS=12 SYN_LIB_PATH=$L/libsynergy_b200_var_s5.so timeout 200 python scripts/quick_variant_check.py 2>&1 | tail -1
SYN_FUSED_WARPS=12 timeout 200 python scripts/quick_variant_check.py 2>&1 | tail -1
echo "== variant parity (s5, occ2)"
for v in _var_s5 _var_occ2; do
  SYN_LIB_PATH=$L/libsynergy_b200$v.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "not native_library" 2>&1 | tail -3
done
echo "== dense: face-major vs vertex-major"
timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1
SYN_DENSE_VERTEX_MAJOR=1 timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/r2_bench_call2.json 2> $OUT/r2_bench_call2.err; tail -c 6000 $OUT/r2_bench_call2.json; tail -5 $OUT/r2_bench_call2.err
echo "== trace"
SYN_LIB_PATH=$L/libsynergy_b200_trace.so timeout 200 python scripts/fused_trace.py 1 2 3 5 8 12 > $OUT/r2_trace_call2.txt 2>&1; grep -A3 "== block" $OUT/r2_trace_call2.txt | head -60
