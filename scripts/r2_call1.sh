#!/bin/bash
# round-2 GPU call 1: probes, never-run variants, worker-warp sweep, phase trace, baseline tests
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
echo "== ffma2 probe"; timeout 60 tools/ffma2_probe | tee $OUT/r2_ffma2_probe.txt
echo "== variants"; scripts/round2_candidates.sh run
echo "== worker warps"
for w in 8 12; do SYN_FUSED_WARPS=$w timeout 120 python scripts/quick_variant_check.py 2>&1 | tail -1; done
echo "== trace"
SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200_trace.so timeout 200 python scripts/fused_trace.py 1 2 3 4 5 8 12 15 > $OUT/r2_trace_base.txt 2>&1; tail -5 $OUT/r2_trace_base.txt
echo "== pytest gpu"; time (timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4)
echo "== dense v2 (elect issuer in dense + tail)"
for v in "" _var_dense2; do
  SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200$v.so timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1
  SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200$v.so timeout 120 python scripts/quick_variant_check.py 2>&1 | tail -1
done
SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200_var_dense2.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "== occ2 (two CTAs per SM)"
SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200_var_occ2.so timeout 200 python scripts/quick_variant_check.py 2>&1 | tail -2
SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200_var_occ2.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('occ2 ms', d['ms_per_step'], d['kernels_ms'])"
SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200_var_occ2.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "== ffma2 / occ2f"
for v in _var_ffma2 _var_occ2f; do
  SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200$v.so timeout 200 python scripts/quick_variant_check.py 2>&1 | tail -1
done
SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200_var_occ2f.so timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('occ2f ms', d['ms_per_step'], d['kernels_ms'])"
SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200_var_occ2f.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
