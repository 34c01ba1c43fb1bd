#!/bin/bash
# round-2 GPU call 3: default = 2x2 depthwise + padded taps + PDL; new coverage (PointNet heads, losses, f1, ResNet-50); ncu evidence
OUT=gpurun_out; mkdir -p $OUT
L=$PWD/synergynet_b200
echo "== pytest gpu (main suite)"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== new coverage tests"; timeout 1200 python -m pytest tests/_wip_gpu_heads.py -q -s 2>&1 | tail -40
echo "== variants"
timeout 200 python scripts/quick_variant_check.py 2>&1 | tail -1
SYN_DEBUG_OCC=1 SYN_LIB_PATH=$L/libsynergy_b200_var_occ2.so timeout 200 python scripts/quick_variant_check.py 2>&1 | grep -E "ms/step|sized for 2" | tail -12
SYN_LIB_PATH=$L/libsynergy_b200_var_occ2.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "not native_library" 2>&1 | tail -2
echo "== bench"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/r2_bench_call3.json 2> $OUT/r2_bench_call3.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_call3.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e_u8']['value'])
print(d['kernels_ms']); print('dense', d.get('dense')); print('lat', d.get('latency_b1')); print('1pass', d.get('single_pass_fp16'))
PY
tail -3 $OUT/r2_bench_call3.err
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>/dev/null | tail -1 | cut -c1-600
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r2_launches.csv python bench.py --steps 2 --warmup 3 --profile > $OUT/r2_ncu_launches.log 2>&1; echo rc=$?
echo "== ncu full: one step of the fused path"
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:fused_mbconv|tail_conv|heads_kernel|dense_" --launch-skip 63 --launch-count 21 -o $OUT/r2_step -f python bench.py --steps 1 --warmup 3 --profile > $OUT/r2_ncu_step.log 2>&1; echo rc=$?; ls -la $OUT/r2_step.ncu-rep
echo "== ncu full: dense mesh"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:dense_recon_fm" -s 5 -c 1 -o $OUT/r2_dense -f python scripts/bench_configs.py dense > $OUT/r2_ncu_dense.log 2>&1; echo rc=$?
echo "== trace"
SYN_LIB_PATH=$L/libsynergy_b200_trace.so timeout 200 python scripts/fused_trace.py 1 2 3 5 8 12 > $OUT/r2_trace_call3.txt 2>&1; grep -A3 "== block" $OUT/r2_trace_call3.txt | head -40
