#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
L=$PWD/synergynet_b200
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== deep weight rings"
for v in "" _var_ring ""  _var_ring; do SYN_LIB_PATH=$L/libsynergy_b200$v.so timeout 200 python scripts/quick_variant_check.py 2>&1 | tail -1; done
SYN_LIB_PATH=$L/libsynergy_b200_var_ring.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-config5 --no-single-pass 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ring ms', d['ms_per_step'], d['kernels_ms'])"
SYN_LIB_PATH=$L/libsynergy_b200_var_ring.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "not native_library" 2>&1 | tail -2
echo "== e2e chunk schedule"
for cc in "512 512" "256 768" "384 640" "512 1024"; do set -- $cc; SYN_HOST_CHUNK0=$1 SYN_HOST_CHUNK=$2 timeout 300 python - <<'PY'
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench
from synergynet_b200 import synthetic
m = bench.build_model('cuda:0'); eng = m._engine(torch.device('cuda', 0))
for kind, mk in (('f32', synthetic.make_inputs), ('u8', synthetic.make_crops_u8)):
    xh = [mk(1024, seed=100 + i).pin_memory() for i in range(2)]
    lh = torch.empty((1024, 3, 68)).pin_memory()
    for i in range(3): eng.forward_landmarks_host(xh[i % 2], lh)
    t0 = time.perf_counter()
    for i in range(40): eng.forward_landmarks_host(xh[i % 2], lh)
    dt = (time.perf_counter() - t0) / 40
    print('chunks', os.environ.get('SYN_HOST_CHUNK0'), os.environ.get('SYN_HOST_CHUNK'), kind, round(dt * 1e3, 3), 'ms', round(1024 / dt), 'faces/s')
PY
done
echo "== bench (full line)"
timeout 1200 python bench.py --steps 20 --warmup 5 > $OUT/r2_bench_call5.json 2> $OUT/r2_bench_call5.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench_call5.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e_u8']['value'])
print('config5', d.get('config5')); print('dense', d.get('dense', {}).get('ms')); print('lat', d.get('latency_b1'))
PY
tail -3 $OUT/r2_bench_call5.err
