#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== debug heads"; timeout 300 python scripts/debug_heads.py 2>&1 | tail -25
echo "== new coverage tests"; timeout 1200 python -m pytest tests/_wip_gpu_heads.py -q --tb=line 2>&1 | tail -15
echo "== dense with L2 evict_last"; timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1
SYN_DENSE_VERTEX_MAJOR=1 timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1
echo "== e2e chunk schedule"
for c0 in 64 128 256 448; do SYN_HOST_CHUNK0=$c0 timeout 300 python - <<'PY'
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
    print('chunk0', os.environ.get('SYN_HOST_CHUNK0'), kind, round(dt * 1e3, 3), 'ms', round(1024 / dt), 'faces/s')
PY
done
echo "== pytest gpu (main suite)"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
