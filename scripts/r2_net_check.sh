#!/bin/bash
# detector after the NMS-scan / ranking rewrite: parity tests of both stages, then the timing
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 300 python -m pytest tests/test_gpu_render.py tests/test_gpu_faceboxes.py -q 2>&1 | tail -8
echo "== bench"; timeout 200 python scripts/bench_render.py --no-cpu > $OUT/render_bench4.json 2> $OUT/render_bench4.err; echo rc=$?; tail -n 3 $OUT/render_bench4.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/render_bench4.json').read().strip().splitlines()[-1])
print({k: v for k, v in d['detect'].items() if k not in ('workload', 'network')})
print({k: v for k, v in d['detect']['network'].items() if k not in ('workload', 'note')})
PY
