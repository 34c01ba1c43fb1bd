#!/bin/bash
# Mid-round check in ONE python start-up where possible: the new pipeline test + the parity subset touched by the
# block-3 retiling, then the render / detect / detector-network timing.
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest"; timeout 400 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_parity.py -q -x 2>&1 | tail -8
echo "== bench render"; timeout 200 python scripts/bench_render.py --no-cpu > $OUT/render_bench2.json 2> $OUT/render_bench2.err; echo rc=$?; tail -n 3 $OUT/render_bench2.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/render_bench2.json').read().strip().splitlines()[-1])
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d['render'].items() if k.endswith('_ms') or k.endswith('per_s')})
print({k: v for k, v in d['detect'].items() if k != 'workload'})
PY
