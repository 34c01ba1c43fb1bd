#!/bin/bash
# Round-2 closing run on one B200: full GPU test-suite, the bench line (profiles/r2_bench.json is a copy of its output),
# the reference arm, smoke().
OUT=gpurun_out; mkdir -p $OUT
echo "== smoke"; timeout 600 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench"; timeout 1500 python bench.py > $OUT/r2_bench.json 2> $OUT/r2_bench.err; echo rc=$?; tail -2 $OUT/r2_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e_u8']['value'], 'launches', d['gpu_launches'])
print(d['kernels_ms']); print('roofline', d['roofline']); print('dense', d.get('dense')); print('config5', d.get('config5'))
print('cpu', d.get('cpu_baseline')); print('clocks', d.get('clocks'))
PY
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 3 2>/dev/null | tail -1 | cut -c1-400
