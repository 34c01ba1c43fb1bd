#!/bin/bash
# Closing run of round 2 (second session): the whole GPU test-suite, the bench line, and ncu
# evidence for the kernels added in this session (Sim3DR, FaceBoxes): launch list + one `--set full` capture of each.
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
echo "== bench"; timeout 900 python bench.py > $OUT/r2b_bench.json 2> $OUT/r2b_bench.err; echo rc=$?; tail -n 2 $OUT/r2b_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2b_bench.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e_u8']['value'], 'launches', d['gpu_launches'])
print({k: v for k, v in list(d['kernels_ms'].items())[:6]}); print('roofline', d['roofline']['frac_hbm'], d['roofline']['frac_tensor'])
print('dense', d.get('dense', {}).get('ms')); print('render', {k: v for k, v in d.get('render', {}).items() if k.endswith('_ms')})
print('detect', {k: v for k, v in d.get('detect', {}).items() if k.endswith('_ms')}, d.get('detect', {}).get('network'))
print('cpu', d.get('cpu_baseline')); print('clocks', d.get('clocks'))
PY
echo "== ncu launch list (render / detect)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:syn:: --csv --log-file $OUT/r2_render_launches.csv \
    python scripts/ncu_render_once.py > $OUT/ncu_render_list.log 2>&1; echo rc=$?
echo "== ncu full (render / detect)"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:syn:: -c 26 -o $OUT/r2_render_full -f \
    python scripts/ncu_render_once.py > $OUT/ncu_render_full.log 2>&1; echo rc=$?; ls -la $OUT/r2_render_full.ncu-rep
