#!/bin/bash
# N-GPU check of bench.py (run under `gpurun --gpus N`): weak scaling line with the overlapped all-gather and the
# on-hardware verification that the gathered landmarks equal the single-GPU result.
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/r2_bench_${N}gpu.json 2> $OUT/r2_bench_${N}gpu.err; echo rc=$?
tail -3 $OUT/r2_bench_${N}gpu.err | cut -c1-300
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'verify', d.get('verify'), 'e2e', d['e2e']['value'], d['e2e_u8']['value'])
PY
