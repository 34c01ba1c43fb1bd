#!/bin/bash
# Run on the B200 box (via gpurun): smoke, GPU parity tests, bench, ncu launch list + one full capture.
# Usage: scripts/gpu_check.sh [tag]   -> everything lands in gpurun_out/<tag>_*
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== smoke" ; timeout 600 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 $OUT/${TAG}_smoke.log
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest.log 2>&1 ; echo "pytest rc=$?" ; tail -15 $OUT/${TAG}_pytest.log
echo "== bench" ; timeout 900 python bench.py --steps 30 --warmup 5 $BENCH_ARGS > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err ; echo "bench rc=$?" ; cat $OUT/${TAG}_bench.json ; tail -3 $OUT/${TAG}_bench.err
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --profile $BENCH_ARGS > $OUT/${TAG}_ncu_launches.log 2>&1 ; echo "ncu-list rc=$?"
if [ -n "$NCU_KERNEL" ]; then
  echo "== ncu full: $NCU_KERNEL"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$NCU_KERNEL -s ${NCU_SKIP:-20} -c ${NCU_COUNT:-3} \
      -o $OUT/${TAG}_prof -f python bench.py --steps 1 --warmup 3 --profile $BENCH_ARGS > $OUT/${TAG}_ncu_full.log 2>&1 ; echo "ncu-full rc=$?"
fi
