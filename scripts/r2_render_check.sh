#!/bin/bash
# First GPU run of the Sim3DR / FaceBoxes post-processing kernels: parity tests, smoke(), the render / detect timing.
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest render"; timeout 600 python -m pytest tests/test_gpu_render.py -q -x 2>&1 | tail -15
echo "== bench render"; timeout 300 python scripts/bench_render.py > $OUT/render_bench.json 2> $OUT/render_bench.err; echo rc=$?; tail -3 $OUT/render_bench.err; cat $OUT/render_bench.json | cut -c1-3000
