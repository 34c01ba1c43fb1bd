#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
L=$PWD/synergynet_b200
echo "== dense: six plane slots (main) ; + 4 copies per plane"
for r in 1 2; do for v in "" _var_six4; do
  SYN_LIB_PATH=$L/libsynergy_b200$v.so timeout 120 python scripts/bench_configs.py dense 2>&1 | tail -1 | cut -c1-130
done; done
echo "== dense timeline (six slots)"
SYN_DENSE_TRACE=$OUT/r2_dense_trace.txt timeout 120 python scripts/bench_configs.py dense > /dev/null 2>&1
python scripts/dense_trace.py $OUT/r2_dense_trace.txt 2>/dev/null | head -16
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
