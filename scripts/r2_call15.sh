#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
scripts/r2_final.sh
echo "== ncu launch list of the bench step"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_launches.csv python bench.py --steps 2 --warmup 1 --profile > $OUT/r2_launches.log 2>&1; echo rc=$?
echo "== ncu dense (final kernel)"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:dense_recon_fm" -s 5 -c 1 -o $OUT/r2_dense6 -f python scripts/bench_configs.py dense > $OUT/r2_ncu_dense6.log 2>&1; echo rc=$?
echo "== variant: mbarrier spin inlined"
for r in 1 2; do
  timeout 300 python scripts/quick_variant_check.py 2>&1 | tail -1
  SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200_var_mbinl.so timeout 300 python scripts/quick_variant_check.py 2>&1 | tail -1
done
