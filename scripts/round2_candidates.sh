#!/bin/bash
# Round-2 starting point: build the three options that were written but never run,
# check each against the fp32 engine and time it next to the default library.  Run the second half under gpurun:
#   scripts/round2_candidates.sh build      (here)
#   gpurun -- 'scripts/round2_candidates.sh run'
set -e
cd "$(dirname "$0")/.."
case "$1" in
  build)
    scripts/build_variant.sh epi2 -DSYN_EPI2_STAGED=1
    scripts/build_variant.sh split -DSYN_DW_SPLIT_LAST=1
    scripts/build_variant.sh pdl -DSYN_PDL=1
    ;;
  run)
    for v in "" _var_epi2 _var_split _var_pdl; do
      SYN_LIB_PATH=$PWD/synergynet_b200/libsynergy_b200$v.so timeout 120 python scripts/quick_variant_check.py 2>&1 | tail -1
    done
    ;;
  *) echo "usage: $0 build|run"; exit 1 ;;
esac
