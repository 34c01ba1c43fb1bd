L=$PWD/synergynet_b200
run() { timeout 120 env "$@" python scripts/quick_kernels.py 2>&1 | tail -n 1; }
run A=1
run SYN_LIB_PATH=$L/libsynergy_b200_var_mw.so SYN_FUSED_WARPS=20
run SYN_LIB_PATH=$L/libsynergy_b200_var_mw.so SYN_FUSED_WARPS=24
run SYN_LIB_PATH=$L/libsynergy_b200_var_b3a.so
run SYN_LIB_PATH=$L/libsynergy_b200_var_b3b.so
