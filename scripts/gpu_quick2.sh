bash scripts/gpu_quick.sh
bash scripts/ncu_one.sh ccl4 k_ccl_local 6
