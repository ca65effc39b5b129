#!/bin/bash
# ncu --set full of the named kernels in one step of a workload:  KERNELS='k_eqbits|k_wsl_remove' bash scripts/r2_full.sh [workload] [tag]
WL=${1:-dist_monuseg_1000}; TAG=${2:-r2_full}
REPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KERNELS}" -c ${COUNT:-12} -f -o gpurun_out/$TAG python scripts/step_pass.py $WL > gpurun_out/$TAG.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/$TAG.log; ls -la gpurun_out/$TAG.ncu-rep
