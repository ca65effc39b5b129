#!/bin/bash
python scripts/full_pass.py > /dev/null 2>&1 || exit 1
REPS=2 ncu --set full --clock-control none --import-source on -k regex:k_pair_local --launch-skip 2 -c 1 -f \
    -o gpurun_out/r2_pair_local python scripts/full_pass.py > gpurun_out/ncu_pair_local.log 2>&1
tail -2 gpurun_out/ncu_pair_local.log
