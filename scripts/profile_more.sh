#!/bin/bash
# Full ncu captures of the next four kernels of the DIST step (after scripts/profile_round.sh): flatten, plateau bitmaps,
# pair table, watershed-line removal.  Digest with scripts/digest_profiles.py.
R=${1:-r1}
python scripts/full_pass.py > /dev/null 2>&1 || exit 1
cap() {  # tag regex skip
    REPS=2 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 -c 1 -f \
        -o gpurun_out/${R}_$1 python scripts/full_pass.py > gpurun_out/ncu_$1_${R}.log 2>&1
    tail -1 gpurun_out/ncu_$1_${R}.log
}
cap flatten k_ccl_flatten 6
cap plateau k_plateau_bits 1
cap pair k_pair_accumulate$ 1
cap wsl_remove k_wsl_remove 2
