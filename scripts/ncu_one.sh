#!/bin/bash
# ncu --set full capture of one kernel of the DIST pass: scripts/ncu_one.sh <tag> <kernel regex> <launch skip>
python scripts/full_pass.py > /dev/null 2>&1 || exit 1
REPS=2 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip $3 -c 1 -f \
    -o gpurun_out/x_$1 python scripts/full_pass.py > gpurun_out/ncu_x_$1.log 2>&1
tail -1 gpurun_out/ncu_x_$1.log
