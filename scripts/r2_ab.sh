#!/bin/bash
# A/B of an environment switch on the per-kernel step times:  AB='TISEG_X=1' bash scripts/r2_ab.sh
timeout 200 python scripts/step_times.py > gpurun_out/r2_times_a.log 2>&1; echo "A rc=$?"; head -${LINES_SHOWN:-14} gpurun_out/r2_times_a.log
env $AB timeout 200 python scripts/step_times.py > gpurun_out/r2_times_b.log 2>&1; echo "B ($AB) rc=$?"; head -${LINES_SHOWN:-14} gpurun_out/r2_times_b.log
