#!/bin/bash
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest.log
TISEG_DEBUG_FLOOD=1 REPS=2 python scripts/full_pass.py > gpurun_out/r2_flood_debug.log 2>&1
tail -5 gpurun_out/r2_flood_debug.log
