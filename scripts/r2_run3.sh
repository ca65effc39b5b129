#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest.log
timeout 120 python scripts/step_times.py > gpurun_out/r2_times_new.log 2>&1
TISEG_PROF_FLOOD=1 REPS=2 timeout 120 python scripts/full_pass.py 2>&1 | tail -3 > gpurun_out/r2_flood_debug.log
