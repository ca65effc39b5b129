#!/bin/bash
TISEG_PROF_FLOOD=1 REPS=2 timeout 120 python scripts/full_pass.py 2>&1 | tail -5 > gpurun_out/r2_flood_debug.log
