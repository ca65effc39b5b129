#!/bin/bash
TISEG_PAIR_DBG=1 python scripts/step_times.py 2>&1 | head -3 > gpurun_out/r2_dbg1.log
TISEG_PAIR_DBG=2 python scripts/step_times.py 2>&1 | head -3 > gpurun_out/r2_dbg2.log
