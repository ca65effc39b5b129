#!/bin/bash
for gx in 16 64 244; do
TISEG_SEM_GX=$gx timeout 120 python scripts/step_times.py 2>&1 | grep "step wall\|sem_counts\|dist_prep\|flood_par" > gpurun_out/r2_sem_$gx.log
done
