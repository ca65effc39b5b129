#!/bin/bash
for v in fused split; do
  if [ $v = split ]; then export TISEG_RANK_SPLIT=1; fi
  timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2_b_$v.json 2> gpurun_out/r2_b_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/r2_b_$v.json').read()); print('$v', d['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['kernel_sum_ms_per_step'])"
done
