#!/bin/bash
# whole-step per-kernel DRAM traffic of every workload (ncu metrics pass) + the bench line of hover
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum
for wl in dist_monuseg_1000 unet_cpm17_256 conic_sweep_256 cdnet_consep_1000 hover_consep_1000; do
  timeout 300 ncu --metrics $M --clock-control none -c 2000 --csv --log-file gpurun_out/r2_step_$wl.csv python scripts/step_pass.py $wl > gpurun_out/r2_step_$wl.log 2>&1
  tail -1 gpurun_out/r2_step_$wl.log
done
TISEG_BENCH_ALLOW_UNKNOWN=1 timeout 400 python bench.py --workload hover_consep_1000 --steps 8 --warmup 3 --distinct 4 > gpurun_out/r2_wl_hover_consep_1000.json 2> gpurun_out/r2_wl_hover_consep_1000.err
echo "hover rc=$?"
