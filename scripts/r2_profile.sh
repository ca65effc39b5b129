#!/bin/bash
# round-2 profile pass: whole-step per-kernel traffic of every workload, the launch list of the bench command, and one
# ncu --set full capture of the top kernels of the headline workload
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum
for wl in dist_monuseg_1000 unet_cpm17_256 conic_sweep_256 cdnet_consep_1000 hover_consep_1000; do
  timeout 300 ncu --metrics $M --clock-control none -c 2000 --csv --log-file gpurun_out/r2_step_$wl.csv python scripts/step_pass.py $wl > gpurun_out/r2_step_$wl.log 2>&1
  echo "$wl rc=$? $(tail -1 gpurun_out/r2_step_$wl.log)"
done
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_b21.json 2> gpurun_out/r2_b21.err; echo "bench(2,1) rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_launches.log 2>&1; echo "launch list rc=$?"
KERNELS='k_ws_flood_par|k_eqbits|k_bitccl_tile|k_wsl_remove|k_plateau_bits|k_pair_bits|k_argmax_logits2|k_rank_fused|k_marker_scatter|k_dist_prep|k_sem_counts' COUNT=24 bash scripts/r2_full.sh dist_monuseg_1000 r2_full_top
