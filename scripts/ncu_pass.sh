# per-kernel ncu metrics of one DIST pass (REPS=2: second pass is captured)
REPS=2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,launch__registers_per_thread,launch__grid_size \
  --launch-skip 60 -c 70 --csv --log-file gpurun_out/ncu_pass.csv python scripts/full_pass.py > gpurun_out/ncu_pass.log 2>&1
tail -2 gpurun_out/ncu_pass.log
