#!/bin/bash
# round-2 baseline: exp probe, GPU parity suite timing, whole-step per-kernel DRAM traffic (ncu metrics pass)
set -x
python scripts/exp_band_probe.py > gpurun_out/exp_band_probe.log 2>&1
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 > gpurun_out/r2_base_pytest.log
python scripts/full_pass.py > /dev/null 2>&1 || exit 1
REPS=2 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum \
  --clock-control none --launch-skip 66 -c 80 --csv --log-file gpurun_out/r2_base_step_traffic.csv python scripts/full_pass.py > gpurun_out/r2_base_ncu.log 2>&1
tail -2 gpurun_out/r2_base_ncu.log
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
