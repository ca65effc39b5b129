#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -x -q -k host_feed 2>&1 | tail -4 > gpurun_out/r2_pytest.log
for wl in dist_monuseg_1000 unet_cpm17_256 conic_sweep_256 cdnet_consep_1000 hover_consep_1000; do
  TISEG_BENCH_ALLOW_UNKNOWN=1 timeout 400 python bench.py --workload $wl --steps 8 --warmup 3 --distinct 4 > gpurun_out/r2_wl_$wl.json 2> gpurun_out/r2_wl_$wl.err
  echo "$wl rc=$?" >> gpurun_out/r2_pytest.log
done
