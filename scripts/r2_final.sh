#!/bin/bash
# round-2 closing pass on one B200: GPU tests, smoke, the default bench line + the reference arm, the other workloads
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference rc=$?"
for wl in unet_cpm17_256 conic_sweep_256 cdnet_consep_1000 hover_consep_1000; do
  timeout 500 python bench.py --workload $wl --steps 8 --warmup 3 > gpurun_out/r2_wl_$wl.json 2> gpurun_out/r2_wl_$wl.err; echo "$wl rc=$?"
done
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/r2_gpu.txt
