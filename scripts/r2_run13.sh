#!/bin/bash
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_pytest.log
timeout 300 python bench.py --workload conic_sweep_256 --steps 8 --warmup 3 --distinct 4 > gpurun_out/r2_wl_conic_sweep_256.json 2> gpurun_out/r2_wl_conic.err
echo "conic rc=$?" >> gpurun_out/r2_pytest.log
