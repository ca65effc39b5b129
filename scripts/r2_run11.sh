#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "uint16 or pair_metrics" 2>&1 | tail -4 > gpurun_out/r2_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err
echo "rc=$?" >> gpurun_out/r2_pytest.log
