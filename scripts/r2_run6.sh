#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "dist" 2>&1 | grep -B5 -A25 "def _diff\|AssertionError" | head -80 > gpurun_out/r2_pytest.log
