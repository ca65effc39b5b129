#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest.log
python scripts/step_times.py > gpurun_out/r2_times_new.log 2>&1
