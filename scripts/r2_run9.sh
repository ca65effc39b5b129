#!/bin/bash
timeout 500 python -m pytest tests -m gpu -q 2>&1 | grep -v "Warning\|warn\|^$\|_finish" | tail -40 > gpurun_out/r2_pytest.log
