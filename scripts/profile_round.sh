#!/bin/bash
# Round profile on the GPU box: tests, bench (plain), ncu launch list of the same bench command, full captures of the
# two heaviest kernels.  Outputs under gpurun_out/ (copied to profiles/ by scripts/digest_profiles.py here).
R=${1:-r1}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err || exit 1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_${R}.log 2>&1
python scripts/full_pass.py > /dev/null 2>&1 || exit 1
REPS=2 ncu --set full --clock-control none --import-source on -k regex:k_ws_flood_u8 --launch-skip 2 -c 1 -f \
    -o gpurun_out/${R}_flood python scripts/full_pass.py > gpurun_out/ncu_flood_${R}.log 2>&1
REPS=2 ncu --set full --clock-control none --import-source on -k regex:k_ccl_local --launch-skip 6 -c 1 -f \
    -o gpurun_out/${R}_ccl_local python scripts/full_pass.py > gpurun_out/ncu_ccl_${R}.log 2>&1
tail -2 gpurun_out/ncu_ccl_${R}.log
