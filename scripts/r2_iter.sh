#!/bin/bash
# one development iteration on the GPU box: parity tests of the touched stages, per-kernel step times, a short bench
set -o pipefail
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest.log
timeout 200 python scripts/step_times.py > gpurun_out/r2_times_new.log 2>&1; echo "times rc=$?"; head -22 gpurun_out/r2_times_new.log
timeout 280 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_tmp.json').read())
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["pipeline_frac"], d.get("check"))
PY
tail -3 gpurun_out/r2_bench_tmp.err
