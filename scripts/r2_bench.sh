#!/bin/bash
timeout 280 python bench.py --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_tmp.json 2> gpurun_out/r2_bench_tmp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_tmp.json').read())
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"], d["roofline"]["pipeline_frac"])
PY
tail -3 gpurun_out/r2_bench_tmp.err
