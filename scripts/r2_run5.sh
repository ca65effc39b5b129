#!/bin/bash
for vl in 2 3 4; do
  timeout 200 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --value-lanes $vl > gpurun_out/r2_vl$vl.json 2> gpurun_out/r2_vl$vl.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2_vl$vl.json').read())
print($vl, d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"])
PY
done
