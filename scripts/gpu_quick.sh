python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tmp.json').read())
print(d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"])
pk=d["roofline"]["per_kernel_ms_per_step"]
print({k:v for k,v in pk.items() if v>0.02})
PY
tail -3 gpurun_out/bench_tmp.err
