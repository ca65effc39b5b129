python -m pytest tests -m gpu -x -q 2>&1 | tail -1
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bt.json 2> gpurun_out/bt.err
python -c "
import json
d=json.loads(open('gpurun_out/bt.json').read())
print(d['value'], d['ms_per_step'], d['roofline']['per_kernel_ms_per_step'].get('k_ws_flood_u8'))"
python scripts/kernel_times.py 3 2>&1 | head -8
