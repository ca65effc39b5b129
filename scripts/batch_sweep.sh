python -m pytest tests -m gpu -x -q -k "sem_counts or semantic or config" 2>&1 | tail -2
for b in 32 64 96; do for vl in 1 2 3; do
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch $b --value-lanes $vl > gpurun_out/bt.json 2> gpurun_out/bt.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bt.json').read())
print("batch $b lanes $vl", round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],3), d["roofline"]["per_kernel_ms_per_step"].get("k_sem_counts"), d["roofline"]["per_kernel_ms_per_step"].get("k_ws_flood_u8"))
PY
done; done
