# one-off tuning sweep: flood geometries at the bench batch
run() {
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bt.json 2> gpurun_out/bt.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bt.json').read())
pk=d["roofline"]["per_kernel_ms_per_step"]
print("$1", round(d["value"]), round(d["ms_per_step"],3), "flood", pk.get("k_ws_flood_u8"))
PY
}
for v in 0 1 2 4; do TISEG_FLOOD_VARIANT=$v run "flood_variant=$v"; done
