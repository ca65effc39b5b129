# one-off tuning sweep: plateau band height and flood geometry at the bench batch
run() {
python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bt.json 2> gpurun_out/bt.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bt.json').read())
pk=d["roofline"]["per_kernel_ms_per_step"]
print("$1", round(d["value"]), round(d["ms_per_step"],3), "plateau", pk.get("k_plateau_bits"), "flood", pk.get("k_ws_flood_u8"))
PY
}
for r in 8 16 32 64; do TISEG_PB_ROWS=$r run "pb_rows=$r"; done
for v in 1 2 3; do TISEG_FLOOD_VARIANT=$v run "flood_variant=$v"; done
