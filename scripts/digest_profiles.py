#!/usr/bin/env python
"""Turn the raw artefacts a `scripts/profile_round.sh rN` run left under gpurun_out/ into the tracked summaries under
profiles/: bench line, ncu launch list (per-kernel totals + share), and for every full capture a details text, the
raw metric CSV and a short JSON; the DRAM traffic of the bench's dominant kernel goes to
profiles/dominant_kernel_traffic.json (read back by bench.py's `roofline.traffic`)."""
import collections, csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(PR, exist_ok=True)

bench = json.loads(open(os.path.join(GO, "bench_%s.json" % R)).read())
shutil.copy(os.path.join(GO, "bench_%s.json" % R), os.path.join(PR, "%s_bench.json" % R))

# ---- launch list
lines = [l for l in open(os.path.join(GO, "launches_%s.csv" % R)) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v
    a = agg.setdefault(r["Kernel Name"], [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
with open(os.path.join(PR, "%s_launches_dist1000_b%d.csv" % (R, bench["config"]["tiles_per_step_per_gpu"])), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, python bench.py --steps 2 --warmup 1 (first 400 launches)\n")
    f.write("kernel,launches,total_us,share\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('"%s",%d,%.1f,%.4f\n' % (k, a[0], a[1], a[1] / tot))

# ---- full captures
def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(out.splitlines()))
    return dict(zip(rr[0], rr[-1])), out

for tag in ("flood", "ccl_local", "flatten", "plateau", "pair", "wsl_remove"):
    rep = os.path.join(GO, "%s_%s.ncu-rep" % (R, tag))
    if not os.path.exists(rep):
        continue
    d, out = raw(rep)
    if tag in ("flood", "ccl_local"):          # the raw metric dump only for the two heaviest kernels
        open(os.path.join(PR, "%s_%s_full_raw.csv" % (R, tag)), "w").write(out)
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    open(os.path.join(PR, "%s_%s_details.txt" % (R, tag)), "w").write(det)
    keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]
    summ = {k: d.get(k) for k in keys}
    json.dump(summ, open(os.path.join(PR, "%s_%s_summary.json" % (R, tag)), "w"), indent=1)
    print(tag, summ)
    dom = bench["roofline"]["kernel"]
    if (tag == "ccl_local" and dom.startswith("(k_ccl_local<Img, 2")) or (tag == "flood" and dom.startswith("k_ws_flood_u8")):
        unit = 1e6            # ncu prints Mbyte for these captures
        traffic = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])) * unit
        json.dump({"kernel": dom, "batch": bench["config"]["tiles_per_step_per_gpu"], "dram_bytes_per_launch": traffic,
                   "source": "profiles/%s_%s_full_raw.csv (ncu --set full, one launch)" % (R, tag)},
                  open(os.path.join(PR, "dominant_kernel_traffic.json"), "w"), indent=1)
print("share of top kernels (launch list):", [(k[:40], round(a[1] / tot, 3)) for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:6]])
