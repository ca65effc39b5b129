#!/usr/bin/env python
"""ncu metrics pass over scripts/step_pass.py (two identical steps; the SECOND is kept) -> per-kernel DRAM bytes, L2
bytes, warp instructions and time of ONE whole step: profiles/r2_step_traffic_<workload>.json.  bench.py reads the file
named by its workload to report `roofline.step_dram_bytes` / `traffic_ratio` and each kernel's fraction on
min(algorithmic, measured) bytes.

    python scripts/digest_step_traffic.py gpurun_out/r2_step_dist.csv profiles/r2_step_traffic_dist_monuseg_1000.json 64 22 1000 dist_monuseg_1000 [all]
"""
import collections, csv, json, re, sys

src, dst, batch, pipe_bpp = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
H = W = int(sys.argv[5]) if len(sys.argv) > 5 else 1000
workload = sys.argv[6] if len(sys.argv) > 6 else "dist_monuseg_1000"
lines = [l for l in open(src) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3,
        "msecond": 1e3, "inst": 1, "": 1}
per = collections.OrderedDict()
for r in rows:
    name = r["Kernel Name"]
    m = re.match(r"(?:void )?(?:tiseg::)?([A-Za-z0-9_]+)", name)
    short = m.group(1) if m else name
    if "at::" in name or "at_cuda" in name:
        short = "torch:" + short
    k = per.setdefault((r["ID"], short), {})
    v = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1)
    k[r["Metric Name"]] = v
if not (len(sys.argv) > 7 and sys.argv[7] == "all"):       # two steps were captured: keep the second (the capture
    items = list(per.items())                                # starts with a launch sequence that repeats right after itself)
    names = [k[1] for k, _ in items]
    L = next(l for l in range(8, len(names) // 2 + 1) if names[l:2 * l] == names[:l])
    per = collections.OrderedDict(items[L:2 * L])
agg = collections.OrderedDict()
for (_, short), m in per.items():
    a = agg.setdefault(short, {"launches": 0, "us": 0.0, "dram_bytes": 0.0, "l2_bytes": 0.0, "warp_inst": 0.0})
    a["launches"] += 1
    a["us"] += m.get("gpu__time_duration.sum", 0.0)
    a["dram_bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
    a["l2_bytes"] += m.get("lts__t_bytes.sum", 0.0)
    a["warp_inst"] += m.get("smsp__inst_executed.sum", 0.0)
tot = sum(a["dram_bytes"] for a in agg.values())
tot_us = sum(a["us"] for a in agg.values())
alg = pipe_bpp * H * W * batch
out = {"workload": workload, "batch": batch, "tile": [H, W],
       "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,"
              "smsp__inst_executed.sum --clock-control none over the second of two steps of scripts/step_pass.py (one whole "
              "step; per-launch times under ncu are serialised and cold-cache; cudaMemset is not a kernel and is not listed)",
       "launches": sum(a["launches"] for a in agg.values()), "step_dram_bytes": tot, "step_us_serialised": tot_us,
       "algorithmic_bytes": alg, "traffic_ratio": tot / alg,
       "kernels": {k: {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in a.items()}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["dram_bytes"])}}
json.dump(out, open(dst, "w"), indent=1)
print("launches %d, step DRAM %.1f MB = %.2fx algorithmic (%.1f MB), serialised %.0f us" % (out["launches"], tot / 1e6, tot / alg, alg / 1e6, tot_us))
for k, a in list(out["kernels"].items())[:24]:
    print("%-28s n=%2d %8.1f us %8.1f MB dram %8.1f MB l2 %7.1f M inst" % (k, a["launches"], a["us"], a["dram_bytes"] / 1e6, a["l2_bytes"] / 1e6, a["warp_inst"] / 1e6))
