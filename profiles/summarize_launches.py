#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv` launch list into a per-kernel table (markdown).

    python profiles/summarize_launches.py gpurun_out/launches.csv [steps] > profiles/rNN_launches.md

Times under ncu are cold-cache and serialised: a kernel's SHARE of the step is what has to agree with
bench.py's CUDA-event numbers, not the absolute.  `DRAM GB/s` = (read+write bytes) / duration of that
launch; `% of peak` uses MEASURED_PEAKS.json (6538.3 GB/s) when present.
"""
import csv
import json
import os
import sys
from collections import OrderedDict

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
peak = 6538.3
pk = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launch = OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]].split("(")[0], "grid": r[ix["Grid Size"]],
                                        "block": r[ix["Block Size"]]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    m = r[ix["Metric Name"]]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)  # -> us
    else:
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    d[m] = v
agg = OrderedDict()
for d in launch.values():
    a = agg.setdefault(d["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "grid": d["grid"], "block": d["block"]})
    a["n"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0)
    a["rd"] += d.get("dram__bytes_read.sum", 0.0)
    a["wr"] += d.get("dram__bytes_write.sum", 0.0)
tot = sum(a["us"] for a in agg.values())
print(f"# ncu launch list: {os.path.basename(path)} ({len(launch)} launches, {steps} step(s) captured, "
      f"total {tot/1e3:.2f} ms under ncu; HBM peak {peak:.0f} GB/s)\n")
print("| kernel | launches/step | us/step | share | DRAM read MB/step | DRAM write MB/step | DRAM GB/s | % of peak | grid | block |")
print("|---|---|---|---|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    gbs = (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0
    print(f"| {name} | {a['n']/steps:g} | {a['us']/steps:.1f} | {100*a['us']/tot:.1f}% | {a['rd']/steps/1e6:.1f} | "
          f"{a['wr']/steps/1e6:.1f} | {gbs:.0f} | {100*gbs/peak:.1f}% | {a['grid']} | {a['block']} |")
