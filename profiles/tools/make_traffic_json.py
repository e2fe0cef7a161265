"""profiles/ncu_traffic.json (bench.py's `roofline.traffic`) from an ncu launch list (see profiles/summarize_launches.py)."""
import csv
import json
import sys
from collections import OrderedDict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
launch = OrderedDict()
for r in rows[1:]:
    d = launch.setdefault(r[ix["ID"]], {"name": r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]})
    v = float(r[ix["Metric Value"]].replace(",", ""))
    v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ix["Metric Unit"]], 1.0)
    d[r[ix["Metric Name"]]] = v
agg = OrderedDict()
for d in launch.values():
    a = agg.setdefault(d["name"], {"launches": 0, "bytes": 0.0})
    a["launches"] += 1
    a["bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
out = {"source": f"{path} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, C2 workload on the pitched raster, 2 chain runs)",
       "kernels": {k: {"launches": v["launches"], "dram_bytes_per_launch": v["bytes"] / v["launches"]} for k, v in agg.items()}}
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(sum(v["bytes"] for v in agg.values()) / 2 / 1e9, "GB of DRAM traffic per chain")
