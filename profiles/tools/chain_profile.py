"""Per-kernel CUDA-event times of one fill+D8+flats+accum chain (the library's own hc_profile_* records).
usage: python profiles/tools/chain_profile.py [size=10801] [steps=5] [workload=c2|c4] ; prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pydrodem_b200 as hc
from pydrodem_b200 import _lib, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10801
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
wl = sys.argv[3] if len(sys.argv) > 3 else "c2"
dev = torch.device("cuda:0")
if wl == "c4":
    z0, _ = synth.make_c4(n, n, 1004, device=dev)
    z = hc.empty_raster(n, n, torch.float32, dev)
    z.copy_(z0)
    del z0
else:
    pitched = os.environ.get("HC_PITCHED", "1") == "1"   # pitched device raster: the TMA / vector path
    z = hc.empty_raster(n, n, torch.float32, dev) if pitched else torch.empty((n, n), dtype=torch.float32, device=dev)
    for r0 in range(0, n, 1024):
        nr = min(1024, n - r0)
        z[r0:r0 + nr] = synth.make_dem(n, n, 1002, row0=r0, nrows=nr, device=dev)
kw = dict(dx=30.0, dy=30.0, seed_mode=1, table=1, flats=True, want_slope=False)
for _ in range(3):
    hc.fill_d8_accum(z, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    hc.fill_d8_accum(z, **kw)
e1.record()
torch.cuda.synchronize()
ms_plain = e0.elapsed_time(e1) / steps
_lib.profile_enable(True, 0)
for _ in range(steps):
    hc.fill_d8_accum(z, **kw)
torch.cuda.synchronize()
prof = _lib.profile_read(0)
_lib.profile_enable(False, 0)
_lib.debug_read(0)
hc.fill_d8_accum(z, **kw)
torch.cuda.synchronize()
dbg = [int(x) for x in _lib.debug_read(0)]
dbg_names = {0: "fast_sweeps", 1: "fast_tiles_with_work", 2: "pending_tiles", 3: "slow_sweeps", 4: "slow_visits",
             5: "lake_visits", 6: "light_visits", 8: "edges_general_path_tiles", 9: "light_cycles", 10: "lake_cycles",
             11: "lake_iterations", 12: "light_cycles_load", 13: "light_cycles_fold", 14: "light_cycles_sweeps", 15: "light_cells"}
rows = sorted(((k, c / steps, t / steps) for k, (c, t, m) in prof.items()), key=lambda r: -r[2])
print(json.dumps({"size": n, "workload": wl, "ms_per_step": ms_plain, "kernel_ms_sum": sum(r[2] for r in rows),
                  "fill_stats": hc.core.fill_stats(0), "debug": {v: dbg[k] for k, v in dbg_names.items()},
                  "kernels": {k: {"launches": c, "ms": round(t, 4)} for k, c, t in rows}}))
