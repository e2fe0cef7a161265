"""N fill+D8+flats+accum chains on the pitched C2 raster (for ncu).  usage: one_chain.py [size=10801] [reps=2]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import pydrodem_b200 as hc
from pydrodem_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10801
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
z = hc.empty_raster(n, n, torch.float32, dev)
for r0 in range(0, n, 1024):
    nr = min(1024, n - r0)
    z[r0:r0 + nr] = synth.make_dem(n, n, 1002, row0=r0, nrows=nr, device=dev)
kw = dict(dx=30.0, dy=30.0, seed_mode=1, table=1, flats=True, want_slope=False)
for _ in range(reps):
    hc.fill_d8_accum(z, **kw)
    torch.cuda.synchronize()
print("done")
