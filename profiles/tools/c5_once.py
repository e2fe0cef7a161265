"""One DEM-NR smoothing core (stencils.smooth_core) on an n x n synthetic DEM, for ncu.  usage: c5_once.py [n=8000] [reps=2]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from pydrodem_b200 import stencils, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
z = torch.empty((n, n), dtype=torch.float32, device=dev)
for r0 in range(0, n, 1024):
    nr = min(1024, n - r0)
    z[r0:r0 + nr] = synth.make_dem(n, n, 1005, row0=r0, nrows=nr, device=dev)
for _ in range(reps):
    zf, filt, slope = stencils.smooth_core(z, 30.0, 30.0)
torch.cuda.synchronize()
print("ok", float(zf.mean()), int(filt.to(torch.int32).sum()), int(slope.to(torch.int32).sum()))
