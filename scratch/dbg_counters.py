import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pydrodem_b200 as hc
from pydrodem_b200 import _lib, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10801
dev = torch.device("cuda:0")
z = torch.empty((n, n), dtype=torch.float32, device=dev)
for r0 in range(0, n, 1024):
    nr = min(1024, n - r0)
    z[r0:r0 + nr] = synth.make_dem(n, n, 1002, row0=r0, nrows=nr, device=dev)
kw = dict(dx=30.0, dy=30.0, seed_mode=1, table=1, flats=True, want_slope=False)
hc.fill_d8_accum(z, **kw); torch.cuda.synchronize()
_lib.debug_read(0)
f, d, s, a = hc.fill_d8_accum(z, **kw); torch.cuda.synchronize()
c = _lib.debug_read(0)
print("fast sweeps", c[0], "fast tiles with work", c[1], "pending tiles", c[2], "slow sweeps", c[3], "slow visits", c[4], "dense visits", c[5], "other", c[6:16])
print("noflow cells after fill d8:", int((d == 0).sum()))
