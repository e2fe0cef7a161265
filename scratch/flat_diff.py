import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pydrodem_b200 as hc
from pydrodem_b200 import synth
n = 10801
dev = torch.device("cuda:0")
z = torch.empty((n, n), dtype=torch.float32, device=dev)
for r0 in range(0, n, 1024):
    nr = min(1024, n - r0)
    z[r0:r0 + nr] = synth.make_dem(n, n, 1002, row0=r0, nrows=nr, device=dev)
kw = dict(dx=30.0, dy=30.0, seed_mode=1, table=1, flats=True, want_slope=False)
f, d, s, a = hc.fill_d8_accum(z, **kw); torch.cuda.synchronize()
d1 = d.clone()
os.environ["HC_FLAT_START_ALL"] = "1"
f, d, s, a = hc.fill_d8_accum(z, **kw); torch.cuda.synchronize()
diff = torch.nonzero(d1 != d)
print("cells differing:", diff.shape[0], "noflow new:", int((d1 == 0).sum()), "noflow all:", int((d == 0).sum()))
for rc in diff[:6].tolist():
    r, c = rc
    print(rc, "tile", r // 64, c // 64, "in-tile", r % 64, c % 64, "new", int(d1[r, c]), "ref", int(d[r, c]), "z", float(f[r, c]))
    print(f[r-2:r+3, c-2:c+3].cpu().numpy())
    print(d[r-2:r+3, c-2:c+3].cpu().numpy())
