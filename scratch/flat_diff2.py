import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np, pydrodem_b200 as hc
from pydrodem_b200 import synth
np.set_printoptions(linewidth=250)
n = 10801
dev = torch.device("cuda:0")
z = torch.empty((n, n), dtype=torch.float32, device=dev)
for r0 in range(0, n, 1024):
    nr = min(1024, n - r0)
    z[r0:r0 + nr] = synth.make_dem(n, n, 1002, row0=r0, nrows=nr, device=dev)
f = hc.fill(z, seed_mode=1)
dA = hc.d8(f, dx=30.0, dy=30.0, table=1, want_slope=False)
dnew = hc.resolve_flats(f, dA, table=1)
os.environ["HC_FLAT_START_ALL"] = "1"
dref = hc.resolve_flats(f, dA, table=1)
diff = torch.nonzero(dnew != dref)
print("diff cells", diff.shape[0])
r, c = 4859, 64
sl = (slice(r - 6, r + 7), slice(40, 90))
flat = (f[sl] == f[r, c])
print("flat mask (1 = same z):"); print(flat.to(torch.uint8).cpu().numpy())
print("dirA:"); print(dA[sl].cpu().numpy())
print("new:"); print(dnew[sl].cpu().numpy())
print("ref:"); print(dref[sl].cpu().numpy())
