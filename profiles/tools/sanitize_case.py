"""Small cases of the kernels that live on shared / global atomics, for compute-sanitizer (memcheck / racecheck):
fill (k_ptr, k_edges hash + pair table, k_bor_persistent), flats (sparse fast tier + slow tier), accumulation
(packed-word shared atomics), labelling (shared union-find), segmented reductions (pit catalog).
usage: compute-sanitizer --tool <memcheck|racecheck> python profiles/tools/sanitize_case.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import pydrodem_b200 as hc
from pydrodem_b200 import synth
from pydrodem_b200.tools import depressions as dep

ND = -9999.0
dev = torch.device("cuda:0")
for (ny, nx, seed, pitched) in ((200, 300, 3, False), (256, 320, 4, True)):
    z = synth.make_dem_numpy(ny, nx, seed, n_bowls=6)
    z[:, 0] = ND
    zt = hc.empty_raster(ny, nx, torch.float32, dev) if pitched else torch.empty((ny, nx), dtype=torch.float32, device=dev)
    zt.copy_(torch.from_numpy(z))
    for seed_mode, table in ((1, 1), (0, 0)):
        f, d, s, a = hc.fill_d8_accum(zt, nodata=ND, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table)
    torch.cuda.synchronize()
    diff = (f - zt).cpu().numpy() if not pitched else (f.contiguous() - zt.contiguous()).cpu().numpy()
    lab, n = hc.label((torch.from_numpy(diff) > 0).to(torch.uint8).cuda(), 8)
    print(ny, nx, "pitched" if pitched else "dense", "raised cells", int((diff > 0).sum()), "clumps", n)
# banded chain (seam graph, band accumulation) and a noise DEM (general path of k_edges, dense flats)
zn = synth.random_dem(150, 170, 5, "noise", nodata_holes=2, levels=6)
hc.fill_d8_accum(torch.from_numpy(zn).cuda(), nodata=ND, dx=1.0, dy=1.0)
from pydrodem_b200 import bands
bands.fill_d8_accum_banded(synth.make_dem_numpy(240, 200, 7, n_bowls=5), 3, dx=30.0, dy=30.0)
# pit catalog / selection / apply (segment.cu)
z = synth.make_dem_numpy(180, 200, 9, n_bowls=8)
z[0, :] = ND
fill = hc.fill(z, nodata=ND, seed_mode=hc.SEED_WBT)
carve = hc.breach_depressions(z, nodata=ND, max_length=10, fill_pits=True)
dfill, dcarve = np.round(fill - z, 3), np.round(carve - z, 3)
fid = hc.label((dfill > 0).astype(np.uint8), 8)[0]
cid = hc.label((dcarve < 0).astype(np.uint8), 8)[0]
cfid = hc.label((dcarve > 0).astype(np.uint8), 8)[0]
DF, fc, cco, cc, cfc = dep.build_pit_catalog(z.ravel(), ND, carve.ravel(), dcarve.ravel(), dfill.ravel(), fid.ravel(), cid.ravel(), cfid.ravel())
DF = dep.select_solution(DF, fc, cc, cco, fill.ravel(), carve.ravel(), fid.ravel(), cid.ravel(), 60, 25, 6,
                         water_mask_array=np.zeros(z.shape, np.uint8))
dep.apply_solution(DF, fc, cc, cfc, z.ravel().copy(), fill.ravel(), carve.ravel(), excludedpits=[])
torch.cuda.synchronize()
print("sanitize case done:", len(DF), "pits")
