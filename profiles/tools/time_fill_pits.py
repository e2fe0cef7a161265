"""Where dep.fill_pits spends its time on a 3601^2 tile (host I/O vs device work).  usage: time_fill_pits.py"""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import pydrodem_b200 as hc
from pydrodem_b200 import raster_io, synth, core
from pydrodem_b200.wbt import WhiteboxToolsGPU
from pydrodem_b200.tools import depressions as dep

z = synth.make_dem_numpy(3601, 3601, 1001)
d = tempfile.mkdtemp()
p = os.path.join(d, "inDEM.tif")
raster_io.write_raster(p, z, None, nodata=-9999.0)
wbt = WhiteboxToolsGPU()
def T(label, fn, n=3):
    fn(); torch.cuda.synchronize()
    t = time.time()
    for _ in range(n):
        r = fn()
    torch.cuda.synchronize()
    print(f"{label:40s} {(time.time() - t) / n * 1e3:8.1f} ms")
    return r
T("read_raster", lambda: raster_io.read_raster(p))
a, meta = raster_io.read_raster(p)
T("core.fill (numpy in/out)", lambda: core.fill(a, nodata=-9999.0, seed_mode=0))
f = core.fill(a, nodata=-9999.0, seed_mode=0)
T("write_raster", lambda: raster_io.write_raster(os.path.join(d, "o.tif"), f, meta, nodata=-9999.0))
T("wbt.fill_depressions_wang_and_liu", lambda: wbt.fill_depressions_wang_and_liu(p, os.path.join(d, "w.tif"), fix_flats=False))
T("core.diff_round (numpy in/out)", lambda: core.diff_round(f, a, 3))
T("np.round(f - a, 3)", lambda: np.round(f - a, 3))
T("dep.fill_pits", lambda: dep.fill_pits(p, wbt))
T("dep.fill_pits(save_tif=False)", lambda: dep.fill_pits(p, wbt, save_tif=False))
za = torch.from_numpy(a)
T("H2D pageable 52 MB", lambda: za.cuda())
zp = za.pin_memory()
T("H2D pinned 52 MB", lambda: zp.cuda(non_blocking=True))
zd = za.cuda()
T("D2H to new numpy", lambda: zd.cpu().numpy())
T("device fill only", lambda: hc.fill(zd, nodata=-9999.0, seed_mode=0))
