"""dev probe: the per-unit depression pipeline on a C1-size (3601^2) synthetic DEM, stage times"""
import json, sys, tempfile, time
import numpy as np
sys.path.insert(0, ".")
from pydrodem_b200 import raster_io, synth
from pydrodem_b200.models.depression_handling import DepHandleCore
import pydrodem_b200.tools.depressions as dep
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3601
z = synth.make_dem_numpy(n, n, 1001)
z[0, :] = -9999.0; z[:, 0] = -9999.0
d, dw = tempfile.mkdtemp(), tempfile.mkdtemp()
raster_io.write_raster(d + "/unit.tif", z, None, nodata=-9999.0)
raster_io.write_raster(dw + "/unit.tif", z[:1200, :1200], None, nodata=-9999.0)
times = {}
def timed(mod, name):
    f = getattr(mod, name)
    def w(*a, **k):
        t = time.time(); r = f(*a, **k); times[name] = times.get(name, 0.0) + time.time() - t; return r
    setattr(mod, name, w)
for nm in ("fill_pits", "breach_pits", "create_clump_array", "fill_shallow_pits", "build_pit_catalog", "select_solution",
           "apply_solution", "combined_solution"):
    timed(dep, nm)
h = DepHandleCore(basename="unit", radius=50, sfill=2.0)
h.run_depression_core(dw + "/unit.tif", dw)     # warm-up of the library (context, arena) on a crop
times.clear()
t0 = time.time()
ff, mod, DF, sm = h.run_depression_core(d + "/unit.tif", d)
tot = time.time() - t0
print(json.dumps({"n": n, "total_s": round(tot, 2), "pits": int(len(DF)), "stages_s": {k: round(v, 2) for k, v in times.items()},
                  "codes": {int(k): int(v) for k, v in DF.groupby("solution").count()["maxfill"].items()}}))
