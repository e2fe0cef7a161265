import sys, tempfile, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_pipeline import unit_dem, ND
from pydrodem_b200 import raster_io
from pydrodem_b200.models.depression_handling import DepHandleCore
d = tempfile.mkdtemp()
z = unit_dem()
raster_io.write_raster(d + "/basin_7.tif", z, None, nodata=ND)
h = DepHandleCore(basename="basin_7", radius=12, sfill=0.5, fill_vol_thresh=60, carve_vol_thresh=25, max_carve_depth=6)
import pydrodem_b200.tools.depressions as dep
orig = dep.combined_solution
def spy(*a, **k):
    r = orig(*a, **k); print("combined pit", a[9], "->", r[0], "rows", len(r[1])); return r
dep.combined_solution = spy
ff, mod, DF, sm = h.run_depression_core(d + "/basin_7.tif", d)
print(DF.groupby("solution").count()["maxfill"])
print("changed cells", int((ff != z).sum()), "of", z.size)
