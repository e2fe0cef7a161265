"""`WhiteboxToolsGPU` — a duck type of the `wbt` object pydrodem passes around.

pydrodem builds one `whitebox_tools.WhiteboxTools()` per process (models/depression_handling.py:237-242,
models/find_sinks.py:47-52) and hands it positionally to `dep.fill_pits(inDEM, wbt, ...)`,
`dep.create_clump_array(..., wbt, ...)` etc.; every method takes GeoTIFF *paths*.  This class exposes
the method subset on the depression/D8 path with the exact keyword names the reference uses
 and runs them on libhydrocore through the C ABI.  Like the real tool it returns 0 on
success; unlike it, it raises on failure (the reference ignores return codes and only notices a
missing output file).

Deliberate differences, all stated where they apply:
  * `flat_increment > 0` (Priority-Flood+epsilon) is served by `hc_fill_eps_dev` (float64 out, as WhiteboxTools
    writes): a stated-tolerance mode, WhiteboxTools' own heap order on ties being unknown.
  * outputs are float32 GeoTIFFs (WhiteboxTools writes float64 and pydrodem narrows them right
    away with `pyct.tifdowncast`, tools/depressions.py:1306); D8 pointers are int16 with -32768 on
    nodata cells as WhiteboxTools writes them.
  * breach_depressions runs the order-independent flow-path breach on the GPU (hc_breach_flowpath_dev; the host
    priority-order formulation of csrc/breach.cu remains as core.breach_depressions(method="flood"));
    breach_depressions_least_cost is not built.
"""
import numpy as np

from . import core, raster_io
from ._lib import SEED_WBT, TABLE_WBT


class WhiteboxToolsGPU:
    def __init__(self):
        self.verbose = False
        self.exe_path = None

    # -- plumbing the reference calls at construction ------------------------------------------------
    def set_verbose_mode(self, val=True):
        self.verbose = bool(val)
        return 0

    def set_whitebox_dir(self, path):
        self.exe_path = path
        return 0

    def _say(self, *a):
        if self.verbose:
            print(*a)

    @staticmethod
    def _load(path):
        arr, meta = raster_io.read_raster(path)
        nodata = meta.nodata if meta.nodata is not None else -9999.0
        return arr, meta, float(nodata)

    @staticmethod
    def _eps(fix_flats, flat_increment):
        """the gradient WhiteboxTools imposes on filled areas: `flat_increment` when given (> 0), nothing when fix_flats is
        off.  fix_flats=True WITHOUT an increment lets WhiteboxTools derive one from the raster's resolution and
        elevation range (its own heuristic, not reproducible without the binary): refused."""
        if flat_increment is not None and flat_increment > 0:
            return float(flat_increment)
        if fix_flats:
            raise NotImplementedError("fix_flats=True needs an explicit flat_increment here (WhiteboxTools' automatic "
                                      "increment is its own heuristic); pydrodem passes fix_flats only together with "
                                      "flat_increment > 0 (tools/depressions.py:1271-1273)")
        return 0.0

    # -- fills (tools/depressions.py:1277-1286, :1006, :1385) ---------------------------------------
    def fill_depressions_wang_and_liu(self, dem, output, fix_flats=True, flat_increment=None, callback=None):
        eps = self._eps(fix_flats, flat_increment)
        z, meta, nodata = self._load(dem)
        if eps > 0.0:
            # Priority-Flood + epsilon, float64 like WhiteboxTools' own output (tolerance class, see core.fill_eps)
            out = core.fill_eps(np.asarray(z, np.float32), eps, nodata=nodata, seed_mode=SEED_WBT)
        else:
            out = core.fill(np.asarray(z, np.float32), nodata=nodata, seed_mode=SEED_WBT)
        raster_io.write_raster(output, out, meta, nodata=nodata)
        self._say("fill_depressions_wang_and_liu ->", output)
        return 0

    def fill_depressions(self, dem, output, fix_flats=True, flat_increment=None, max_depth=None, callback=None):
        if max_depth is not None:
            raise NotImplementedError("fill_depressions(max_depth=...) is not built; pydrodem passes None "
                                      "(tools/depressions.py:1278)")
        # with fix_flats=False and no depth cap this is the same unique surface as the Wang & Liu tool
        return self.fill_depressions_wang_and_liu(dem, output, fix_flats=fix_flats, flat_increment=flat_increment)

    # -- single-cell pits (models/depression_handling.py:259-261, models/find_sinks.py:118) ----------
    def fill_single_cell_pits(self, dem, output, callback=None):
        z, meta, nodata = self._load(dem)
        raster_io.write_raster(output, core.fill_single_cell_pits(np.asarray(z, np.float32), nodata), meta, nodata=nodata)
        return 0

    def breach_single_cell_pits(self, dem, output, callback=None):
        z, meta, nodata = self._load(dem)
        raster_io.write_raster(output, core.breach_single_cell_pits(np.asarray(z, np.float32), nodata), meta,
                               nodata=nodata)
        return 0

    # -- D8 (run_aws.py:539, :544) ------------------------------------------------------------------
    def d8_pointer(self, dem, output, esri_pntr=False, callback=None):
        if esri_pntr:
            raise NotImplementedError("esri_pntr=True is not used by pydrodem")
        z, meta, nodata = self._load(dem)
        dx, dy = meta.pixel_size
        d = core.d8(np.asarray(z, np.float32), nodata=nodata, dx=dx, dy=dy, table=TABLE_WBT, want_slope=False)
        p = d.astype(np.int16)
        p[d == 255] = -32768
        raster_io.write_raster(output, p, meta, nodata=-32768)
        return 0

    def d8_flow_accumulation(self, i, output, out_type="cells", log=False, clip=False, pntr=False, esri_pntr=False,
                             callback=None):
        if not pntr:
            raise NotImplementedError("pydrodem calls d8_flow_accumulation with pntr=True (run_aws.py:544)")
        if out_type != "cells" or log or clip or esri_pntr:
            raise NotImplementedError("only out_type='cells' without log/clip is on the reference path")
        p, meta, nodata = self._load(i)
        d = np.where(p == nodata, 255, p).astype(np.uint8)
        a = core.accum(d, table=TABLE_WBT).astype(np.float32)
        a[d == 255] = -32768.0
        raster_io.write_raster(output, a, meta, nodata=-32768.0)
        return 0

    # -- breach (tools/depressions.py:1376-1380, :1012-1014) ------------------------------------------
    def breach_depressions(self, dem, output, max_depth=None, max_length=None, flat_increment=None, fill_pits=False,
                           callback=None):
        """Lindsay (2016) constrained breaching + fill of the rest; see core.breach_depressions for what runs where"""
        z, meta, nodata = self._load(dem)
        out = core.breach_depressions(np.asarray(z, np.float32), nodata=nodata, max_length=max_length,
                                      max_depth=max_depth, fill_pits=fill_pits, flat_increment=flat_increment)
        raster_io.write_raster(output, out, meta, nodata=nodata)
        self._say("breach_depressions ->", output)
        return 0

    # -- gap interpolation (tools/water.py:1396, :1774; models/burn_rivers.py:1283, :1582; dem_noise_removal.py:739) --
    def fill_missing_data(self, i, output, filter=11, weight=2.0, no_edges=False, callback=None):
        """IDW infill of the cells pydrodem blanks before re-interpolating them (hc_fill_missing_data_dev): a nodata
        cell takes the 1 / d^weight weighted mean of the valid cells bordering nodata within filter / 2 cells.  Restated
        from the tool's documentation (parity unpinned).  no_edges=True (skip nodata connected to the raster frame) is
        not what pydrodem passes and is not built."""
        if no_edges:
            raise NotImplementedError("fill_missing_data(no_edges=True) is not used by pydrodem")
        from .tools import water
        z, meta, nodata = self._load(i)
        out = water.fill_missing_data(np.asarray(z, np.float32), nodata, int(filter), float(weight))
        raster_io.write_raster(output, out, meta, nodata=nodata)
        self._say("fill_missing_data ->", output)
        return 0

    # -- least-cost breach (tools/depressions.py:1370-1374, :996-1000: carve_method 'LC') ----------------
    def breach_depressions_least_cost(self, dem, output, dist, max_cost=None, min_dist=True, flat_increment=None,
                                      fill=True, callback=None):
        """Lindsay & Dhun (2015) least-cost breaching, one CTA per pit on the GPU; see core.breach_depressions_least_cost
        (order-independent restatement, parity unpinned).  dist is the search radius in cells (at most 65 here: the
        window lives in shared memory; the reference passes radius = 50, config_files/config_local.ini:162)."""
        z, meta, nodata = self._load(dem)
        out = core.breach_depressions_least_cost(np.asarray(z, np.float32), int(dist), nodata=nodata, max_cost=max_cost,
                                                 min_dist=min_dist, flat_increment=flat_increment, fill=fill)
        raster_io.write_raster(output, out, meta, nodata=nodata)
        self._say("breach_depressions_least_cost ->", output)
        return 0
