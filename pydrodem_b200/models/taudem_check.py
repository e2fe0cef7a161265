"""`models/taudem_check.py` of the reference (:32-247) on the GPU path: after the region-wide PitRemove / D8FlowDir /
AreaD8 run, every basin is checked for flow-direction nodata on valid flattened-water cells and, if any is found (and
the basin is not endorheic), its water is raised by 10 cm for a second pass.

Same class, constructor, method names, file names and `-c config -l log` entry point.  The reference cuts the four
per-basin rasters out of region VRTs along the basin polygon (GDAL cutlines + geopandas: out of scope);
here a basin directory must already hold them on one grid -- `taudem_DEM.tif`, `taudem_d8_flowdir.tif` (int16, nodata
99 after the cut, :108-110), `taudem_flatten_mask.tif`, `taudem_EDM.tif` -- or a `window.json` {"row0", "row1",
"col0", "col1"} from which the first two are cut out of the same-grid `fel` / `p` mosaics, with
`Water_flatten_mask.tif` (and optionally `EDM.tif`) beside it.  The problem mask and the +0.10 m fix are
`hc_taudem_check_dev` / `hc_taudem_fix_dev`."""
import glob
import json
import os
import time
import traceback

import numpy as np

from .. import core, mosaic, raster_io
from ..tools import log as logt


def _erase(outpath, search):
    for f in glob.glob(os.path.join(outpath, search)):
        if os.path.isfile(f):
            os.remove(f)


class TaudemCheck():

    def __init__(self, work_dir, basin_work_dir, taudem_dir, projectname):
        self.starttime = time.time()
        self.basin_work_dir = basin_work_dir
        self.taudem_dir = taudem_dir
        self.projectname = projectname
        self.taudem_projdir = os.path.join(self.taudem_dir, self.projectname)
        self.filled_DEM_vrt = os.path.join(self.taudem_projdir, 'fel.vrtd', 'orig.vrt')
        self.flow_dir_vrt = os.path.join(self.taudem_projdir, 'p.vrtd', 'orig.vrt')
        self.edm_mosaic = os.path.join(os.path.dirname(os.path.dirname(work_dir)), "DTM_output_mosaics",
                                       "{0}_EDM.tif".format(self.projectname))
        self.nodataval, self.nodataval_d8 = -9999.0, 99
        self.n_fixed = 0

    @staticmethod
    def _read_any(path):
        return mosaic.read_vrt(path) if path.endswith('.vrt') else raster_io.read_raster(path)

    def _basin_rasters(self, outpath_basin):
        """the four per-basin rasters (see the module docstring)"""
        dem_tif = os.path.join(outpath_basin, 'taudem_DEM.tif')
        d8_tif = os.path.join(outpath_basin, 'taudem_d8_flowdir.tif')
        flat_tif = os.path.join(outpath_basin, 'taudem_flatten_mask.tif')
        edm_tif = os.path.join(outpath_basin, 'taudem_EDM.tif')
        if not (os.path.exists(dem_tif) and os.path.exists(d8_tif)):
            with open(os.path.join(outpath_basin, 'window.json')) as fh:
                w = json.load(fh)
            win = (slice(w["row0"], w["row1"]), slice(w["col0"], w["col1"]))
            fel, fmeta = self._read_any(self.filled_DEM_vrt)
            p, pmeta = self._read_any(self.flow_dir_vrt)
            raster_io.write_raster(dem_tif, np.ascontiguousarray(fel[win], np.float32), fmeta,
                                   nodata=fmeta.nodata if fmeta.nodata is not None else -9999.0)
            d8 = np.ascontiguousarray(p[win]).astype(np.int16)
            if pmeta.nodata is not None:
                d8[d8 == np.int16(pmeta.nodata)] = 99
            raster_io.write_raster(d8_tif, d8, pmeta, nodata=99)
        if not os.path.exists(flat_tif):
            fm, m = raster_io.read_raster(os.path.join(outpath_basin, 'Water_flatten_mask.tif'))
            raster_io.write_raster(flat_tif, fm.astype(np.uint8), m, nodata=0)
        if not os.path.exists(edm_tif):
            src = os.path.join(outpath_basin, 'EDM.tif')
            dem, m = raster_io.read_raster(dem_tif)
            edm = raster_io.read_raster(src)[0].astype(np.uint8) if os.path.exists(src) else np.zeros(dem.shape, np.uint8)
            raster_io.write_raster(edm_tif, edm, m, nodata=0)
        return dem_tif, d8_tif, flat_tif, edm_tif

    def check_basin(self, basin_id):
        """:54-179.  Returns the number of problem cells found (None for skipped / failed basins)."""
        outpath_basin = os.path.join(self.basin_work_dir, 'basin_{0}'.format(basin_id))
        try:
            if os.path.exists(os.path.join(outpath_basin, 'coast.txt')):
                print('   {0} - coastal basin - skip'.format(basin_id))
                return None
            dem_tif, d8_tif, flat_tif, edm_tif = self._basin_rasters(outpath_basin)
            DEM_array, dmeta = raster_io.read_raster(dem_tif)
            d8_array, pmeta = raster_io.read_raster(d8_tif)
            flatten_mask_array = raster_io.read_raster(flat_tif)[0]
            edm_array = raster_io.read_raster(edm_tif)[0]
            self.nodataval = dmeta.nodata if dmeta.nodata is not None else -9999.0
            self.nodataval_d8 = pmeta.nodata if pmeta.nodata is not None else 99
            p8 = np.where(d8_array == self.nodataval_d8, 255, d8_array).astype(np.uint8)
            _, n_nodata_d8 = core.taudem_check(p8, DEM_array.astype(np.float32), flatten_mask_array.astype(np.uint8),
                                               edm_array.astype(np.uint8), nodata=float(self.nodataval))
            if n_nodata_d8 == 0 or os.path.exists(os.path.join(outpath_basin, 'endo_sink.tif')):
                _erase(outpath_basin, 'taudem_*')
                return n_nodata_d8
            print('  basin {0}: nodata values in flow direction, raise river elevations by 10 cm'.format(basin_id))
            self.fix_basin(outpath_basin, dem_tif, DEM_array, None, flatten_mask_array)
            return n_nodata_d8
        except Exception as e:  # noqa: BLE001  (reported and skipped, as upstream :168-174)
            print("Exception occurred on basin {0}: {1}".format(basin_id, e))
            print(traceback.format_exc())
            _erase(outpath_basin, 'taudem_*')
            return None

    def fix_basin(self, outpath_basin, basinDEM_tif, DEM_array, basin_shpfile, flatten_mask_array):
        """:183-194: water cells (1 < mask < 6) += 0.10 m; nodata becomes NaN in `conditionedDEM_taudem_adjust.tif`"""
        _, meta = raster_io.read_raster(basinDEM_tif)
        out = core.taudem_fix(np.ascontiguousarray(DEM_array, np.float32), np.ascontiguousarray(flatten_mask_array, np.uint8),
                              nodata=float(self.nodataval), bump=0.10)
        out = np.asarray(out, np.float32)
        out[np.asarray(DEM_array) == self.nodataval] = np.nan
        raster_io.write_raster(os.path.join(outpath_basin, "conditionedDEM_taudem_adjust.tif"), out, meta, nodata=float('nan'))
        self.n_fixed += 1
        return out


def main(argv=None):
    """`python -m pydrodem_b200.models.taudem_check -c config.ini -l log` (:204-247)"""
    import argparse
    from .. import config as cfgmod, units
    starttime = time.time()
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config", required=True, help="config file")
    ap.add_argument("-l", "--logfile", required=False, help="log file")
    args = ap.parse_args(argv)
    config = cfgmod.load(args.config)
    taudem_dir = config.get("paths-taudem", "taudem_dir")
    projectname = config.get("outputs", "projectname")
    basin_levels = json.loads(config.get("parameters-DH-basin", "basin_levels"))
    work_dir = os.path.join(config.get("paths-DH", "output_dir"), projectname)
    basin_work_dir = os.path.join(work_dir, 'basins_lvl{0:02d}'.format(basin_levels[0]))
    tau_check = TaudemCheck(work_dir, basin_work_dir, taudem_dir, projectname)
    basinIDs = logt.find_all_basins_IDs(basin_work_dir)
    units.run_units(tau_check.check_basin, basinIDs)
    if args.logfile and int(os.environ.get("RANK", "0")) == 0:
        with open(args.logfile, 'w') as f:
            f.write(logt.log_time("TAUDEM CHECK", starttime))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
