"""`models/find_sinks.py` of the reference (:31-247) on the GPU path: one sink point per endorheic basin -- the lowest
pixel of the largest connected blob among the lowest 1 % of the basin's cells, preferring persistent water (SWO >= 80).

Same class, constructor, method name and `-c config -l log` entry point.  The reference crops the region mosaic to the
basin polygon with a GDAL cutline and writes `sink.shp` through geopandas (both out of scope); here a
basin directory must already hold the crop `DEM_endo.tif` (nodata = +99999 so that it never wins a minimum, :55) and
optionally `SWO_endo.tif` on the same grid; the sink goes to `sink.geojson` (a one-point FeatureCollection in the
raster's coordinates, EPSG:4326 upstream) and, as the raster the later stages use, `endo_sink.tif` (1 at the sink).
Single-cell pits are filled first (:118, `hc_fill_single_cell_pits_dev`); the selection is `tools/endo.find_sink_pixel`."""
import json
import os
import time
import traceback

import numpy as np

from .. import raster_io
from ..tools import endo
from ..tools import log as logt
from ..wbt import WhiteboxToolsGPU


class FindSinks():

    def __init__(self, inDS_vrt, endo_work_dir, SWO_vrt, projectname, WBT_path=None, verbose_print=True, rank=0, wbt=None):
        self.WBT_path = WBT_path
        self.rank = rank
        self.starttime = time.time()
        self.inDS_vrt = inDS_vrt
        self.endo_work_dir = endo_work_dir
        self.SWO_vrt = SWO_vrt
        self.projectname = projectname
        self.verbose_print = verbose_print
        self.wbt = wbt if wbt is not None else WhiteboxToolsGPU()

    def endo_basin_setup(self, input_id, endo_nodata=99999, overwrite=False):
        """:55-138.  Returns (x, y) of the sink, or None (skipped / failed)."""
        outpath = os.path.join(self.endo_work_dir, 'basin_{0}'.format(input_id))
        sink_point_file = os.path.join(outpath, "sink.geojson")
        if os.path.exists(sink_point_file) and not overwrite:
            return None
        try:
            endoDEM = os.path.join(outpath, 'DEM_endo.tif')
            if not os.path.exists(endoDEM):
                raise FileNotFoundError(f"{endoDEM}: the basin crop must exist (see the module docstring)")
            endoDEM_sc = os.path.join(outpath, 'DEM_endo_sc.tif')
            self.wbt.fill_single_cell_pits(endoDEM, endoDEM_sc)
            z, meta = raster_io.read_raster(endoDEM_sc)
            swo_tif = os.path.join(outpath, 'SWO_endo.tif')
            swo = raster_io.read_raster(swo_tif)[0] if os.path.exists(swo_tif) else None
            gt = meta.geotransform if hasattr(meta, "geotransform") else None
            if gt is None:
                dx, dy = meta.pixel_size
                tie = meta.get("geo", {}).get(33922)
                x0, y0 = (float(tie[1][3]), float(tie[1][4])) if tie else (0.0, 0.0)
                gt = (x0, dx, 0.0, y0, 0.0, -dy)
            r, c = endo.find_sink_pixel(z, swo, geotransform=gt)
            x, y = endo.get_xy(r, c, gt)
            with open(sink_point_file, 'w') as fh:
                json.dump({"type": "FeatureCollection", "features": [
                    {"type": "Feature", "properties": {"basin_id": int(input_id), "id": 0, "row": int(r), "col": int(c)},
                     "geometry": {"type": "Point", "coordinates": [x, y]}}]}, fh)
            mask = np.zeros(z.shape, np.uint8)
            mask[r, c] = 1
            raster_io.write_raster(os.path.join(outpath, 'endo_sink.tif'), mask, meta, nodata=0)
            os.remove(endoDEM_sc)
            return (x, y)
        except Exception as e:  # noqa: BLE001  (reported and skipped, as upstream :131-137)
            print("Exception occurred on rank {0}, basin {1}: {2}".format(self.rank, input_id, e))
            print(traceback.format_exc())
            return None


def main(argv=None):
    """`python -m pydrodem_b200.models.find_sinks -c config.ini -l log` (:146-247)"""
    import argparse
    from .. import config as cfgmod, units
    starttime = time.time()
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config", required=True, help="config file")
    ap.add_argument("-l", "--logfile", required=False, help="log file")
    args = ap.parse_args(argv)
    config = cfgmod.load(args.config)
    overwrite = config.getboolean("processes-depression-handling", "overwrite")
    projectname = config.get("outputs", "projectname")
    endo_work_dir = os.path.join(config.get("paths", "output_dir_cond"), projectname, 'endo_basins')
    if overwrite:
        endo_ids = logt.find_all_ids(endo_work_dir, unit='basin', dir_search='basin_')
    else:
        endo_ids = logt.find_unfinished_ids(endo_work_dir, unit='basin', dir_search='basin_', tifname='sink.geojson')
    fs = FindSinks(config.get("outputs", "dtm_mosaic"), endo_work_dir, config.get("paths", "SWO_vrt"), projectname,
                   config.get("paths", "WBT_path", fallback=None), verbose_print=False,
                   rank=int(os.environ.get("RANK", "0")))
    units.run_units(lambda b: fs.endo_basin_setup(b, overwrite=overwrite), endo_ids)
    if args.logfile and int(os.environ.get("RANK", "0")) == 0:
        with open(args.logfile, 'w') as f:
            f.write(logt.log_time("ENDO SINK FINDING", starttime))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
