"""The per-unit depression pipeline of `DepHandle.depression_handling` + the final fill of
`run_depression_handling` (models/depression_handling.py:717-830, 905-913, 470-484), without the GIS around it.

The reference method spends its first 170 lines building masks from shapefiles (OSM water, basin walls, SWO
cutlines: GDAL / geopandas, out of scope) and then runs the sequence mirrored here on one buffered DEM GeoTIFF:

    initial_fill_carve -> build_pit_catalog -> select_solution -> apply_solution (solved pits)
      -> combined_solution for every pit coded 4000..4999 -> final fill (dep.fill_pits)

`depression_core` is that sequence with the same helper calls, argument order and intermediate file names
(`*_apply_TEMP.tif`, `*_conditionedTEMP.tif`, `<basename>_solution_mask.tif`); the masks arrive as arrays.
Fills, labels, catalog reductions and selection run on the GPU (libhydrocore); the breach's priority-order part runs
on the host (csrc/breach.cu).
"""
import os

import numpy as np

from .. import raster_io
from ..tools import depressions as dep
from ..wbt import WhiteboxToolsGPU


class DepHandleCore:
    """The subset of DepHandle's configuration (models/depression_handling.py:41-230; defaults of
    config_files/config_local.ini [parameters-depression-handling]) that the array-level pipeline reads."""

    def __init__(self, basename="DEM", nodataval=-9999, fill_method='WL', carvemethod='SC', maxcost=200, radius=50,
                 min_dist=False, sfill=1.0, flat_increment=0.0, fill_vol_thresh=12000, carve_vol_thresh=4000,
                 max_carve_depth=100, combined_fill_interval=0.2, del_temp_files=False, verbose_print=False, wbt=None):
        self.basename = basename
        self.nodataval = nodataval
        self.fill_method = fill_method
        self.carvemethod = carvemethod
        self.maxcost = maxcost
        self.radius = radius
        self.min_dist = min_dist
        self.sfill = sfill
        self.flat_increment = flat_increment
        self.fill_vol_thresh = fill_vol_thresh
        self.carve_vol_thresh = carve_vol_thresh
        self.max_carve_depth = max_carve_depth
        self.combined_fill_interval = combined_fill_interval
        self.del_temp_files = del_temp_files
        self.verbose_print = verbose_print
        self.starttime = None
        self.wbt = wbt if wbt is not None else WhiteboxToolsGPU()

    def depression_core(self, inDEM, outpath, water_mask_array=None, wall_array=None, sink_mask=None):
        """models/depression_handling.py:717-830 and :905-913.  Returns (modDEMoutput_array, DFpits,
        solution_mask_array)."""
        _, meta = raster_io.read_raster(inDEM)
        modDEM_vec, fill_vec, carve_vec, diff_vec_fill, diff_vec_carve, fillID_vec, carveID_vec, carvefillID_vec, \
            nx, ny = dep.initial_fill_carve(inDEM, self.basename, outpath, self.fill_method, self.maxcost, self.radius,
                                            self.min_dist, self.nodataval, self.starttime, self.wbt, sfill=self.sfill,
                                            flat_increment=self.flat_increment, watermask=None,
                                            water_mask_array=water_mask_array, carveout=False,
                                            carve_method=self.carvemethod, verbose_print=self.verbose_print,
                                            del_temp_files=self.del_temp_files)
        DFpits, fillcells_dct, carvecells_only_dct, carvecells_dct, carve_fillcells_dct = dep.build_pit_catalog(
            modDEM_vec, self.nodataval, carve_vec, diff_vec_carve, diff_vec_fill, fillID_vec, carveID_vec,
            carvefillID_vec, verbose_print=self.verbose_print)
        if water_mask_array is None:
            # the reference raises NameError here (tools/depressions.py:606-607 vs :647); an all-zero mask is the
            # configuration that completes
            water_mask_array = np.zeros((ny, nx), np.uint8)
        DFpits = dep.select_solution(DFpits, fillcells_dct, carvecells_dct, carvecells_only_dct, fill_vec, carve_vec,
                                     fillID_vec, carveID_vec, self.fill_vol_thresh, self.carve_vol_thresh,
                                     self.max_carve_depth, apply_shallow_fill=False, wall_array=wall_array,
                                     water_mask_array=water_mask_array, sink_array=sink_mask)
        excludedpits = list(DFpits[DFpits.solution > 3000].index.values)
        modDEM_vec, solution_mask_vec = dep.apply_solution(DFpits, fillcells_dct, carvecells_dct, carve_fillcells_dct,
                                                           modDEM_vec, fill_vec, carve_vec, excludedpits=excludedpits)
        modDEM_vec = np.asarray(modDEM_vec)
        solution_mask_vec = np.asarray(solution_mask_vec)
        modDEM_array = modDEM_vec.reshape(ny, nx)
        modDEM_tif = inDEM.replace('.tif', '_apply_TEMP.tif')
        raster_io.write_raster(modDEM_tif, modDEM_array.astype(np.float32), meta, nodata=self.nodataval)

        DFpits_3 = DFpits[(DFpits.solution >= 4000) & (DFpits.solution < 5000)].copy()
        if len(DFpits_3.index.values) > 0:
            modDEM_array, _ = raster_io.read_raster(modDEM_tif)
            # every pit's partial-fill candidates are evaluated from the same pre-combination surface (the reference's
            # loop reads modDEM_array once, :768), so all of them go through one batched device pass
            results = dep.combined_solutions(list(DFpits_3.index.values), modDEM_array, nx, ny, DFpits, fillcells_dct,
                                             carvecells_dct, carve_fillcells_dct, self.combined_fill_interval, fillID_vec,
                                             fill_vec, carve_vec, self.radius, self.max_carve_depth,
                                             nodataval=self.nodataval, carve_method=self.carvemethod,
                                             maxcost=self.maxcost, min_dist=self.min_dist,
                                             flat_increment=self.flat_increment if self.carvemethod == 'LC' else 0.0)
            for pit_id in DFpits_3.index.values:
                solution_name, sol_df = results[pit_id]
                DFpits.at[pit_id, 'solution'] = solution_name
                combined_footprint_ind = sol_df.footprint.values[0]
                combined_footprint_elevs = sol_df.elevs.values[0]
                modDEM_vec[combined_footprint_ind] = combined_footprint_elevs
                solution_mask_vec[combined_footprint_ind] = solution_name
        modDEMoutput_array = modDEM_vec.reshape(ny, nx)

        nodata_ind = np.where(modDEMoutput_array == self.nodataval)
        solution_mask_array = solution_mask_vec.reshape(ny, nx)
        solution_mask_array[nodata_ind] = -99
        if not self.del_temp_files:
            solution_mask_tif = os.path.join(outpath, f'{self.basename}_solution_mask.tif')
            raster_io.write_raster(solution_mask_tif, solution_mask_array.astype(np.int16), meta, nodata=-99)
        return modDEMoutput_array, DFpits, solution_mask_array

    def run_depression_core(self, inDEM, outpath, water_mask_array=None, wall_array=None, sink_mask=None):
        """depression_core + the final fill of run_depression_handling (:470-484).  Returns (ffill_array,
        modDEMoutput_array, DFpits, solution_mask_array); ffill_array is the conditioned DEM."""
        # (rasters nobody reads back at once -- difference maps, id rasters, masks -- are written behind the pipeline's
        #  back; every file is complete when this method returns)
        with raster_io.write_behind():
            _, meta = raster_io.read_raster(inDEM)
            outDEM_array, DFpits, solution_mask_array = self.depression_core(inDEM, outpath, water_mask_array, wall_array,
                                                                            sink_mask)
            outDEM_tif_temp = inDEM.replace('.tif', '_conditionedTEMP.tif')
            raster_io.write_raster(outDEM_tif_temp, outDEM_array.astype(np.float32), meta, nodata=self.nodataval)
            ffill_array, _ = dep.fill_pits(outDEM_tif_temp, self.wbt, save_tif=False, nodataval=self.nodataval)
            raster_io.write_raster(outDEM_tif_temp, ffill_array.astype(np.float32), meta, nodata=self.nodataval)
        return ffill_array, outDEM_array, DFpits, solution_mask_array


class DepHandle(DepHandleCore):
    """`models/depression_handling.py:DepHandle` (:45-260) with the reference's constructor signature, unit directory
    layout and output files, for the raster half of the flow.

    The reference cuts every unit's rasters out of region mosaics with shapefile cutlines (GDAL / geopandas: out of
    scope); here a unit directory `<work_dir>/<base_unit>_<id>/` must already hold the buffered window
    `inDEM.tif` the reference creates at :348 (or a `window.json` {"row0", "row1", "col0", "col1"} from which it is
    cut out of `inDS_vrt`, a same-grid GeoTIFF / .vrt), optionally `inDEM_Water_mask.tif` (:428) and
    `inDEM_boundary_donut.tif` (:451).  `run_depression_handling` then follows :340-536: single-cell pits (if asked),
    the depression pipeline, the final fill, `conditionedDEM.tif` + `conditionedDEM_DIFF.tif` (LZW, as
    pyct.tifcompress leaves them), `Water_mask.tif`, `DEMcompleted.txt`; an exception is reported and swallowed per
    unit exactly as upstream (:535-540).  River burning / flattening, no-flow walls and endorheic sink enforcement
    need the vector half and raise NotImplementedError when switched on."""

    def __init__(self, inDS_vrt, work_dir, SWO_vrt, projectname, basin_levels, WBT_path=None, base_unit='tile', OSM_dir=None,
                 edm_vrt=None, taudem_dir=None, overwrite=False, run_parallel=True, verbose_print=True, del_temp_files=False,
                 noflow_bound=False, burn_rivers=False, flattened_mask_vrt=None, basins_merge_shp=None,
                 basins_dissolve_shp=None, reservoir_barrier=False, fill_h20_adjacent=False, buffer=100, coastal_buffer=0,
                 sfill=2.0, fill_single_cell_pits=False, flat_increment=0.0, fill_method="WL", carve_method='SC', maxcost=200,
                 carveout=False, radius=25, min_dist=False, wall_height=50.0, fill_vol_thresh=12000, carve_vol_thresh=4000,
                 max_carve_depth=100, combined_fill_interval=.2, enforce_endo=False, enforce_monotonicity=True,
                 noflow_points=None, wall_remove_min=0, osm_coast_clip=True, wbt=None):
        super().__init__(basename="inDEM", nodataval=-9999, fill_method=fill_method, carvemethod=carve_method,
                         maxcost=maxcost, radius=radius, min_dist=min_dist, sfill=sfill, flat_increment=flat_increment,
                         fill_vol_thresh=fill_vol_thresh, carve_vol_thresh=carve_vol_thresh,
                         max_carve_depth=max_carve_depth, combined_fill_interval=combined_fill_interval,
                         del_temp_files=del_temp_files, verbose_print=verbose_print, wbt=wbt)
        import time
        self.starttime = time.time()
        self.inDS_vrt, self.work_dir, self.SWO_vrt, self.projectname = inDS_vrt, work_dir, SWO_vrt, projectname
        self.basin_levels, self.WBT_path, self.base_unit = basin_levels, WBT_path, base_unit
        self.overwrite, self.run_parallel = overwrite, run_parallel
        self.fill_single_cell_pits = fill_single_cell_pits
        self.buffer, self.coastal_buffer, self.carveout = buffer, coastal_buffer, carveout
        for name, on in (("burn_rivers", burn_rivers), ("noflow_bound", noflow_bound), ("reservoir_barrier", reservoir_barrier),
                         ("fill_h20_adjacent", fill_h20_adjacent), ("enforce_endo", enforce_endo)):
            if on:
                raise NotImplementedError(f"DepHandle({name}=True) needs the reference's vector / OSM half (out of scope)")
        self.enforce_monotonicity = enforce_monotonicity

    def fillcarve_single_cell_pits(self, inDEM, outDEM):
        """:257-261"""
        self.wbt.breach_single_cell_pits(inDEM, outDEM)
        self.wbt.fill_single_cell_pits(outDEM, outDEM)

    def _unit_inputs(self, outpath):
        import json
        from .. import mosaic
        inDEM = os.path.join(outpath, 'inDEM.tif')
        if not os.path.exists(inDEM):
            win = os.path.join(outpath, 'window.json')
            if not os.path.exists(win):
                raise FileNotFoundError(f"{outpath}: neither inDEM.tif nor window.json (see DepHandle's docstring)")
            with open(win) as fh:
                w = json.load(fh)
            arr, meta = (mosaic.read_vrt(self.inDS_vrt) if self.inDS_vrt.endswith('.vrt')
                         else raster_io.read_raster(self.inDS_vrt))
            sub = np.ascontiguousarray(arr[w["row0"]:w["row1"], w["col0"]:w["col1"]])
            raster_io.write_raster(inDEM, sub.astype(np.float32), dep._window_meta(meta, w["row0"], w["col0"]),
                                   nodata=meta.nodata if meta.nodata is not None else self.nodataval)
        return inDEM

    def run_depression_handling(self, input_id, SWO_cutoff=20, initial_burn=2.0, SWO_burn=2.0, burn_cutoff=6.0,
                                gap_dist=240, apply_relative_min=True):
        """:262-540 for one unit; returns the conditioned DEM array (None when the unit failed: see error_log_main.txt)"""
        import traceback
        outpath = os.path.join(self.work_dir, f'{self.base_unit}_{input_id}')
        self.outpath, self.input_id = outpath, input_id
        errorlog = os.path.join(outpath, 'error_log_main.txt')
        completed_file = os.path.join(outpath, 'DEMcompleted.txt')
        outDEM_tif = os.path.join(outpath, "conditionedDEM.tif")
        try:
            for f in (errorlog, completed_file, outDEM_tif):
                if os.path.exists(f):
                    os.remove(f)
            inDEM = self._unit_inputs(outpath)
            inDEM_array, meta = raster_io.read_raster(inDEM)
            if meta.nodata is not None:
                self.nodataval = meta.nodata
            wm_tif = os.path.join(outpath, f"{self.basename}_Water_mask.tif")
            water_mask_array = raster_io.read_raster(wm_tif)[0] if os.path.exists(wm_tif) else \
                np.zeros(inDEM_array.shape, np.uint8)      # (None crashes the reference's selection)
            donut_tif = os.path.join(outpath, f"{self.basename}_boundary_donut.tif")
            wall_array = raster_io.read_raster(donut_tif)[0] if os.path.exists(donut_tif) else None
            work_dem = inDEM
            if self.fill_single_cell_pits:
                work_dem = inDEM.replace('.tif', '_scp.tif')
                self.fillcarve_single_cell_pits(inDEM, work_dem)
            ffill_array, outDEM_array, DFpits, solmask = self.run_depression_core(work_dem, outpath, water_mask_array,
                                                                                 wall_array=wall_array)
            DFpits.to_csv(os.path.join(outpath, 'DFpits.csv'))        # (upstream: DFpits.h5; pytables is not a dependency here)
            raster_io.write_raster(os.path.join(outpath, "Water_mask.tif"), np.asarray(water_mask_array, np.uint8), meta,
                                   nodata=0, compress='lzw')
            diff = (ffill_array - inDEM_array).astype(np.float32)
            raster_io.write_raster(outDEM_tif.replace('.tif', '_DIFF.tif'), diff, meta, nodata=self.nodataval, compress='lzw')
            raster_io.write_raster(outDEM_tif, ffill_array.astype(np.float32), meta, nodata=self.nodataval, compress='lzw')
            with open(completed_file, 'w') as f:
                f.write('DEM processing complete')
            if self.del_temp_files:
                import glob
                for f in glob.glob(os.path.join(outpath, f"{self.basename}*")):
                    if os.path.isfile(f):
                        os.remove(f)
            return ffill_array
        except Exception as e:  # noqa: BLE001  (the reference reports and moves on to the next unit, :535-540)
            msg = f"Exception occurred on {input_id}: {e}\n{traceback.format_exc()}"
            print(msg)
            try:
                with open(errorlog, 'w') as f:
                    f.write(msg)
            except OSError:
                pass
            return None


def main(argv=None):
    """`python -m pydrodem_b200.models.depression_handling -c config.ini -l log` (:1419-1646): same sections and keys; the
    units (`<dh_work_dir>/<base_unit>_*`, unfinished ones unless overwrite) are spread over the job's ranks by
    units.run_units (one process per GPU under torchrun; a plain loop otherwise)."""
    import argparse
    import json
    import time
    from .. import config as cfgmod, units
    from ..tools import log as logt
    starttime = time.time()
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config", required=True, help="config file")
    ap.add_argument("-l", "--logfile", required=False, help="log file")
    args = ap.parse_args(argv)
    config = cfgmod.load(args.config)
    p, s = "processes-depression-handling", "parameters-depression-handling"
    opts = cfgmod.dephandle_options(config)
    burn = cfgmod.burn_options(config)
    if opts.get("sfill") == 0.0:
        opts["sfill"] = None
    base_unit = config.get("region", "base_unit")
    burn_rivers = config.getboolean(p, "burn_rivers")
    dh = DepHandle(config.get('outputs', 'burned_mosaic' if burn_rivers else 'dtm_mosaic'), config.get("outputs", "dh_work_dir"),
                   config.get("paths", "SWO_vrt"), config.get("outputs", "projectname"), json.loads(config.get(s, "basin_levels")),
                   config.get("paths", "WBT_path", fallback=None), base_unit=base_unit,
                   overwrite=config.getboolean(p, "overwrite"), verbose_print=opts.pop("verbose_print", True),
                   del_temp_files=opts.pop("del_temp_files", False), burn_rivers=burn_rivers,
                   noflow_bound=config.getboolean(p, "noflow_walls", fallback=False),
                   reservoir_barrier=config.getboolean(p, "reservoir_barrier", fallback=False),
                   fill_h20_adjacent=config.getboolean(p, "fill_h20_adjacent", fallback=False),
                   enforce_endo=config.getboolean(p, "enforce_endo", fallback=False),
                   enforce_monotonicity=config.getboolean(p, "enforce_monotonicity", fallback=True),
                   buffer=config.getfloat(s, "buffer", fallback=100), sfill=opts.pop("sfill"),
                   carve_method=opts.pop("carvemethod"), **opts)
    if dh.overwrite:
        ids = logt.find_all_ids(dh.work_dir, unit=base_unit, dir_search=f'{base_unit}_')
    else:
        ids = logt.find_unfinished_ids(dh.work_dir, unit=base_unit, dir_search=f'{base_unit}_', tifname='DEMcompleted.txt')
    import torch
    import torch.distributed as dist
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    units.run_units(lambda u: dh.run_depression_handling(u, **burn) is not None, ids)
    if args.logfile and int(os.environ.get("RANK", "0")) == 0:
        with open(args.logfile, 'w') as f:
            f.write(logt.log_time("Depression handling", starttime))
    return 0


if __name__ == '__main__':
    raise SystemExit(main())
