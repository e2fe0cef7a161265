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
            if self.carvemethod == 'LC':
                raise NotImplementedError("carve_method='LC' (breach_depressions_least_cost) is not built")
            results = dep.combined_solutions(list(DFpits_3.index.values), modDEM_array, nx, ny, DFpits, fillcells_dct,
                                             carvecells_dct, carve_fillcells_dct, self.combined_fill_interval, fillID_vec,
                                             fill_vec, carve_vec, self.radius, self.max_carve_depth,
                                             nodataval=self.nodataval)
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
