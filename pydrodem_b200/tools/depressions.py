"""Mirror of the hot-path functions of the reference's `tools/depressions.py`.

Same signatures and file protocol (`<in>_WL.tif`, `<in>_WL_DIFF.tif`, clump-ID rasters); GDAL is
replaced by pydrodem_b200.raster_io, WhiteboxTools by the `wbt` duck type (pydrodem_b200.wbt) and
scipy.ndimage.label / convolve2d by the CUDA kernels.  The pit catalog / solution selection
functions (build_pit_catalog, select_solution, apply_solution, combined_solution) and breach_pits
are not built yet.
"""
import os

import numpy as np

from .. import core, raster_io, stencils


def fill_pits(inDEM, wbt, fill_method='WL', nodataval=-9999, fixflats=False, flat_increment=0.0, save_tif=True,
              downcast=True):
    """tools/depressions.py:1236-1316.  Returns (fill_array, diff_fill_array), writes `<in>_<method>.tif`
    and (save_tif) `<in>_<method>_DIFF.tif`.  `downcast` is moot: the fill is written as float32 already."""
    fill_tif = inDEM.replace('.tif', '_{0}.tif'.format(fill_method))
    if flat_increment > 0.0:
        fixflats = True
    if fill_method == "SF":
        wbt.fill_depressions(inDEM, fill_tif, max_depth=None, fix_flats=fixflats, flat_increment=flat_increment)
    elif fill_method == "WL":
        wbt.fill_depressions_wang_and_liu(inDEM, fill_tif, fix_flats=fixflats, flat_increment=flat_increment)
    else:
        raise ValueError("   Improper fill method specified")
    inDEM_array, meta = raster_io.read_raster(inDEM)
    fill_array, _ = raster_io.read_raster(fill_tif)
    diff_fill_array = np.round((fill_array - inDEM_array), 3)  # :1308
    if save_tif:
        diff_tif_fill = fill_tif.replace('.tif', '_DIFF.tif')
        raster_io.write_raster(diff_tif_fill, diff_fill_array.astype(np.float32), meta, nodata=nodataval)
    return fill_array, diff_fill_array


def create_clump_array(input_array, clump_ID_tif, stencil_tif, wbt, nodataval=-9999, noflow_array=None,
                       condition='greaterthan', clumpbuffer=None, split_mask=None, save_tif=True):
    """tools/depressions.py:1593-1675: sign mask of a difference map -> optional (2b+1)^2 box dilation ->
    labels (8-connectivity, or 4-connectivity when buffered), ids in scipy's raster-scan order."""
    input_array = np.asarray(input_array)
    clump_array = np.zeros(input_array.shape, dtype=np.int8)
    if condition == 'greaterthan':
        clump_array[input_array > 0] = 1
    if condition == 'lessthan':
        if noflow_array is not None:
            clump_array[noflow_array == 1] = 1
        clump_array[input_array < 0] = 1
    if condition == 'nonzero':
        clump_array[input_array != 0] = 1
    if clumpbuffer is not None:
        sz = (2 * clumpbuffer) + 1
        clump_array = stencils.apply_convolve_filter(clump_array, np.ones((sz, sz)))
        clump_array = (clump_array > 0).astype(np.int8)
        conn = 4
    else:
        conn = 8
    if split_mask is not None:
        split_ind = np.nonzero(split_mask)
        split_clump_array = np.zeros_like(clump_array)
        split_clump_array[split_ind] = clump_array[split_ind]
        clump_array[split_ind] = 0
        clump_ID_array, num_feat = core.label(clump_array, conn)
        split_ID_array, _ = core.label(split_clump_array, conn)
        split_ind = np.nonzero(split_ID_array)
        split_ID_array[split_ind] += num_feat
        clump_ID_array[split_ind] = split_ID_array[split_ind]
    else:
        clump_ID_array, num_feat = core.label(clump_array, conn)
    if save_tif:
        meta = raster_io.read_raster(stencil_tif)[1] if (stencil_tif and os.path.exists(stencil_tif)) else None
        raster_io.write_raster(clump_ID_tif, clump_ID_array, meta, nodata=0)
    return clump_ID_array


def fix_shifted_data(in_array, shift_array, nodataval=-9999, failtest=False, vprint=False):
    """tools/depressions.py:1406-1468 repairs a one-pixel shift artefact of WhiteboxTools' GeoTIFF
    writer; the GPU path never shifts, so the arrays are returned unchanged."""
    return shift_array
