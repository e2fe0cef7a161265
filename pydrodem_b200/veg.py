"""Vegetation-bias lookup of CreateDTM (models/dem_noise_removal.py:345-364, 556-631, 633-807), array level.

The reference works on GeoTIFF paths and a JSON table on disk; the arithmetic between `gdal_array.LoadFile` and
`return bias_array, slope_veg_array, max_treeheight` is what lives here, on the GPU: the 3x3 means of tree cover /
tree height (`hc_box3_index_dev`), the 9x9 DW smoothing + `slope_D2` of the DEM (stencils), and the per-pixel table
lookup the reference does in a Python list comprehension (:802-805) as one gather kernel (`hc_lut3d_gather_dev`).
The nodata in-fill of the tree-cover raster (:692-762: two fall-back mosaics, WhiteboxTools fill_missing_data) is raster
I/O plus an off-path tool and is not mirrored: more than 10 cells above 100 raise NotImplementedError.

Index arrays are bit-exact for integer-valued tree cover / height rasters; the slope index inherits the float32
convolution's tolerance class (a cell whose smoothed slope sits within ~1e-6 relative of a whole degree may land in
the neighbouring slope bin).
"""
import ctypes
import json

import numpy as np

from . import _lib
from . import stencils
from .core import _is_torch, _stream, _tdev, _tp

SLOPE_CAP_3D = 19   # models/dem_noise_removal.py:792
SLOPE_CAP_2D = 50   # :616
HALO_ROWS = 5       # 9x9 smoothing (4) + central-difference slope (1): halo of a row band (bands.apply_banded)


def load_bias_cube(lookup_table):
    """The LUT preparation of compute_veg_bias_3D (:660-683): JSON {slope: [[bias over tree height] over tree cover]}
    -> float64 cube [tree cover, tree height, slope], zero-cover / zero-height planes zeroed, smoothed along the slope
    axis with [.25 .5 1 .5 .25]/2.5.  `lookup_table` is a path or the already-loaded dict.  Host side: the table has a
    few 10^4 entries and is prepared once per run; scipy's convolve1d is used as the reference uses it."""
    import scipy.ndimage
    data = json.load(open(lookup_table)) if isinstance(lookup_table, (str, bytes)) or hasattr(lookup_table, "__fspath__") \
        else lookup_table
    keys = list(data.keys())
    sample_arr = np.array(data.get(keys[0]))
    slope_axis = np.array(keys).astype(int)
    bias_cube = np.empty((sample_arr.shape[0], sample_arr.shape[1], len(keys)))
    for s in slope_axis:
        bias_cube[:, :, s] = np.array(data.get(str(int(s))))
    bias_cube[0, :, :] = 0.0
    bias_cube[:, 0, :] = 0.0
    weights = np.array([.25, .5, 1.0, .5, .25])
    weights = (1 / np.sum(weights)) * weights
    return scipy.ndimage.convolve1d(bias_cube, weights, axis=2)


def _dev2d(a, dtype, what):
    import torch
    if _is_torch(a):
        t = a
    else:
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensor required (no CPU fallback)")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def box3_index(array, zero_mask=None, in_max=None, zero_above=None, return_smooth=False):
    """astype(int) of apply_convolve_filter(array, np.ones((3, 3)), normalize_kernel=True), with the reference's
    pre/post steps folded in: `array[array > in_max] = in_max` (:777) and `array[zero_mask == 1] = 0` (:687) before,
    `smooth[smooth > zero_above] = 0` (:771) after.  Returns int32 indices (and the float32 smoothed array)."""
    import torch
    host = not _is_torch(array)
    a = _dev2d(array, torch.float32, "box3_index")
    m = None if zero_mask is None else _dev2d(zero_mask, torch.uint8, "box3_index")
    idx = torch.empty(a.shape, dtype=torch.int32, device=a.device)
    sm = torch.empty_like(a) if return_smooth else None
    inf = float("inf")
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().hc_box3_index_dev(_lib.context(a.device.index or 0), _tp(a), _tp(m),
                                                 inf if in_max is None else float(in_max), a.shape[0], a.shape[1],
                                                 inf if zero_above is None else float(zero_above), _tp(idx), _tp(sm),
                                                 _stream(a)), "hc_box3_index_dev")
    if host:
        idx = idx.cpu().numpy()
        sm = None if sm is None else sm.cpu().numpy()
    return (idx, sm) if return_smooth else idx


def lut_gather(cube, i0, i1, i2, cap2, zero_mask=None, dem=None):
    """bias[p] = cube[i0[p], i1[p], min(i2[p], cap2)] (float64); i2 (int8) is capped in place; bias is zeroed under
    zero_mask; with `dem` also returns dem - bias in float64 (:361-364).  Device tensors in, device tensors out."""
    import torch
    cube = np.ascontiguousarray(cube, dtype=np.float64)
    if cube.ndim == 2:
        cube = cube[:, None, :]
    dev = i0.device
    cube_d = torch.from_numpy(cube).to(dev)
    n = i0.numel()
    bias = torch.empty(i0.shape, dtype=torch.float64, device=dev)
    out = torch.empty(i0.shape, dtype=torch.float64, device=dev) if dem is not None else None
    oob = ctypes.c_int(0)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().hc_lut3d_gather_dev(_lib.context(dev.index or 0), _tp(i0), _tp(i1), _tp(i2), int(cap2),
                                                   _tp(cube_d), cube.shape[0], cube.shape[1], cube.shape[2],
                                                   _tp(zero_mask), _tp(bias), _tp(dem), _tp(out), n, ctypes.byref(oob),
                                                   _stream(i0)), "hc_lut3d_gather_dev")
    return bias, out


def _count_nodata(tc):
    if _is_torch(tc):
        return int((tc > 100).sum().item())
    return int(np.count_nonzero(np.asarray(tc) > 100))


def compute_veg_bias_3D(TC_array, Theight_array, Zarr, bias_cube, dx, dy, ocean_mask=None, watermaskTX=None,
                        return_dem=False):
    """CreateDTM.compute_veg_bias_3D (:633-807) from the loaded rasters: TC_array / Theight_array are what
    `gdal_array.LoadFile(treecover_tif / treeheight_tif)` returns, `bias_cube` is `load_bias_cube(...)`, `ocean_mask`
    is `self.ocean_mask` (:338).  Returns (bias_array float64, slope_veg_array, max_treeheight) like the reference.

    `watermaskTX` (a 0/1 raster of `(edm == 2) | (edm == 3)`, :337) folds in the two lines that follow the call in
    run_tile -- `bias_array[watermaskTX] = 0.0` and, with return_dem=True, `dem_array - bias_array` (float64) as a
    fourth return value."""
    import torch
    host = not _is_torch(Zarr)
    if _count_nodata(TC_array) > 10:
        raise NotImplementedError("tree-cover nodata in-fill (models/dem_noise_removal.py:692-762) is raster I/O + "
                                  "WhiteboxTools fill_missing_data and is not on the GPU path; fill the raster first")
    bias_cube = np.asarray(bias_cube, dtype=np.float64)
    max_treeheight = bias_cube.shape[1] - 1
    z = _dev2d(Zarr, torch.float32, "compute_veg_bias_3D")
    z = _tdev(z, torch.float32, "compute_veg_bias_3D")
    tc = box3_index(_dev2d(TC_array, torch.float32, "TC_array"),
                    None if ocean_mask is None else _dev2d(ocean_mask, torch.uint8, "ocean_mask"), zero_above=100.0)
    th = box3_index(_dev2d(Theight_array, torch.float32, "Theight_array"), in_max=max_treeheight)
    kernel = stencils.build_kernel([1., 0.8, 0.6, 0.4, 0.2], dtype=np.float32, normalize=True)
    slope = stencils.slope_D2(stencils.apply_convolve_filter(z, kernel), dx, dy, unit='degree')
    wm = None if watermaskTX is None else _dev2d(watermaskTX, torch.uint8, "watermaskTX")
    bias, dem = lut_gather(bias_cube, tc, th, slope, SLOPE_CAP_3D, wm, z if return_dem else None)
    if host:
        bias, slope = bias.cpu().numpy(), slope.cpu().numpy()
        dem = None if dem is None else dem.cpu().numpy()
    return (bias, slope, max_treeheight, dem) if return_dem else (bias, slope, max_treeheight)


def compute_veg_bias_2Dtable(TC_array, Zarr, bias_table, dx, dy, watermaskTX=None, return_dem=False):
    """The '2Dtable' branch of CreateDTM.compute_veg_bias_2D (:580-590, :610-631): tree cover above 100 -> 0, slope of
    the 9x9-smoothed DEM capped at 50, bias = bias_table[slope, treecover] (`bias_table` = the pivoted
    slope x treecover table, :622-624).  Returns (bias_array float64, slope_veg_array)."""
    import torch
    host = not _is_torch(Zarr)
    z = _tdev(_dev2d(Zarr, torch.float32, "compute_veg_bias_2Dtable"), torch.float32, "compute_veg_bias_2Dtable")
    tc = _dev2d(TC_array, torch.int32, "TC_array")          # Byte raster: 0..255
    kernel = stencils.build_kernel([1., 0.8, 0.6, 0.4, 0.2], dtype=np.float32, normalize=True)
    slope = stencils.slope_D2(stencils.apply_convolve_filter(z, kernel), dx, dy, unit='degree')
    table = np.asarray(bias_table, dtype=np.float64).T                         # -> [treecover, slope]
    # `TC_array[TC_array > 100] = 0` (:583) as table rows: cover 101..255 reads the cover-0 row
    table = np.ascontiguousarray(np.concatenate([table[:101], np.repeat(table[:1], 155, axis=0)]))
    wm = None if watermaskTX is None else _dev2d(watermaskTX, torch.uint8, "watermaskTX")
    bias, dem = lut_gather(table, tc, None, slope, SLOPE_CAP_2D, wm, z if return_dem else None)
    if host:
        bias, slope = bias.cpu().numpy(), slope.cpu().numpy()
        dem = None if dem is None else dem.cpu().numpy()
    return (bias, slope, dem) if return_dem else (bias, slope)
