"""The array-level water stencils of `tools/water.py` on the GPU, with the reference's signatures.

`river_minimize` (:786), `smooth_rivers` (:928), `raise_small_streams` (:960) and `raise_small_rivers` (:1010)
are chains of mask predicates, `scipy.ndimage.minimum_filter` / `maximum_filter` over square windows of up to 51
cells, a 5x5 mean and a per-cell select; csrc/water.cu runs each chain as a handful of kernels with its scratch
rasters in the context's arena.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.  Like the reference,
`DEM_burn_arr` is modified in place and returned.  Rasters are float32 (the reference's DEMs are); the water mask
may be any integer type with values 0..255 and the SWO raster any integer type that fits int16.

`minimum_filter` / `maximum_filter` (size=k, mode 'reflect') are exported too, for uint8 and float32 rasters.
Not mirrored: `enforce_monotonic_rivers` / `enforce_monotonic_streams` / `stream_minimize` / `raise_canals`
(GeoTIFF paths, WhiteboxTools and TauDEM subprocesses inside the function body).
"""
import numpy as np

from .. import _lib
from ..core import _is_torch, _stream, _tp


def _dev(a, dtype, what):
    import torch
    t = a if _is_torch(a) else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensor required (no CPU fallback)")
    if t.dim() != 2:
        raise ValueError(f"{what}: 2-D raster expected")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _call(name, dev, *args):
    import torch
    with torch.cuda.device(dev):
        _lib.check(getattr(_lib.load(), name)(_lib.context(dev.index or 0), *args), name)


def _minmax(array, size, is_max):
    import torch
    host = not _is_torch(array)
    src = array if not host else np.asarray(array)
    if (src.dtype == torch.uint8) if not host else (src.dtype in (np.uint8, np.bool_)):
        a, eb = _dev(src if not host else src.astype(np.uint8), torch.uint8, "minmax_filter"), 1
    else:
        a, eb = _dev(src, torch.float32, "minmax_filter"), 4
    out = torch.empty_like(a)
    _call("hc_minmax_filter_dev", a.device, _tp(a), _tp(out), a.shape[0], a.shape[1], int(size), int(is_max), eb,
          _stream(a))
    return out.cpu().numpy() if host else out


def minimum_filter(array, size):
    """scipy.ndimage.minimum_filter(array, size=size) for a uint8 or float32 raster"""
    return _minmax(array, size, False)


def maximum_filter(array, size):
    """scipy.ndimage.maximum_filter(array, size=size) for a uint8 or float32 raster"""
    return _minmax(array, size, True)


def _finish_inplace(given, result):
    """the reference returns the (modified) DEM_burn_arr object itself"""
    if _is_torch(given):
        if given.data_ptr() != result.data_ptr():
            given.copy_(result)
        return given
    if isinstance(given, np.ndarray) and given.dtype == np.float32 and given.flags.writeable:
        given[...] = result.cpu().numpy()
        return given
    return result.cpu().numpy()


def river_minimize(inDEM_array, water_mask_array, SWO_array, max_burn=5, river_filter_size=3, nodataval=-9999):
    """tools/water.py:786-847"""
    import torch
    host = not _is_torch(inDEM_array)
    dem = _dev(inDEM_array, torch.float32, "river_minimize")
    wm = _dev(water_mask_array, torch.uint8, "water_mask_array")
    swo = _dev(SWO_array, torch.int16, "SWO_array")
    out = torch.empty_like(dem)
    _call("hc_river_minimize_dev", dem.device, _tp(dem), _tp(wm), _tp(swo), float(nodataval), float(max_burn),
          int(river_filter_size), _tp(out), dem.shape[0], dem.shape[1], _stream(dem))
    return out.cpu().numpy() if host else out


def raise_small_rivers(inDEM_array, DEM_burn_arr, water_mask_array, SWO_array, nodata, SWO_cutoff_high=50,
                       SWO_cutoff_low=5, extra_raise=0.10, max_raise=3.0, filter_size=51):
    """tools/water.py:1010-1045"""
    import torch
    same = inDEM_array is DEM_burn_arr
    burn = _dev(DEM_burn_arr, torch.float32, "DEM_burn_arr")
    dem = burn if same else _dev(inDEM_array, torch.float32, "raise_small_rivers")
    wm = _dev(water_mask_array, torch.uint8, "water_mask_array")
    swo = _dev(SWO_array, torch.int16, "SWO_array")
    _call("hc_raise_small_rivers_dev", burn.device, _tp(dem), _tp(burn), _tp(wm), _tp(swo), float(nodata),
          int(SWO_cutoff_high), int(SWO_cutoff_low), float(np.float32(extra_raise)), float(np.float32(max_raise)),
          int(filter_size), burn.shape[0], burn.shape[1], _stream(burn))
    return _finish_inplace(DEM_burn_arr, burn)


def raise_small_streams(inDEM_array, DEM_burn_arr, water_mask_array, nodata, extra_raise=0.50, max_raise=5.0, fsize=51):
    """tools/water.py:960-1007"""
    import torch
    same = inDEM_array is DEM_burn_arr
    burn = _dev(DEM_burn_arr, torch.float32, "DEM_burn_arr")
    dem = burn if same else _dev(inDEM_array, torch.float32, "raise_small_streams")
    wm = _dev(water_mask_array, torch.uint8, "water_mask_array")
    _call("hc_raise_small_streams_dev", burn.device, _tp(dem), _tp(burn), _tp(wm), float(nodata),
          float(np.float32(extra_raise)), float(np.float32(max_raise)), int(fsize), burn.shape[0], burn.shape[1],
          _stream(burn))
    return _finish_inplace(DEM_burn_arr, burn)


def smooth_rivers(inDEM_array, DEM_burn_arr, water_mask_array, SWO_array, nodata, SWO_cut=90, max_lower=10.0,
                  filter_size=21):
    """tools/water.py:928-957"""
    import torch
    same = inDEM_array is DEM_burn_arr
    burn = _dev(DEM_burn_arr, torch.float32, "DEM_burn_arr")
    dem = burn if same else _dev(inDEM_array, torch.float32, "smooth_rivers")
    wm = _dev(water_mask_array, torch.uint8, "water_mask_array")
    swo = _dev(SWO_array, torch.int16, "SWO_array")
    _call("hc_smooth_rivers_dev", burn.device, _tp(dem), _tp(burn), _tp(wm), _tp(swo), float(nodata), int(SWO_cut),
          float(np.float32(max_lower)), int(filter_size), burn.shape[0], burn.shape[1], _stream(burn))
    return _finish_inplace(DEM_burn_arr, burn)


def halo_rows(name, **kw):
    """Rows of halo a row band needs for `name` to give the whole-raster result on its own rows (the dependency
    radius of the chain: mask dilations + 5x5 mean + moving window), for bands.apply_banded / apply_rank."""
    if name == "river_minimize":
        return 1 + int(kw.get("river_filter_size", 3)) // 2
    if name == "raise_small_rivers":
        return int(kw.get("filter_size", 51)) // 2
    if name == "raise_small_streams":
        return 2 + int(kw.get("fsize", 51)) // 2
    if name == "smooth_rivers":
        return 2 + int(kw.get("filter_size", 21)) // 2
    raise KeyError(name)
