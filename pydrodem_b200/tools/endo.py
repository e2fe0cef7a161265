"""Raster logic of `endo.find_sink` (tools/endo.py:251-432) on the GPU.

The reference function is wrapped in GDAL reads, a GDAL cutline of the SWO mosaic and geopandas /
shapely output; what it computes on arrays is: 1st-percentile threshold -> 8-connected labels of the
low mask -> largest blob -> intersection with SWO >= 80 when non-empty -> the minimum-elevation pixel of
that area, or, on ties, the tied pixel nearest their centroid.  `find_sink_pixel` is that, with the
arrays as arguments; the percentile's two order statistics come from an exact radix select
(hc_kth_value_dev), the labels from hc_label_dev, the blob sizes from hc_seg_stats_dev.
"""
import ctypes

import numpy as np

from .. import _lib, core


def _percentile_from_order_stats(n, q, kth):
    """np.percentile(a, q) (method 'linear') for a float32 array of n elements, given kth(i) -> i-th smallest.
    Uses numpy's own index / gamma / lerp helpers so that every rounding step is numpy's."""
    try:
        from numpy.lib import _function_base_impl as f      # numpy >= 2
    except ImportError:                                      # pragma: no cover
        from numpy.lib import function_base as f
    quantiles = np.asanyarray(np.true_divide(q, np.float32(100)))
    props = f._QuantileMethods["linear"]
    virtual_indexes = np.asanyarray(props["get_virtual_index"](n, quantiles))
    stub = np.empty((n,), dtype=np.bool_) if n < (1 << 20) else np.lib.stride_tricks.as_strided(np.zeros(1, np.bool_), (n,), (0,))
    previous_indexes, next_indexes = f._get_indexes(stub, virtual_indexes, n)
    gamma = f._get_gamma(virtual_indexes, previous_indexes, props)
    prev = np.asanyarray(np.float32(kth(int(previous_indexes))))
    nxt = np.asanyarray(np.float32(kth(int(next_indexes))))
    return f._lerp(prev, nxt, np.asanyarray(gamma))


def kth_value(t, k):
    """k-th smallest (0-based) element of a float32 CUDA tensor (exact)."""
    import torch
    v = ctypes.c_float(0)
    x = t.reshape(-1).contiguous()
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hc_kth_value_dev(_lib.context(x.device.index or 0), ctypes.c_void_p(x.data_ptr()),
                                                x.numel(), int(k), ctypes.byref(v),
                                                ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                   "hc_kth_value_dev")
    return np.float32(v.value)


def find_sink_pixel(raster_array, SWO_array=None, geotransform=(0.0, 1.0, 0.0, 0.0, 0.0, -1.0)):
    """(row, col) of the sink pixel endo.find_sink would place (tools/endo.py:285-420).  `raster_array`: the
    basin DEM with nodata = +99999 (so it never wins a minimum, models/find_sinks.py:55); `SWO_array`: uint8
    surface-water occurrence on the same grid or None; `geotransform`: GDAL-style, only used for the
    centroid distance when several pixels tie for the minimum."""
    import torch
    a = torch.from_numpy(np.ascontiguousarray(raster_array, dtype=np.float32)).cuda() \
        if not core._is_torch(raster_array) else raster_array
    ny, nx = a.shape
    n = a.numel()
    low_threshold = _percentile_from_order_stats(n, 1, lambda i: kth_value(a, i))            # :286
    low = a <= float(low_threshold) if np.float32(low_threshold) == low_threshold else a.double() <= float(low_threshold)
    nlow = int(low.sum())
    if nlow == 1:                                                                           # :296-300
        idx = int(torch.nonzero(low.reshape(-1))[0])
        return idx // nx, idx % nx
    lab, nfeat = core.label(low.to(torch.uint8), 8)                                          # :320-323
    # size of every blob, largest one (first maximum, as np.argmax over count_list) :324-329
    lib = _lib.load()
    cnt = torch.empty(nfeat + 1, dtype=torch.int64, device=a.device)
    labf = lab.reshape(-1).contiguous()
    _lib.check(lib.hc_seg_stats_dev(_lib.context(a.device.index or 0), ctypes.c_void_p(labf.data_ptr()), None, None, n, nfeat, 1,
                                    None, None, None, ctypes.c_void_p(cnt.data_ptr()),
                                    ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)), "hc_seg_stats_dev")
    cnt[0] = -1
    biggest = int(torch.nonzero(cnt == cnt.max())[0])   # first maximum, like np.argmax
    area = lab == biggest
    if SWO_array is not None:                                                               # :332-368
        swo = torch.from_numpy(np.ascontiguousarray(SWO_array)).cuda() if not core._is_torch(SWO_array) else SWO_array
        both = area & (swo >= 80)
        if bool(both.any()):
            area = both
    zmin = a[area].min()                                                                    # :379
    ties = torch.nonzero(area & (a == zmin))                                                # :380-383
    if ties.shape[0] == 1:
        return int(ties[0, 0]), int(ties[0, 1])
    # several minima: the one nearest their centroid in map coordinates (:393-417, shapely nearest_points;
    # on equal distances the first pixel in raster order is taken here)
    x0, dx, _rx, y0, _ry, dy = geotransform                  # get_xy, tools/endo.py:213-216 (pixel centres)
    rows, cols = ties[:, 0].double(), ties[:, 1].double()
    xs = x0 + cols * dx + dx / 2.0
    ys = y0 + rows * dy + dy / 2.0
    d2 = (xs - xs.sum() / len(xs)) ** 2 + (ys - ys.sum() / len(ys)) ** 2
    j = int(torch.nonzero(d2 == d2.min())[0])
    return int(ties[j, 0]), int(ties[j, 1])


# ---- the scalar geometry around the sink pixel (tools/endo.py:158-250): host arithmetic on two numbers ----
def get_xy(r, c, gt):
    """centre coordinate of raster cell (r, c) for GDAL geotransform `gt` (tools/endo.py:197-215)"""
    x0, dx, rx, y0, ry, dy = gt
    return (x0 + c * dx + dx / 2.0, y0 + r * dy + dy / 2.0)


def coords_to_pixel(x, y, dx, dy, xmin, ymax):
    """(col, row) of a coordinate; dy is the geotransform's (negative) row step (tools/endo.py:218-250)"""
    col = int((x - xmin) / dx)
    row = int((y - ymax) / dy)
    return col, row


def set_nodata_points(point_xy, raster_array, dx, dy, xmin, ymax, nodata, set_buffer=False):
    """tools/endo.py:158-194 with the point given as (x, y) instead of a one-row GeoDataFrame: the sink pixel becomes
    nodata (in place for a numpy array or a CUDA tensor; `set_buffer` is unused upstream too)."""
    col, row = coords_to_pixel(point_xy[0], point_xy[1], dx, dy, xmin, ymax)
    raster_array[row][col] = nodata
    return raster_array
