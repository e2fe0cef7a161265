"""The three raster-file helpers of `tools/convert.py` the depression path leans on, over pydrodem_b200.raster_io
instead of GDAL: `npy2tif` (:128-169), `tifdowncast` (:211-271), `tifcompress` (:274-320).  Same arguments; a GDAL
data-type code (gdal.GDT_Byte = 1 ... GDT_Float64 = 7) or a numpy dtype is accepted for `dtype`.  Projection /
geotransform are the GeoTIFF tags of the stencil file, carried through verbatim."""
import os

import numpy as np

from .. import raster_io

_GDT = {1: np.uint8, 2: np.uint16, 3: np.int16, 4: np.uint32, 5: np.int32, 6: np.float32, 7: np.float64}


def _np_dtype(dtype):
    if isinstance(dtype, int) and not isinstance(dtype, bool):
        return np.dtype(_GDT[dtype])
    return np.dtype(dtype)


def npy2tif(array, orig_tif, outfile, nodata=None, dtype=6, outformat="GTiff"):
    """array -> single-band GeoTIFF with the georeferencing of `orig_tif` (tools/convert.py:128-169)"""
    if outformat != "GTiff":
        raise NotImplementedError("only GTiff output is on the reference path")
    _, meta = raster_io.read_raster(orig_tif)
    meta = raster_io.RasterMeta({"geo": meta.get("geo", {})})
    raster_io.write_raster(outfile, np.asarray(array).astype(_np_dtype(dtype)), meta, nodata=nodata)


def tifdowncast(input_tif, dtype=6, dstSRS=None, nodataval=None):
    """rewrite a (float64) GeoTIFF in place at a narrower type, keeping nodata and georeferencing
    (tools/convert.py:211-271; WhiteboxTools writes float64, the GPU tools write float32 already)"""
    if dstSRS is not None:
        raise NotImplementedError("re-projection (dstSRS) needs GDAL and is not on the reference path")
    arr, meta = raster_io.read_raster(input_tif)
    nd = meta.nodata if nodataval is None else nodataval
    out = os.path.join(os.path.dirname(input_tif), os.path.splitext(os.path.basename(input_tif))[0] + "_dc.tif")
    raster_io.write_raster(out, arr.astype(_np_dtype(dtype)), meta, nodata=nd)
    os.remove(input_tif)
    os.rename(out, input_tif)


def tifcompress(input_tif, dstSRS=None):
    """rewrite a GeoTIFF in place with COMPRESS=LZW (tools/convert.py:274-320)"""
    if dstSRS is not None:
        raise NotImplementedError("re-projection (dstSRS) needs GDAL and is not on the reference path")
    arr, meta = raster_io.read_raster(input_tif)
    out = os.path.join(os.path.dirname(input_tif), os.path.splitext(os.path.basename(input_tif))[0] + "_compressed.tif")
    raster_io.write_raster(out, arr, meta, nodata=meta.nodata, compress='lzw')
    os.remove(input_tif)
    os.rename(out, input_tif)
