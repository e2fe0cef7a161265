"""Mosaicking of unit outputs: the ORDER in which the reference stacks basin DEMs into its
conditioned mosaic (run_local.py:636-677) and the VRT that carries it (`pyct.build_vrt`, tools/convert.py:813-870 ->
gdal.BuildVRT), without GDAL.

* `mosaic_order(basin_dirs)`  -- coastal basins at the bottom, ordinary basins above them, then the water-flattened
  ones, then the stream-burned ones on top (a later source overwrites an earlier one where both have data).
* `build_vrt(tifs, vrt_path, srcnodata=, VRTnodata=)`  -- writes a GDAL-readable .vrt (one band, the sources as
  ComplexSource / SimpleSource rectangles in the given order) for north-up GeoTIFFs of one resolution on one pixel grid.
* `read_vrt(vrt_path)`  -- composes the mosaic array (painter's algorithm, source nodata transparent): what
  `gdal_array.LoadFile(vrt)` returns for such a file.  Host assembly of file windows, no arithmetic on elevations.
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import raster_io

_WKT_4326 = ('GEOGCS["WGS 84",DATUM["WGS_1984",SPHEROID["WGS 84",6378137,298.257223563,AUTHORITY["EPSG","7030"]],'
             'AUTHORITY["EPSG","6326"]],PRIMEM["Greenwich",0,AUTHORITY["EPSG","8901"]],UNIT["degree",0.0174532925199433,'
             'AUTHORITY["EPSG","9122"]],AXIS["Latitude",NORTH],AXIS["Longitude",EAST],AUTHORITY["EPSG","4326"]]')
_GDAL_TYPE = {"uint8": "Byte", "uint16": "UInt16", "int16": "Int16", "uint32": "UInt32", "int32": "Int32",
              "float32": "Float32", "float64": "Float64"}


def mosaic_order(basin_dirs, flatten_name='Water_flatten_mask_overlap.tif'):
    """run_local.py:636-671: returns the basin directories bottom-to-top.  Island chains are skipped, `coast.txt`
    marks coastal basins (bottom), a flatten mask with any cell == 1 marks water-flattened basins, a flatten mask
    without one marks stream-burned basins (top)."""
    fpaths, f2paths, coast_paths, skipped = [], [], [], []
    for bpath in basin_dirs:
        if os.path.exists(os.path.join(bpath, 'island_chain.txt')):
            skipped.append(bpath)
            continue
        if os.path.exists(os.path.join(bpath, 'coast.txt')):
            coast_paths.append(bpath)
            continue
        flatten_tif = os.path.join(bpath, flatten_name)
        if os.path.exists(flatten_tif):
            flatten_array, _ = raster_io.read_raster(flatten_tif)
            (fpaths if np.any(flatten_array == 1) else f2paths).append(bpath)
    special = set(coast_paths) | set(fpaths) | set(f2paths)
    # (the reference keeps island-chain directories in the plain group: they never enter a special list)
    plain = [p for p in basin_dirs if p not in special]
    return coast_paths + plain + fpaths + f2paths


def _geo(meta, path):
    g = meta.get("geo", {})
    if 33550 not in g or 33922 not in g:
        raise ValueError(f"{path}: no ModelPixelScale / ModelTiepoint georeferencing")
    sx, sy = float(g[33550][1][0]), float(g[33550][1][1])
    t = g[33922][1]
    x0 = float(t[3]) - float(t[0]) * sx
    y0 = float(t[4]) + float(t[1]) * sy
    return x0, y0, sx, sy


def build_vrt(tifs, vrt_path, srcnodata=None, VRTnodata=None, resolution='highest', xres=None, yres=None,
              alignedPixels=True, resampleAlg=None, outputSRS=None):
    """tools/convert.py:813-870.  All sources must share pixel size and lie on one pixel grid (the basin / tile rasters
    of one project do: they are cut from one reference grid); no resampling."""
    if resampleAlg is not None or xres is not None or yres is not None:
        raise NotImplementedError("resampling needs GDAL; the project's rasters share one grid")
    src = []
    for p in tifs:
        arr, meta = raster_io.read_raster(p)
        x0, y0, sx, sy = _geo(meta, p)
        src.append(dict(path=p, ny=arr.shape[0], nx=arr.shape[1], x0=x0, y0=y0, sx=sx, sy=sy, dtype=str(arr.dtype),
                        nodata=meta.nodata))
    if not src:
        raise ValueError("build_vrt: no sources")
    sx, sy = src[0]["sx"], src[0]["sy"]
    for s_ in src:
        if abs(s_["sx"] - sx) > 1e-9 * abs(sx) or abs(s_["sy"] - sy) > 1e-9 * abs(sy):
            raise NotImplementedError(f"{s_['path']}: pixel size differs (resampling needs GDAL)")
    xmin = min(s_["x0"] for s_ in src)
    ymax = max(s_["y0"] for s_ in src)
    xmax = max(s_["x0"] + s_["nx"] * sx for s_ in src)
    ymin = min(s_["y0"] - s_["ny"] * sy for s_ in src)
    X, Y = int(round((xmax - xmin) / sx)), int(round((ymax - ymin) / sy))
    root = ET.Element("VRTDataset", rasterXSize=str(X), rasterYSize=str(Y))
    if outputSRS is not None:
        if str(outputSRS).upper() != "EPSG:4326":
            raise NotImplementedError("only outputSRS='EPSG:4326' (what the reference passes) can be written without GDAL")
        ET.SubElement(root, "SRS").text = _WKT_4326
    ET.SubElement(root, "GeoTransform").text = ", ".join(repr(float(v)) for v in (xmin, sx, 0.0, ymax, 0.0, -sy))
    band = ET.SubElement(root, "VRTRasterBand", dataType=_GDAL_TYPE[src[0]["dtype"]], band="1")
    if VRTnodata is not None:
        ET.SubElement(band, "NoDataValue").text = str(VRTnodata)
    for s_ in src:
        xoff = (s_["x0"] - xmin) / sx
        yoff = (ymax - s_["y0"]) / sy
        if abs(xoff - round(xoff)) > 1e-3 or abs(yoff - round(yoff)) > 1e-3:
            raise NotImplementedError(f"{s_['path']}: not on the mosaic's pixel grid (resampling needs GDAL)")
        e = ET.SubElement(band, "ComplexSource" if srcnodata is not None else "SimpleSource")
        rel = os.path.relpath(s_["path"], os.path.dirname(os.path.abspath(vrt_path)))
        ET.SubElement(e, "SourceFilename", relativeToVRT="1").text = rel
        ET.SubElement(e, "SourceBand").text = "1"
        ET.SubElement(e, "SourceProperties", RasterXSize=str(s_["nx"]), RasterYSize=str(s_["ny"]),
                      DataType=_GDAL_TYPE[s_["dtype"]], BlockXSize=str(s_["nx"]), BlockYSize="1")
        ET.SubElement(e, "SrcRect", xOff="0", yOff="0", xSize=str(s_["nx"]), ySize=str(s_["ny"]))
        ET.SubElement(e, "DstRect", xOff=str(int(round(xoff))), yOff=str(int(round(yoff))), xSize=str(s_["nx"]),
                      ySize=str(s_["ny"]))
        if srcnodata is not None:
            ET.SubElement(e, "NODATA").text = str(srcnodata)
    ET.indent(root)
    ET.ElementTree(root).write(vrt_path)


def read_vrt(vrt_path):
    """-> (mosaic array, RasterMeta).  Sources are painted in file order; cells equal to a source's NODATA (nan
    included) are transparent; cells no source covers hold the band's NoDataValue (0 when there is none, as GDAL does)."""
    root = ET.parse(vrt_path).getroot()
    X, Y = int(root.get("rasterXSize")), int(root.get("rasterYSize"))
    band = root.find("VRTRasterBand")
    inv = {v: k for k, v in _GDAL_TYPE.items()}
    dtype = np.dtype(inv[band.get("dataType")])
    nd = band.find("NoDataValue")
    fill = float(nd.text) if nd is not None else 0.0
    out = np.full((Y, X), fill if dtype.kind == "f" or not np.isnan(fill) else 0, dtype)
    base = os.path.dirname(os.path.abspath(vrt_path))
    for e in band:
        if e.tag not in ("SimpleSource", "ComplexSource"):
            continue
        fn = e.find("SourceFilename")
        p = os.path.join(base, fn.text) if fn.get("relativeToVRT") == "1" else fn.text
        arr, _ = raster_io.read_raster(p)
        s, d = e.find("SrcRect"), e.find("DstRect")
        sx0, sy0, sw, sh = (int(float(s.get(k))) for k in ("xOff", "yOff", "xSize", "ySize"))
        dx0, dy0, dw, dh = (int(float(d.get(k))) for k in ("xOff", "yOff", "xSize", "ySize"))
        if (sw, sh) != (dw, dh):
            raise NotImplementedError("scaled sources need GDAL")
        win = arr[sy0:sy0 + sh, sx0:sx0 + sw]
        dst = out[dy0:dy0 + dh, dx0:dx0 + dw]
        ndt = e.find("NODATA")
        if ndt is not None:
            v = float(ndt.text)
            keep = ~np.isnan(win) if np.isnan(v) else (win != v)
            dst[keep] = win[keep]
        else:
            dst[...] = win
    gt = [float(v) for v in root.find("GeoTransform").text.split(",")]
    meta = raster_io.RasterMeta({"nodata": (float(nd.text) if nd is not None else None),
                                 "geo": {33550: (12, [gt[1], -gt[5], 0.0]), 33922: (12, [0.0, 0.0, 0.0, gt[0], gt[3], 0.0])}})
    return out, meta
