"""Mosaicking order and VRT composition (run_local.py:636-677, tools/convert.py:813-870) without GDAL."""
import os

import numpy as np
import pytest

from pydrodem_b200 import mosaic, raster_io


def _meta(x0, y0, s=0.5):
    return raster_io.RasterMeta({"geo": {33550: (12, [s, s, 0.0]), 33922: (12, [0.0, 0.0, 0.0, x0, y0, 0.0])}})


def test_mosaic_order_follows_the_reference(tmp_path):
    kinds = {"basin_1": "plain", "basin_2": "coast", "basin_3": "flat", "basin_4": "burn", "basin_5": "island",
             "basin_6": "plain", "basin_7": "flat", "basin_8": "coast"}
    dirs = []
    for name, kind in kinds.items():
        d = tmp_path / name
        d.mkdir()
        dirs.append(str(d))
        if kind == "coast":
            (d / "coast.txt").write_text("x")
        if kind == "island":
            (d / "island_chain.txt").write_text("x")
        if kind in ("flat", "burn"):
            m = np.zeros((4, 4), np.uint8)
            m[1, 1] = 1 if kind == "flat" else 2
            raster_io.write_raster(str(d / "Water_flatten_mask_overlap.tif"), m, None, nodata=0)
    got = [os.path.basename(p) for p in mosaic.mosaic_order(dirs)]
    # coastal bottom; plain (island chains stay in this group, as upstream); water-flattened; stream-burned on top
    assert got == ["basin_2", "basin_8", "basin_1", "basin_5", "basin_6", "basin_3", "basin_7", "basin_4"]


def test_vrt_paints_in_order_with_transparent_nodata(tmp_path):
    a = np.full((6, 8), 1.0, np.float32)
    b = np.full((6, 8), 2.0, np.float32)
    b[:, :3] = np.nan                                  # transparent where nan (srcNodata='nan')
    c = np.full((4, 4), 3.0, np.float32)
    pa, pb, pc = (str(tmp_path / f"{n}.tif") for n in "abc")
    raster_io.write_raster(pa, a, _meta(10.0, 50.0), nodata=-9999)
    raster_io.write_raster(pb, b, _meta(12.0, 49.0), nodata=-9999)      # 4 cells east, 2 cells south of a
    raster_io.write_raster(pc, c, _meta(9.0, 51.0), nodata=-9999)       # 2 cells west, 2 cells north of a
    vrt = str(tmp_path / "m.vrt")
    mosaic.build_vrt([pa, pb, pc], vrt, srcnodata='nan', VRTnodata=None, outputSRS="EPSG:4326")
    m, meta = mosaic.read_vrt(vrt)
    assert m.shape == (10, 14) and meta.pixel_size == (0.5, 0.5)
    want = np.zeros((10, 14), np.float32)
    want[2:8, 2:10] = 1.0
    want[4:10, 6 + 3:14] = 2.0                          # b's first three columns stay transparent
    want[0:4, 0:4] = 3.0                                # c is last: on top
    assert np.array_equal(m, want)
    # reversed order: a ends on top of b and c where they overlap
    mosaic.build_vrt([pc, pb, pa], vrt, srcnodata='nan')
    m2, _ = mosaic.read_vrt(vrt)
    assert m2[2, 2] == 1.0 and m2[5, 9] == 1.0 and m2[9, 13] == 2.0
    txt = open(vrt).read()
    assert "<ComplexSource>" in txt and 'relativeToVRT="1"' in txt and "<NODATA>nan</NODATA>" in txt
    # without source nodata: SimpleSource, nan is painted like any value
    mosaic.build_vrt([pa, pb], vrt)
    m3, _ = mosaic.read_vrt(vrt)
    assert "<SimpleSource>" in open(vrt).read() and np.isnan(m3[4, 6])


def test_vrt_refuses_what_needs_gdal(tmp_path):
    a = np.ones((4, 4), np.float32)
    pa, pb = str(tmp_path / "a.tif"), str(tmp_path / "b.tif")
    raster_io.write_raster(pa, a, _meta(0.0, 10.0, 0.5))
    raster_io.write_raster(pb, a, _meta(0.25, 10.0, 0.5))               # half a pixel off the grid
    with pytest.raises(NotImplementedError):
        mosaic.build_vrt([pa, pb], str(tmp_path / "m.vrt"))
    raster_io.write_raster(pb, a, _meta(0.0, 10.0, 0.25))               # other resolution
    with pytest.raises(NotImplementedError):
        mosaic.build_vrt([pa, pb], str(tmp_path / "m.vrt"))
    with pytest.raises(NotImplementedError):
        mosaic.build_vrt([pa], str(tmp_path / "m.vrt"), outputSRS="EPSG:3857")
