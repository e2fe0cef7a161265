"""GeoTIFF subset reader/writer round trips (no GPU)."""
import numpy as np
import pytest


@pytest.mark.parametrize("dtype", [np.float32, np.float64, np.int16, np.uint8, np.int32, np.uint32])
@pytest.mark.parametrize("ext", [".tif", ".npy"])
def test_round_trip(tmp_path, dtype, ext):
    from pydrodem_b200 import raster_io
    rng = np.random.default_rng(0)
    a = (rng.random((37, 53)) * 100).astype(dtype)
    geo = {33550: (12, [30.0, 30.0, 0.0]), 33922: (12, [0.0, 0.0, 0.0, 5e5, 4.1e6, 0.0]),
           34735: (3, [1, 1, 0, 1, 1024, 0, 1, 1]), 34737: (2, "WGS 84 / UTM|")}
    p = str(tmp_path / ("r" + ext))
    raster_io.write_raster(p, a, {"geo": geo}, nodata=-9999)
    b, meta = raster_io.read_raster(p)
    assert b.dtype == a.dtype and np.array_equal(a, b)
    assert meta.nodata == -9999.0 and meta.pixel_size == (30.0, 30.0)
    assert meta["geo"][34735][1] == [1, 1, 0, 1, 1024, 0, 1, 1]


def test_large_strips_and_empty_meta(tmp_path):
    from pydrodem_b200 import raster_io
    a = np.arange(3000 * 1500, dtype=np.float32).reshape(3000, 1500)
    p = str(tmp_path / "big.tif")
    raster_io.write_raster(p, a)
    b, meta = raster_io.read_raster(p)
    assert np.array_equal(a, b) and meta.nodata is None and meta.pixel_size == (1.0, 1.0)


def test_reads_lzw_and_rejects_what_it_does_not_know(tmp_path):
    from PIL import Image
    from pydrodem_b200 import raster_io
    p = str(tmp_path / "lzw.tif")
    a = (np.arange(64, dtype=np.uint8).reshape(8, 8) // 3)
    Image.fromarray(a).save(p, compression="tiff_lzw")
    assert np.array_equal(raster_io.read_raster(p)[0], a)
    q = str(tmp_path / "packbits.tif")
    Image.fromarray(a).save(q, compression="packbits")
    with pytest.raises(ValueError):
        raster_io.read_raster(q)
    with pytest.raises(ValueError):
        raster_io.write_raster(str(tmp_path / "x.tif"), a, None, compress="jpeg")
