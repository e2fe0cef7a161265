"""GeoTIFF wire format: LZW / deflate strips and tiles, predictors, checked against libtiff (through
Pillow, which is installed) in both directions, plus the convert.py helpers."""
import ctypes
import os
import struct

import numpy as np
import pytest

from pydrodem_b200 import _lib, raster_io
from pydrodem_b200.tools import convert

PIL = pytest.importorskip("PIL.Image")


def _enc(b):
    lib = _lib.load()
    src = np.frombuffer(b, np.uint8)
    dst = np.empty(2 * len(b) + 16, np.uint8)
    n = ctypes.c_int64()
    assert lib.hc_lzw_encode_host(src.ctypes.data if len(b) else None, len(b), dst.ctypes.data, len(dst), ctypes.byref(n)) == 0
    return dst[:n.value].tobytes()


def _dec(b, nout):
    lib = _lib.load()
    src = np.frombuffer(b, np.uint8)
    dst = np.empty(max(nout, 1), np.uint8)
    n = ctypes.c_int64()
    assert lib.hc_lzw_decode_host(src.ctypes.data, len(b), dst.ctypes.data, nout, ctypes.byref(n)) == 0
    return dst[:n.value].tobytes()


def scene():
    rng = np.random.default_rng(1)
    a = (rng.random((301, 257)) * 100).astype(np.float32)
    a[50:190, :] = -9999
    return a


def test_lzw_round_trip_including_table_resets():
    rng = np.random.default_rng(0)
    for n in (0, 1, 2, 100, 5000, 70000, 300000):      # > 4094 codes: the table is cleared and rebuilt several times
        for b in (rng.integers(0, 256, n, dtype=np.uint8).tobytes(),
                  np.repeat(rng.integers(0, 4, max(n // 50, 1), dtype=np.uint8), 50)[:n].tobytes(),
                  (b"the quick brown fox " * (n // 20 + 1))[:n]):
            e = _enc(b)
            assert _dec(e, len(b)) == b
    assert len(_enc(b"\x00" * 100000)) < 2000
    lib = _lib.load()
    bad = np.frombuffer(b"\x80\x7f\xff\xff\xff", np.uint8)     # clear, then a code far beyond the table
    out = np.empty(64, np.uint8)
    assert lib.hc_lzw_decode_host(bad.ctypes.data, len(bad), out.ctypes.data, 64, None) != 0


@pytest.mark.parametrize("comp", [None, "lzw", "deflate"])
def test_written_files_are_read_by_libtiff_and_by_us(tmp_path, comp):
    a = scene()
    meta = raster_io.RasterMeta({"geo": {33550: (12, [30.0, 30.0, 0.0]), 33922: (12, [0.0, 0.0, 0.0, 5e5, 4.1e6, 0.0])}})
    for arr in (a, (a * 3).astype(np.int16), (a > 50).astype(np.uint8), a.astype(np.float64), (a * 7).astype(np.int32)):
        p = str(tmp_path / f"t_{arr.dtype}.tif")
        raster_io.write_raster(p, arr, meta, nodata=-9999, compress=comp)
        b, m = raster_io.read_raster(p)
        assert np.array_equal(b, arr) and b.dtype == arr.dtype and m.nodata == -9999 and m.pixel_size == (30.0, 30.0)
        if arr.dtype in (np.float32, np.uint8):
            assert np.array_equal(np.array(PIL.open(p)), arr)           # libtiff decodes our stream
    if comp:
        assert os.path.getsize(str(tmp_path / "t_float32.tif")) < 0.8 * a.nbytes


def test_libtiff_written_files_are_read(tmp_path):
    a = scene()
    u = ((a > 50) * 200).astype(np.uint8)
    i16 = (a * 3).astype(np.uint16)
    PIL.fromarray(a).save(str(tmp_path / "p1.tif"), compression="tiff_lzw")
    PIL.fromarray(a).save(str(tmp_path / "p2.tif"), compression="tiff_adobe_deflate")
    PIL.fromarray(u).save(str(tmp_path / "p3.tif"), compression="tiff_lzw", tiffinfo={317: 2})
    PIL.fromarray(i16).save(str(tmp_path / "p4.tif"), compression="tiff_lzw", tiffinfo={317: 2})
    for f, w in (("p1", a), ("p2", a), ("p3", u), ("p4", i16)):
        b, _ = raster_io.read_raster(str(tmp_path / f"{f}.tif"))
        assert b.dtype == w.dtype and np.array_equal(b, w), f


def _tiled_tiff(path, arr, tw, th, comp, pred):
    """a classic little-endian tiled TIFF built by hand (GDAL's COMPRESS=LZW TILED=YES layout, tools/convert.py:202)"""
    ny, nx = arr.shape
    bps = arr.dtype.itemsize
    tiles = []
    for r0 in range(0, ny, th):
        for c0 in range(0, nx, tw):
            t = np.zeros((th, tw), arr.dtype)
            blk = arr[r0:r0 + th, c0:c0 + tw]
            t[:blk.shape[0], :blk.shape[1]] = blk
            if pred == 2:
                u = t.view(np.dtype(f"<u{bps}")).copy()
                u[:, 1:] = u[:, 1:] - u[:, :-1]
                raw = u.tobytes()
            elif pred == 3:
                planes = t.view(np.uint8).reshape(th, tw, bps).transpose(0, 2, 1)[:, ::-1, :].reshape(th, tw * bps).copy()
                planes[:, 1:] = planes[:, 1:] - planes[:, :-1]
                raw = planes.tobytes()
            else:
                raw = t.tobytes()
            tiles.append(_enc(raw) if comp == 5 else raw)
    offs, pos = [], 8
    for t in tiles:
        offs.append(pos)
        pos += len(t) + (len(t) & 1)
    fmt = {"f": 3, "u": 1, "i": 2}[arr.dtype.kind]
    ents = [(256, 4, [nx]), (257, 4, [ny]), (258, 3, [bps * 8]), (259, 3, [comp]), (262, 3, [1]), (277, 3, [1]),
            (317, 3, [pred]), (322, 4, [tw]), (323, 4, [th]), (324, 4, offs), (325, 4, [len(t) for t in tiles]),
            (339, 3, [fmt])]
    ifd_off = pos
    extra_off = ifd_off + 2 + 12 * len(ents) + 4
    ifd, extra = struct.pack("<H", len(ents)), b""
    for tag, typ, vals in ents:
        f = {3: "H", 4: "I"}[typ]
        data = struct.pack("<" + f * len(vals), *vals)
        ifd += struct.pack("<HHI", tag, typ, len(vals))
        if len(data) <= 4:
            ifd += data.ljust(4, b"\x00")
        else:
            ifd += struct.pack("<I", extra_off + len(extra))
            extra += data + (b"\x00" if len(data) & 1 else b"")
    ifd += struct.pack("<I", 0)
    with open(path, "wb") as fh:
        fh.write(b"II" + struct.pack("<HI", 42, ifd_off))
        for t in tiles:
            fh.write(t + (b"\x00" if len(t) & 1 else b""))
        fh.write(ifd + extra)


@pytest.mark.parametrize("comp,pred", [(1, 1), (5, 1), (5, 2), (5, 3)])
def test_tiled_lzw_with_predictors(tmp_path, comp, pred):
    a = scene()[:150, :200]
    for arr in ((a, (a * 3).astype(np.int16)) if pred != 3 else (a, a.astype(np.float64))):
        p = str(tmp_path / f"tile_{comp}_{pred}_{arr.dtype}.tif")
        _tiled_tiff(p, arr, 64, 32, comp, pred)
        b, _ = raster_io.read_raster(p)
        assert b.dtype == arr.dtype and np.array_equal(b, arr)
        if arr.dtype == np.float32 and pred != 3:
            assert np.array_equal(np.array(PIL.open(p)), arr)       # the hand-built file is a valid TIFF for libtiff too


def test_convert_helpers(tmp_path):
    a = scene()
    stencil = str(tmp_path / "stencil.tif")
    meta = raster_io.RasterMeta({"geo": {33550: (12, [0.5, 0.5, 0.0]), 33922: (12, [0.0, 0.0, 0.0, 10.0, 20.0, 0.0])}})
    raster_io.write_raster(stencil, a, meta, nodata=-9999)
    out = str(tmp_path / "out.tif")
    convert.npy2tif(a.astype(np.float64) * 2, stencil, out, nodata=-9999, dtype=7)
    b, m = raster_io.read_raster(out)
    assert b.dtype == np.float64 and m.pixel_size == (0.5, 0.5) and m.nodata == -9999
    convert.tifdowncast(out)
    b, m = raster_io.read_raster(out)
    assert b.dtype == np.float32 and np.array_equal(b, (a.astype(np.float64) * 2).astype(np.float32)) and m.nodata == -9999
    assert not os.path.exists(str(tmp_path / "out_dc.tif"))
    size = os.path.getsize(out)
    convert.tifcompress(out)
    b2, m2 = raster_io.read_raster(out)
    assert np.array_equal(b2, b) and m2.nodata == -9999 and os.path.getsize(out) < size
    assert np.array_equal(np.array(PIL.open(out)), b)
    convert.npy2tif(a > 0, stencil, str(tmp_path / "mask.tif"), nodata=0, dtype=1)
    assert raster_io.read_raster(str(tmp_path / "mask.tif"))[0].dtype == np.uint8


@pytest.mark.parametrize("comp", [None, "lzw", "deflate"])
def test_tiled_writing_is_read_by_libtiff_and_by_us(tmp_path, comp):
    """`["COMPRESS=LZW", "TILED=YES"]` (tools/convert.py:202): 256 x 256 tiles, edge tiles padded"""
    rng = np.random.default_rng(2)
    a = (rng.random((600, 515)) * 100).astype(np.float32)
    a[100:300] = 7
    for arr in (a, (a > 50).astype(np.uint8)):
        p = str(tmp_path / f"tiled_{arr.dtype}.tif")
        raster_io.write_raster(p, arr, None, nodata=-9999, compress=comp, tiled=True)
        b, m = raster_io.read_raster(p)
        assert np.array_equal(b, arr) and m.nodata == -9999
        assert np.array_equal(np.array(PIL.open(p)), arr)
