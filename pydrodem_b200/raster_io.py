"""Minimal single-band GeoTIFF / .npy raster I/O for the file-path boundary.

pydrodem hands rasters to WhiteboxTools / TauDEM as GeoTIFF paths written by `pyct.npy2tif`
(tools/convert.py:128-169: one band, `SetNoDataValue`, geotransform + projection copied from a
stencil).  GDAL is not available in this environment, so this module reads and writes the subset
that protocol needs: classic or BigTIFF, little endian, one sample per pixel, strips (or tiles on read),
uncompressed, LZW (what `pyct.tifcompress` / `tifdowncast` write, tools/convert.py:202,312; host codec in
csrc/tiffcodec.cu) or deflate, predictors 1-3 on read, float32/float64/int16/uint8/int32/uint32, the GDAL_NODATA tag (42113) and the
GeoTIFF georeferencing tags (33550 ModelPixelScale, 33922 ModelTiepoint, 34735-34737 GeoKeys), which
are carried through verbatim.  Anything else raises.  `.npy` paths are accepted everywhere with a
`<path>.json` side-car for nodata / georeferencing.
"""
import json
import os
import struct
import threading

import numpy as np

_TYPES = {1: ("B", 1), 2: ("c", 1), 3: ("H", 2), 4: ("I", 4), 5: ("II", 8), 6: ("b", 1), 8: ("h", 2), 9: ("i", 4),
          11: ("f", 4), 12: ("d", 8), 16: ("Q", 8), 17: ("q", 8), 18: ("Q", 8)}
_GEO_TAGS = (33550, 33922, 34735, 34736, 34737)
_SF = {1: {8: np.uint8, 16: np.uint16, 32: np.uint32}, 2: {8: np.int8, 16: np.int16, 32: np.int32},
       3: {32: np.float32, 64: np.float64}}


class RasterMeta(dict):
    """nodata (float or None), geo (dict tag -> (type, values)), dx, dy convenience."""

    @property
    def nodata(self):
        return self.get("nodata")

    @property
    def pixel_size(self):
        g = self.get("geo", {}).get(33550)
        if g:
            return float(g[1][0]), float(g[1][1])
        return 1.0, 1.0


def _read_ifd(f, big, off):
    f.seek(off)
    n = struct.unpack("<Q" if big else "<H", f.read(8 if big else 2))[0]
    tags = {}
    for _ in range(n):
        if big:
            tag, typ, cnt = struct.unpack("<HHQ", f.read(12))
            raw = f.read(8)
        else:
            tag, typ, cnt = struct.unpack("<HHI", f.read(8))
            raw = f.read(4)
        tags[tag] = (typ, cnt, raw)
    return tags


def _tag_values(f, big, entry):
    typ, cnt, raw = entry
    fmt, size = _TYPES[typ]
    total = size * cnt
    if total <= (8 if big else 4):
        data = raw[:total]
    else:
        off = struct.unpack("<Q" if big else "<I", raw)[0]
        pos = f.tell()
        f.seek(off)
        data = f.read(total)
        f.seek(pos)
    if typ == 2:
        return data.rstrip(b"\x00").decode("ascii", "replace")
    if typ == 5:
        v = struct.unpack("<" + "II" * cnt, data)
        return [v[2 * i] / max(v[2 * i + 1], 1) for i in range(cnt)]
    return list(struct.unpack("<" + fmt * cnt, data))


def read_raster(path):
    hit = _WB.snapshot_of(path) if _WB.depth > 0 else None
    if hit is not None:
        return hit
    _WB.wait_for(path)
    return _read_raster_now(path)


def _read_raster_now(path):
    """-> (2-D numpy array, RasterMeta)."""
    if path.endswith(".npy"):
        arr = np.load(path)
        meta = RasterMeta()
        if os.path.exists(path + ".json"):
            with open(path + ".json") as fh:
                meta.update(json.load(fh))
            meta["geo"] = {int(k): tuple(v) for k, v in meta.get("geo", {}).items()}
        return arr, meta
    with open(path, "rb") as f:
        head = f.read(16)
        if head[:2] != b"II":
            raise ValueError(f"{path}: only little-endian TIFF is supported")
        magic = struct.unpack("<H", head[2:4])[0]
        big = magic == 43
        if magic not in (42, 43):
            raise ValueError(f"{path}: not a TIFF")
        off = struct.unpack("<Q", head[8:16])[0] if big else struct.unpack("<I", head[4:8])[0]
        tags = _read_ifd(f, big, off)
        val = lambda t, d=None: _tag_values(f, big, tags[t]) if t in tags else d  # noqa: E731
        nx, ny = val(256)[0], val(257)[0]
        bits, spp = val(258, [1])[0], val(277, [1])[0]
        comp, fmtc = val(259, [1])[0], val(339, [1])[0]
        pred = val(317, [1])[0]
        if spp != 1 or comp not in (1, 5, 8, 32946) or pred not in (1, 2, 3):
            raise ValueError(f"{path}: need 1 band, uncompressed / LZW / deflate (got bands={spp}, compression={comp}, "
                             f"predictor={pred})")
        dtype = np.dtype(_SF[fmtc][bits])
        arr = np.empty((ny, nx), dtype)

        def block(off, nbytes, rows, cols):
            """one strip / tile: read, decompress, undo the predictor -> [rows, cols] array"""
            f.seek(off)
            raw = f.read(nbytes)
            want = rows * cols * dtype.itemsize
            if comp != 1:
                raw = _decompress(raw, comp, want)
            if len(raw) < want:
                raise ValueError(f"{path}: short strip / tile ({len(raw)} of {want} bytes)")
            return _unpredict(np.frombuffer(raw, np.uint8, want), pred, rows, cols, dtype)

        if 322 in tags:  # tiled
            tw, th = val(322)[0], val(323)[0]
            offs, cnts = val(324), val(325)
            tx = (nx + tw - 1) // tw
            for i, (o, c) in enumerate(zip(offs, cnts)):
                t = block(o, c if comp != 1 else tw * th * dtype.itemsize, th, tw)
                r0, c0 = (i // tx) * th, (i % tx) * tw
                arr[r0:r0 + th, c0:c0 + tw] = t[:ny - r0, :nx - c0]
        else:
            rps = min(val(278, [ny])[0], ny)
            offs = val(273)
            cnts = val(279, [None] * len(offs))
            for i, (o, c) in enumerate(zip(offs, cnts)):
                r0 = i * rps
                nr = min(rps, ny - r0)
                arr[r0:r0 + nr] = block(o, c if (comp != 1 and c is not None) else nr * nx * dtype.itemsize, nr, nx)
        meta = RasterMeta()
        nd = val(42113)
        if nd is not None:
            try:
                meta["nodata"] = float(nd)
            except ValueError:
                meta["nodata"] = None
        meta["geo"] = {t: (tags[t][0], val(t)) for t in _GEO_TAGS if t in tags}
        return arr, meta


def _decompress(raw, comp, want):
    """TIFF compression 5 (LZW, libhydrocore's host codec) or 8 / 32946 (deflate, zlib)"""
    if comp in (8, 32946):
        import zlib
        return zlib.decompress(raw)
    import ctypes
    from . import _lib
    src = np.frombuffer(raw, np.uint8)
    dst = np.empty(max(want, 1), np.uint8)
    n = ctypes.c_int64()
    _lib.check(_lib.load().hc_lzw_decode_host(ctypes.c_void_p(src.ctypes.data), len(raw), ctypes.c_void_p(dst.ctypes.data),
                                             want, ctypes.byref(n)), "hc_lzw_decode_host")
    return dst[:n.value].tobytes()


def _lzw(raw):
    import ctypes
    from . import _lib
    src = np.frombuffer(raw, np.uint8)
    dst = np.empty(2 * len(raw) + 16, np.uint8)
    n = ctypes.c_int64()
    _lib.check(_lib.load().hc_lzw_encode_host(ctypes.c_void_p(src.ctypes.data), len(raw), ctypes.c_void_p(dst.ctypes.data),
                                             len(dst), ctypes.byref(n)), "hc_lzw_encode_host")
    return dst[:n.value].tobytes()


def _unpredict(b, pred, rows, cols, dtype):
    """undo TIFF predictor 2 (horizontal differencing of the sample words) or 3 (floating point: bytes of a row
    regrouped most-significant plane first, then byte-differenced) on the bytes of one strip / tile"""
    bps = dtype.itemsize
    if pred == 1:
        return b.view(dtype).reshape(rows, cols)
    if pred == 2:
        u = b.view(np.dtype(f"<u{bps}")).reshape(rows, cols)
        return np.cumsum(u, axis=1, dtype=u.dtype).view(dtype)
    cum = np.cumsum(b.reshape(rows, cols * bps), axis=1, dtype=np.uint8)
    planes = cum.reshape(rows, bps, cols)[:, ::-1, :]              # little-endian byte order
    return np.ascontiguousarray(planes.transpose(0, 2, 1)).view(dtype).reshape(rows, cols)


class _WriteBehind:
    """Write-behind for a unit pipeline: inside `with write_behind():` a write_raster() call snapshots the array, creates
    the (empty) file at once -- os.path.exists() is true from then on -- and hands the encoding + file write to a small
    pool of background threads; read_raster() of a path whose write is still queued is served from the snapshot (a copy, with the
    metadata a read of the finished file would give); leaving the block waits for all writes and re-raises the first
    failure.  The pipeline's own kernels and numpy work overlap the ~50-120 ms a 3601^2 raster takes
    to reach the page cache.  Off by default: outside the block every write is synchronous."""

    def __init__(self):
        self.depth = 0
        self.pool = None
        self.pending = {}
        self.lock = threading.Lock()

    def __enter__(self):
        with self.lock:
            if self.depth == 0:
                from concurrent.futures import ThreadPoolExecutor
                self.pool = ThreadPoolExecutor(max_workers=4)
            self.depth += 1
        return self

    def __exit__(self, *exc):
        with self.lock:
            self.depth -= 1
            last = self.depth == 0
        if last:
            self.flush()
            self.pool.shutdown(wait=True)
            self.pool = None
        return False

    def flush(self):
        with self.lock:
            futs = [e[0] for e in self.pending.values()]
            self.pending.clear()
        err = None
        for f in futs:
            try:
                f.result()
            except Exception as e:  # noqa: BLE001  (re-raised below)
                err = err or e
        if err is not None:
            raise err

    def wait_for(self, path):
        with self.lock:
            e = self.pending.pop(os.path.abspath(path), None)
        if e is not None:
            e[0].result()

    def snapshot_of(self, path):
        """(array copy, RasterMeta) of a queued write, or None"""
        with self.lock:
            e = self.pending.get(os.path.abspath(path))
        if e is None:
            return None
        if e[0].done():
            e[0].result()          # surfaces a failed write here rather than at the end of the block
        return e[1].copy(), RasterMeta({"nodata": e[2].get("nodata"), "geo": dict(e[2].get("geo", {}))})


_WB = _WriteBehind()


def write_behind():
    """context manager: see _WriteBehind"""
    return _WB


def write_raster(path, arr, meta=None, nodata=None, compress=None, tiled=False):
    """Write a 2-D array as a single-band GeoTIFF (BigTIFF above 3.9 GB): uncompressed strips by default,
    `compress='lzw'` (what pyct.tifcompress produces, tools/convert.py:312) or `'deflate'`; `tiled=True` lays the
    raster out in 256 x 256 tiles (GDAL's TILED=YES block size; `["COMPRESS=LZW", "TILED=YES"]`, tools/convert.py:202).
    Inside `with write_behind():` the write is queued (see _WriteBehind)."""
    if _WB.depth > 0 and _WB.pool is not None:
        snap = np.array(arr, copy=True, order="C")
        if snap.ndim != 2:
            raise ValueError("write_raster needs a 2-D array")
        _WB.wait_for(path)                       # an earlier queued write of the same path goes first
        open(path, "wb").close()
        m = RasterMeta(meta or {})
        seen = {"nodata": float(nodata) if nodata is not None else m.get("nodata"),
                "geo": {int(t): (int(v[0]), tuple(v[1]) if not isinstance(v[1], str) else v[1])
                        for t, v in m.get("geo", {}).items() if int(t) in _GEO_TAGS}}
        if snap.dtype.byteorder == ">":
            snap = snap.astype(snap.dtype.newbyteorder("<"))
        fut = _WB.pool.submit(_write_raster_now, path, snap, meta, nodata, compress, tiled)
        with _WB.lock:
            _WB.pending[os.path.abspath(path)] = (fut, snap, seen)
        return
    _write_raster_now(path, arr, meta, nodata, compress, tiled)


def _write_raster_now(path, arr, meta=None, nodata=None, compress=None, tiled=False):
    arr = np.ascontiguousarray(arr)
    if arr.ndim != 2:
        raise ValueError("write_raster needs a 2-D array")
    if compress not in (None, 'lzw', 'deflate'):
        raise ValueError("compress must be None, 'lzw' or 'deflate'")
    meta = RasterMeta(meta or {})
    if nodata is not None:
        meta["nodata"] = float(nodata)
    if path.endswith(".npy"):
        np.save(path, arr)
        side = {"nodata": meta.get("nodata"), "geo": {str(k): list(v) for k, v in meta.get("geo", {}).items()}}
        with open(path + ".json", "w") as fh:
            json.dump(side, fh)
        return
    kind = arr.dtype.kind
    fmtc = {"u": 1, "i": 2, "f": 3}.get(kind)
    if fmtc is None or arr.dtype.itemsize * 8 not in _SF[fmtc]:
        raise ValueError(f"unsupported dtype {arr.dtype}")
    if arr.dtype.byteorder == ">":
        arr = arr.astype(arr.dtype.newbyteorder("<"))
    ny, nx = arr.shape
    row_bytes = nx * arr.dtype.itemsize
    import zlib
    pack = {None: (lambda b: b), 'lzw': _lzw, 'deflate': (lambda b: zlib.compress(b, 6))}[compress]
    rps = max(1, min(ny, ((8 << 20) if compress is None else (1 << 20)) // max(row_bytes, 1)))
    blobs = None
    if tiled:
        tw = th = 256
        blobs = []
        for r0 in range(0, ny, th):
            for c0 in range(0, nx, tw):
                t = np.zeros((th, tw), arr.dtype)
                blk = arr[r0:r0 + th, c0:c0 + tw]
                t[:blk.shape[0], :blk.shape[1]] = blk
                blobs.append(pack(t.tobytes()))
        cnts = [len(b) for b in blobs]
    elif compress is not None:
        nstrips = (ny + rps - 1) // rps
        blobs = [pack(arr[i * rps:min(ny, (i + 1) * rps)].tobytes()) for i in range(nstrips)]
        cnts = [len(b) for b in blobs]
    else:
        nstrips = (ny + rps - 1) // rps
        cnts = [min(rps, ny - i * rps) * row_bytes for i in range(nstrips)]
    data_bytes = sum(cnts)
    big = data_bytes > 3_900_000_000
    head = 16 if big else 8
    offs = [int(o) for o in head + np.concatenate([[0], np.cumsum(cnts[:-1], dtype=np.int64)])] if cnts else []
    LONG = 16 if big else 4
    entries = [(256, 4, [nx]), (257, 4, [ny]), (258, 3, [arr.dtype.itemsize * 8]),
               (259, 3, [{None: 1, 'lzw': 5, 'deflate': 8}[compress]]), (262, 3, [1]), (277, 3, [1]), (284, 3, [1]),
               (339, 3, [fmtc])]
    if tiled:
        entries += [(322, 4, [tw]), (323, 4, [th]), (324, LONG, offs), (325, LONG, cnts)]
    else:
        entries += [(273, LONG, offs), (278, 4, [rps]), (279, LONG, cnts)]
    for t, (typ, vals) in sorted(meta.get("geo", {}).items()):
        entries.append((int(t), int(typ), vals))
    if meta.get("nodata") is not None:
        nd = meta["nodata"]
        entries.append((42113, 2, (repr(int(nd)) if float(nd).is_integer() else repr(float(nd)))))
    entries.sort(key=lambda e: e[0])
    ifd_off = head + data_bytes
    ifd_off += ifd_off & 1
    esize = 20 if big else 12
    ifd_len = (8 if big else 2) + esize * len(entries) + (8 if big else 4)
    extra_off = ifd_off + ifd_len
    ifd, extra = bytearray(), bytearray()
    ifd += struct.pack("<Q" if big else "<H", len(entries))
    inline = 8 if big else 4
    for tag, typ, vals in entries:
        if typ == 2:
            data = vals.encode("ascii") + b"\x00"
            cnt = len(data)
        else:
            fmt, _ = _TYPES[typ]
            data = struct.pack("<" + fmt * len(vals), *vals)
            cnt = len(vals)
        ifd += struct.pack("<HHQ" if big else "<HHI", tag, typ, cnt)
        if len(data) <= inline:
            ifd += data.ljust(inline, b"\x00")
        else:
            if len(extra) & 1:
                extra += b"\x00"
            ifd += struct.pack("<Q" if big else "<I", extra_off + len(extra))
            extra += data
    ifd += struct.pack("<Q" if big else "<I", 0)
    with open(path, "wb") as f:
        if big:
            f.write(b"II" + struct.pack("<HHHQ", 43, 8, 0, ifd_off))
        else:
            f.write(b"II" + struct.pack("<HI", 42, ifd_off))
        if blobs is None:
            f.write(memoryview(arr).cast("B"))
        else:
            for b in blobs:
                f.write(b)
        if (head + data_bytes) & 1:
            f.write(b"\x00")
        f.write(ifd)
        f.write(extra)
