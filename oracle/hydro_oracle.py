"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (ctypes front for oracle/hydro_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under pydrodem_b200/ does.

PARITY UNPINNED for fill / D8 / flats / accumulation (see the header of hydro_oracle.c and
DESIGN.md §oracle): the arithmetic of those passes lives in WhiteboxTools / TauDEM binaries
that are not in /root/reference and cannot be run here.  Labelling and the smoothing stencils
ARE pinned: they are scipy.ndimage.label / scipy.signal.convolve2d / np.gradient, the very
calls the reference makes (tools/depressions.py:1669, tools/derive.py:94,124,230-233).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "hydro_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libhydro_oracle.so")

SEED_WBT, SEED_TAUDEM = 0, 1
TABLE_WBT, TABLE_TAUDEM = 0, 1
DIR_NOFLOW, DIR_NODATA = 0, 255


def build(force=False):
    """gcc -O2 the C restatement into oracle/_build/ (git-ignored, travels with gpurun)."""
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    os.makedirs(_OUT_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-std=c99", "-fPIC", "-shared", "-o", _SO, _SRC, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2
    return a


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _chk(rc, name):
    if rc != 0:
        raise RuntimeError(f"oracle {name} failed rc={rc}")


def seed_mask(z, nodata=-9999.0, seed_mode=SEED_WBT):
    z = _f32(z)
    out = np.empty(z.shape, np.uint8)
    _chk(lib().orc_seed_mask(_p(z), _p(out), ctypes.c_int64(z.shape[0]), ctypes.c_int64(z.shape[1]),
                             ctypes.c_float(nodata), ctypes.c_int(seed_mode)), "seed_mask")
    return out


def fill(z, nodata=-9999.0, seed_mode=SEED_WBT):
    z = _f32(z)
    out = np.empty_like(z)
    _chk(lib().orc_fill(_p(z), _p(out), ctypes.c_int64(z.shape[0]), ctypes.c_int64(z.shape[1]),
                        ctypes.c_float(nodata), ctypes.c_int(seed_mode)), "fill")
    return out


def fill_eps(z, eps, nodata=-9999.0, seed_mode=SEED_WBT):
    z = _f32(z)
    out = np.empty(z.shape, np.float64)
    _chk(lib().orc_fill_eps(_p(z), _p(out), ctypes.c_int64(z.shape[0]), ctypes.c_int64(z.shape[1]),
                            ctypes.c_float(nodata), ctypes.c_int(seed_mode), ctypes.c_double(eps)), "fill_eps")
    return out


def fill_single_cell_pits(z, nodata=-9999.0):
    z = _f32(z)
    out = np.empty_like(z)
    _chk(lib().orc_fill_single_cell_pits(_p(z), _p(out), ctypes.c_int64(z.shape[0]),
                                         ctypes.c_int64(z.shape[1]), ctypes.c_float(nodata)), "fscp")
    return out


def breach_single_cell_pits(z, nodata=-9999.0):
    z = _f32(z)
    out = np.empty_like(z)
    _chk(lib().orc_breach_single_cell_pits(_p(z), _p(out), ctypes.c_int64(z.shape[0]),
                                           ctypes.c_int64(z.shape[1]), ctypes.c_float(nodata)), "bscp")
    return out


def d8(z, nodata=-9999.0, dx=1.0, dy=1.0, table=TABLE_WBT, want_slope=True):
    z = _f32(z)
    d = np.empty(z.shape, np.uint8)
    s = np.empty(z.shape, np.float32) if want_slope else None
    _chk(lib().orc_d8(_p(z), _p(d), _p(s) if want_slope else None, ctypes.c_int64(z.shape[0]),
                      ctypes.c_int64(z.shape[1]), ctypes.c_float(nodata), ctypes.c_double(dx),
                      ctypes.c_double(dy), ctypes.c_int(table)), "d8")
    return (d, s) if want_slope else d


def resolve_flats(z, dirs, nodata=-9999.0, table=TABLE_WBT):
    z = _f32(z)
    d = np.array(dirs, dtype=np.uint8, order="C", copy=True)
    _chk(lib().orc_resolve_flats(_p(z), _p(d), ctypes.c_int64(z.shape[0]), ctypes.c_int64(z.shape[1]),
                                 ctypes.c_float(nodata), ctypes.c_int(table)), "resolve_flats")
    return d


def accum(dirs, table=TABLE_WBT):
    d = np.ascontiguousarray(dirs, dtype=np.uint8)
    a = np.empty(d.shape, np.uint32)
    _chk(lib().orc_accum(_p(d), _p(a), ctypes.c_int64(d.shape[0]), ctypes.c_int64(d.shape[1]),
                         ctypes.c_int(table)), "accum")
    return a


def fill_d8_accum(z, nodata=-9999.0, dx=1.0, dy=1.0, seed_mode=SEED_TAUDEM, table=TABLE_TAUDEM,
                  flats=True):
    """The north-star chain PitRemove -> D8FlowDir -> AreaD8 (run_aws.py:580-617)."""
    f = fill(z, nodata, seed_mode)
    d, s = d8(f, nodata, dx, dy, table)
    if flats:
        d = resolve_flats(f, d, nodata, table)
    a = accum(d, table)
    return f, d, s, a


# --- pinned parts: these ARE the reference's own library calls -------------------------------

def label(mask, conn=8):
    """scipy.ndimage.label exactly as tools/depressions.py:1646-1669 calls it."""
    from scipy import ndimage
    s = [[1, 1, 1], [1, 1, 1], [1, 1, 1]] if conn == 8 else [[0, 1, 0], [1, 1, 1], [0, 1, 0]]
    lab, n = ndimage.label(np.asarray(mask) != 0, structure=s)
    return lab.astype(np.int32), int(n)


def fill_planchon_darboux(z, nodata=-9999.0, seed_mode=SEED_WBT, max_iter=100000):
    """Independent second opinion for small grids: Planchon & Darboux (2001) style relaxation
    W <- max(z, min over neighbours W) from W=+inf off the seeds, iterated to a fixed point in
    numpy.  Must equal fill() bit for bit (the eps=0 surface is unique)."""
    z = _f32(z)
    valid = ~((z == np.float32(nodata)) | np.isnan(z))
    seed = seed_mask(z, nodata, seed_mode).astype(bool)
    w = np.where(seed, z, np.float32(np.inf)).astype(np.float32)
    w[~valid] = np.inf
    for _ in range(max_iter):
        p = np.pad(w, 1, constant_values=np.inf)
        m = np.full_like(w, np.inf)
        for dr in (0, 1, 2):
            for dc in (0, 1, 2):
                if dr == 1 and dc == 1:
                    continue
                m = np.minimum(m, p[dr:dr + w.shape[0], dc:dc + w.shape[1]])
        nw = np.where(valid & ~seed, np.minimum(w, np.maximum(z, m)), w)
        if np.array_equal(nw, w):
            break
        w = nw
    out = np.where(valid & np.isfinite(w), w, z).astype(np.float32)
    return out
