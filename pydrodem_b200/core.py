"""Array-level front of the C ABI: numpy arrays go through the *_host entry points (H2D/D2H inside
the call), torch CUDA tensors through the *_dev entry points on torch's current stream.

Torch is plumbing only (device buffers, streams); every kernel lives in libhydrocore.so.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import SEED_TAUDEM, SEED_WBT, TABLE_TAUDEM, TABLE_WBT  # noqa: F401


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _np2d(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.ndim != 2:
        raise ValueError("expected a 2-D raster")
    return a


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def _tdev(t, dtype, name):
    import torch
    if not t.is_cuda:
        raise ValueError(f"{name}: torch tensors must live on a CUDA device (no CPU path)")
    if t.dtype != dtype or t.dim() != 2:
        raise ValueError(f"{name}: expected 2-D {dtype}, got {t.dtype} {tuple(t.shape)}")
    return t.contiguous()


def _stream(t):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _tp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def fill(z, nodata=-9999.0, seed_mode=SEED_WBT):
    """Depression fill with epsilon 0 (dep.fill_pits' WBT call, tools/depressions.py:1284; TauDEM
    PitRemove, run_aws.py:582-588).  Returns float32 of the same shape/container as `z`."""
    lib = _lib.load()
    if _is_torch(z):
        import torch
        z = _tdev(z, torch.float32, "fill")
        out = torch.empty_like(z)
        ctx = _lib.context(z.device.index or 0)
        with torch.cuda.device(z.device):
            _lib.check(lib.hc_fill_dev(ctx, _tp(z), _tp(out), z.shape[0], z.shape[1], nodata, seed_mode,
                                       _stream(z)), "hc_fill_dev")
        return out
    z = _np2d(z, np.float32)
    out = np.empty_like(z)
    _lib.check(lib.hc_fill_host(_lib.context(0), _vp(z), _vp(out), z.shape[0], z.shape[1], nodata, seed_mode),
               "hc_fill_host")
    return out


def d8(z, nodata=-9999.0, dx=1.0, dy=1.0, table=TABLE_WBT, want_slope=True):
    """D8 code (uint8) and optional slope (float32) — wbt.d8_pointer run_aws.py:539 /
    TauDEM D8FlowDir run_aws.py:594-603.  Device tensors only (numpy: use fill_d8_accum)."""
    import torch
    lib = _lib.load()
    host = not _is_torch(z)
    if host:
        z = torch.from_numpy(_np2d(z, np.float32)).cuda()
    z = _tdev(z, torch.float32, "d8")
    d = torch.empty(z.shape, dtype=torch.uint8, device=z.device)
    s = torch.empty(z.shape, dtype=torch.float32, device=z.device) if want_slope else None
    with torch.cuda.device(z.device):
        _lib.check(lib.hc_d8_dev(_lib.context(z.device.index or 0), _tp(z), _tp(d), _tp(s), z.shape[0], z.shape[1],
                                 nodata, dx, dy, table, _stream(z)), "hc_d8_dev")
    if host:
        d = d.cpu().numpy()
        s = s.cpu().numpy() if s is not None else None
    return (d, s) if want_slope else d


def resolve_flats(z, dirs, nodata=-9999.0, table=TABLE_WBT):
    """Assign directions over drainable flats (returns a new array; input untouched)."""
    import torch
    lib = _lib.load()
    host = not _is_torch(z)
    if host:
        z = torch.from_numpy(_np2d(z, np.float32)).cuda()
        dirs = torch.from_numpy(_np2d(dirs, np.uint8)).cuda()
    z = _tdev(z, torch.float32, "resolve_flats")
    d = _tdev(dirs, torch.uint8, "resolve_flats").clone()
    with torch.cuda.device(z.device):
        _lib.check(lib.hc_resolve_flats_dev(_lib.context(z.device.index or 0), _tp(z), _tp(d), z.shape[0],
                                            z.shape[1], nodata, table, _stream(z)), "hc_resolve_flats_dev")
    return d.cpu().numpy() if host else d


def accum(dirs, table=TABLE_WBT):
    """Cell-count accumulation (uint32 stored as int64-safe torch.int32 view -> returned as uint32
    numpy or torch.int32 bit pattern reinterpreted by the caller)."""
    import torch
    lib = _lib.load()
    host = not _is_torch(dirs)
    if host:
        dirs = torch.from_numpy(_np2d(dirs, np.uint8)).cuda()
    d = _tdev(dirs, torch.uint8, "accum")
    a = torch.empty(d.shape, dtype=torch.int32, device=d.device)
    with torch.cuda.device(d.device):
        _lib.check(lib.hc_accum_dev(_lib.context(d.device.index or 0), _tp(d), _tp(a), d.shape[0], d.shape[1],
                                    table, _stream(d)), "hc_accum_dev")
    return a.cpu().numpy().view(np.uint32) if host else a


def fill_d8_accum(z, nodata=-9999.0, dx=1.0, dy=1.0, seed_mode=SEED_TAUDEM, table=TABLE_TAUDEM, flats=True,
                  want_slope=True):
    """The north-star chain (PitRemove -> D8FlowDir -> AreaD8, run_aws.py:580-617; or WBT fill ->
    d8_pointer -> d8_flow_accumulation, run_aws.py:528-544).  Returns (filled, dir, slope, acc)."""
    lib = _lib.load()
    if _is_torch(z):
        import torch
        z = _tdev(z, torch.float32, "fill_d8_accum")
        f = torch.empty_like(z)
        d = torch.empty(z.shape, dtype=torch.uint8, device=z.device)
        s = torch.empty_like(z) if want_slope else None
        a = torch.empty(z.shape, dtype=torch.int32, device=z.device)
        with torch.cuda.device(z.device):
            _lib.check(lib.hc_fill_d8_accum_dev(_lib.context(z.device.index or 0), _tp(z), _tp(f), _tp(d), _tp(s),
                                                _tp(a), z.shape[0], z.shape[1], nodata, dx, dy, seed_mode, table,
                                                int(bool(flats)), _stream(z)), "hc_fill_d8_accum_dev")
        return f, d, s, a
    z = _np2d(z, np.float32)
    f = np.empty_like(z)
    d = np.empty(z.shape, np.uint8)
    s = np.empty_like(z) if want_slope else None
    a = np.empty(z.shape, np.uint32)
    _lib.check(lib.hc_fill_d8_accum_host(_lib.context(0), _vp(z), _vp(f), _vp(d), _vp(s), _vp(a), z.shape[0],
                                         z.shape[1], nodata, dx, dy, seed_mode, table, int(bool(flats))),
               "hc_fill_d8_accum_host")
    return f, d, s, a


def label(mask, conn=8):
    """scipy.ndimage.label-compatible labels (tools/depressions.py:1646-1669): returns (int32 labels, n)."""
    lib = _lib.load()
    n = ctypes.c_int32(0)
    if _is_torch(mask):
        import torch
        m = _tdev(mask, torch.uint8, "label")
        lab = torch.empty(m.shape, dtype=torch.int32, device=m.device)
        with torch.cuda.device(m.device):
            _lib.check(lib.hc_label_dev(_lib.context(m.device.index or 0), _tp(m), _tp(lab), m.shape[0], m.shape[1],
                                        conn, ctypes.byref(n), _stream(m)), "hc_label_dev")
        return lab, int(n.value)
    m = _np2d(np.asarray(mask) != 0, np.uint8)
    lab = np.empty(m.shape, np.int32)
    _lib.check(lib.hc_label_host(_lib.context(0), _vp(m), _vp(lab), m.shape[0], m.shape[1], conn, ctypes.byref(n)),
               "hc_label_host")
    return lab, int(n.value)


def fill_stats(device=0):
    """(basins, Boruvka rounds) of the last fill on this device's context."""
    nb, nr = ctypes.c_int64(0), ctypes.c_int32(0)
    _lib.check(_lib.load().hc_fill_stats(_lib.context(device), ctypes.byref(nb), ctypes.byref(nr)), "hc_fill_stats")
    return int(nb.value), int(nr.value)
