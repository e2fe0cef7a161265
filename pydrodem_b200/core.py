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


def fill_eps(z, eps, nodata=-9999.0, seed_mode=SEED_WBT, return_sweeps=False):
    """Priority-Flood + epsilon: the fill with an imposed gradient `eps` per step over filled areas (WhiteboxTools'
    fix_flats / flat_increment > 0; tools/depressions.py:195-198, :1271-1286).  float64 out (WhiteboxTools writes
    float64; dep.fill_pits narrows it unless downcast=False).  numpy in -> numpy out, CUDA tensor in -> CUDA tensor."""
    import torch
    lib = _lib.load()
    host = not _is_torch(z)
    zt = torch.from_numpy(_np2d(z, np.float32)).cuda() if host else _tdev(z, torch.float32, "fill_eps")
    out = torch.empty(zt.shape, dtype=torch.float64, device=zt.device)
    n = ctypes.c_int64(0)
    with torch.cuda.device(zt.device):
        _lib.check(lib.hc_fill_eps_dev(_lib.context(zt.device.index or 0), _tp(zt), _tp(out), zt.shape[0], zt.shape[1],
                                       nodata, seed_mode, float(eps), ctypes.byref(n), _stream(zt)), "hc_fill_eps_dev")
    res = out.cpu().numpy() if host else out
    return (res, int(n.value)) if return_sweeps else res


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


RASTER_ALIGN = 64   # elements: rows of a pitched device raster start 64 elements apart at least (16-byte aligned for uint8)


def empty_raster(ny, nx, dtype, device):
    """Pitched device raster: a (ny, nx) view of a (ny, ld) tensor, ld = nx rounded up to RASTER_ALIGN elements, so
    that every row is 16-byte aligned whatever nx is (what TMA tile copies and 128-bit loads need; a dense
    10801-wide raster offers neither).  fill_d8_accum() recognises such views (attribute `_hc_ld`), runs
    hc_fill_d8_accum_ld_dev on them and returns pitched views; the padding columns belong to the library."""
    import torch
    ld = (nx + RASTER_ALIGN - 1) // RASTER_ALIGN * RASTER_ALIGN
    base = torch.empty((ny, ld), dtype=dtype, device=device)
    v = base[:, :nx]
    v._hc_ld = ld
    return v


def _pitched(t):
    ld = getattr(t, "_hc_ld", None)
    if ld is None or t.dim() != 2 or t.stride(1) != 1 or t.stride(0) != ld or t.storage_offset() != 0:
        return None
    return ld


def fill_d8_accum(z, nodata=-9999.0, dx=1.0, dy=1.0, seed_mode=SEED_TAUDEM, table=TABLE_TAUDEM, flats=True,
                  want_slope=True):
    """The north-star chain (PitRemove -> D8FlowDir -> AreaD8, run_aws.py:580-617; or WBT fill ->
    d8_pointer -> d8_flow_accumulation, run_aws.py:528-544).  Returns (filled, dir, slope, acc).
    A pitched device raster from empty_raster() takes the TMA / vector path and gets pitched views back."""
    lib = _lib.load()
    if _is_torch(z):
        import torch
        ld = _pitched(z) if z.is_cuda and z.dtype == torch.float32 else None
        if ld is not None:
            ny, nx = z.shape
            f = empty_raster(ny, nx, torch.float32, z.device)
            d = empty_raster(ny, nx, torch.uint8, z.device)
            s = empty_raster(ny, nx, torch.float32, z.device) if want_slope else None
            a = empty_raster(ny, nx, torch.int32, z.device)
            with torch.cuda.device(z.device):
                _lib.check(lib.hc_fill_d8_accum_ld_dev(_lib.context(z.device.index or 0), _tp(z), _tp(f), _tp(d), _tp(s),
                                                       _tp(a), ny, nx, ld, nodata, dx, dy, seed_mode, table,
                                                       int(bool(flats)), _stream(z)), "hc_fill_d8_accum_ld_dev")
            return f, d, s, a
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


def _unary_f32(z, fn_name, nodata):
    import torch
    lib = _lib.load()
    host = not _is_torch(z)
    if host:
        z = torch.from_numpy(_np2d(z, np.float32)).cuda()
    z = _tdev(z, torch.float32, fn_name)
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        _lib.check(getattr(lib, fn_name)(_lib.context(z.device.index or 0), _tp(z), _tp(out), z.shape[0], z.shape[1],
                                         nodata, _stream(z)), fn_name)
    return out.cpu().numpy() if host else out


def fill_single_cell_pits(z, nodata=-9999.0):
    """wbt.fill_single_cell_pits (models/depression_handling.py:261, models/find_sinks.py:118)."""
    return _unary_f32(z, "hc_fill_single_cell_pits_dev", nodata)


def breach_single_cell_pits(z, nodata=-9999.0):
    """wbt.breach_single_cell_pits (models/depression_handling.py:259)."""
    return _unary_f32(z, "hc_breach_single_cell_pits_dev", nodata)


def _breach_flowpath(zh, nodata, max_length, max_depth, fill_pits):
    """Host-walk composition of the order-independent flow-path breach (specification oracle/breach_flowpath.py):
    single-cell pits, fill, D8 and flat resolution on the GPU, the pointer walks on the host
    (`hc_breach_flowpath_host`), closing fill on the GPU.  Kept as a second implementation the device form is
    checked against.  Returns (surface, stats)."""
    import ctypes
    import numpy as np
    w = _np2d(fill_single_cell_pits(zh, nodata) if fill_pits else zh, np.float32)
    filled = _np2d(fill(w, nodata=nodata, seed_mode=SEED_WBT), np.float32)
    # flats on a seed's own level must drain through the seed: pointers come from the filled surface with every seed
    # lowered by one ulp (a host touch-up of the seed cells only, until the device pass has the seed mask itself)
    seed = np.empty(w.shape, np.uint8)
    _lib.check(_lib.load().hc_wbt_seed_mask_host(_vp(w), _vp(seed), w.shape[0], w.shape[1], float(nodata)),
               "hc_wbt_seed_mask_host")
    lowered = filled.copy()
    sm = seed.astype(bool)
    lowered[sm] = np.nextafter(filled[sm], np.float32(-np.inf))
    dirs = d8(lowered, nodata=nodata, table=TABLE_WBT, want_slope=False)
    dirs = resolve_flats(lowered, dirs, nodata=nodata, table=TABLE_WBT)
    cut = np.empty_like(w)
    st = (ctypes.c_int64 * 4)()
    _lib.check(_lib.load().hc_breach_flowpath_host(_vp(w), _vp(_np2d(filled, np.float32)), _vp(_np2d(dirs, np.uint8)),
                                                  _vp(cut), w.shape[0], w.shape[1], float(nodata),
                                                  int(max_length) if max_length else 0,
                                                  float(max_depth) if max_depth else 0.0, st),
               "hc_breach_flowpath_host")
    return cut, st


def _breach_flowpath_dev(zh, nodata, max_length, max_depth, fill_pits, dx=1.0, dy=1.0):
    """The flow-path breach entirely on the device (`hc_breach_flowpath_dev`: seeds, fill, D8 + flats of the
    seed-lowered filled surface, one thread per pit walking its flow path with a CAS-min on the cut surface).
    `zh`: numpy array or CUDA tensor.  Returns (cut surface as a CUDA tensor, stats)."""
    import ctypes
    import torch
    zt = zh if _is_torch(zh) else torch.from_numpy(np.ascontiguousarray(zh, dtype=np.float32)).cuda()
    zt = _tdev(zt, torch.float32, "breach_depressions")
    w = fill_single_cell_pits(zt, nodata) if fill_pits else zt
    out, filled = torch.empty_like(w), torch.empty_like(w)
    dirs = torch.empty(w.shape, dtype=torch.uint8, device=w.device)
    seed = torch.empty(w.shape, dtype=torch.uint8, device=w.device)
    st = (ctypes.c_int64 * 4)()
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().hc_breach_flowpath_dev(_lib.context(w.device.index or 0), _tp(w), _tp(out), _tp(filled),
                                                     _tp(dirs), _tp(seed), w.shape[0], w.shape[1], float(nodata),
                                                     float(dx), float(dy), int(max_length) if max_length else 0,
                                                     float(max_depth) if max_depth else 0.0, st, _stream(w)),
                   "hc_breach_flowpath_dev")
    return out, st


def breach_depressions(z, nodata=-9999.0, max_length=None, max_depth=None, fill_pits=False, flat_increment=None,
                       return_stats=False, method='flowpath_dev'):
    """wbt.breach_depressions (Lindsay 2016) as dep.breach_pits calls it (tools/depressions.py:1376-1380): constrained
    breaching, then the depressions that could not be breached are filled with WhiteboxTools' seed rule.
    float32 in, float32 out (numpy in -> numpy out, CUDA tensor in -> CUDA tensor out); with flat_increment > 0 the closing
    fill is Priority-Flood + epsilon (core.fill_eps) and the result is float64.

    method='flowpath_dev' (default): the order-independent flow-path formulation, every pass on the GPU
    (specification: oracle/breach_flowpath.py).  'flowpath': the same formulation with the pointer walks
    on the host.  'flood': the priority-order formulation on the host (`hc_breach_depressions_host`, csrc/breach.cu).
    WhiteboxTools itself is unavailable, so all three are restatements of Lindsay (2016): parity unpinned."""
    import ctypes
    import numpy as np
    eps = float(flat_increment) if (flat_increment is not None and flat_increment > 0) else 0.0
    if eps > 0.0 and method != 'flowpath_dev':
        raise NotImplementedError("flat_increment > 0 is served by the default method only")
    if method not in ('flood', 'flowpath', 'flowpath_dev'):
        raise ValueError("method must be 'flowpath_dev' (default), 'flowpath' or 'flood'")
    host = not _is_torch(z)
    if method == 'flowpath_dev':
        cut, st = _breach_flowpath_dev(z if not host else _np2d(z, np.float32), nodata, max_length, max_depth, fill_pits)
        # what could not be breached is filled; with flat_increment > 0 by the epsilon fill (float64 out, like WhiteboxTools)
        out = fill_eps(cut, eps, nodata=nodata, seed_mode=SEED_WBT) if eps > 0.0 else fill(cut, nodata=nodata, seed_mode=SEED_WBT)
        if host:
            out = out.cpu().numpy()
        if return_stats:
            return out, dict(pits=st[0], breached=st[1], left_to_fill=st[2], cells_lowered=st[3])
        return out
    zh = _np2d(z if host else z.detach().cpu().numpy(), np.float32)
    if method == 'flowpath':
        cut, st = _breach_flowpath(zh, nodata, max_length, max_depth, fill_pits)
        out = fill(cut, nodata=nodata, seed_mode=SEED_WBT)
        if not host:
            import torch
            out = torch.from_numpy(out).to(z.device)
        if return_stats:
            return out, dict(pits=st[0], breached=st[1], left_to_fill=st[2], cells_lowered=st[3])
        return out
    cut = np.empty_like(zh)
    st = (ctypes.c_int64 * 5)()
    _lib.check(_lib.load().hc_breach_depressions_host(_vp(zh), _vp(cut), zh.shape[0], zh.shape[1], float(nodata),
                                                     int(max_length) if max_length else 0,
                                                     float(max_depth) if max_depth else 0.0, int(bool(fill_pits)), 0,
                                                     st), "hc_breach_depressions_host")
    if host:
        out = fill(cut, nodata=nodata, seed_mode=SEED_WBT)
    else:
        import torch
        out = fill(torch.from_numpy(cut).to(z.device), nodata=nodata, seed_mode=SEED_WBT)
    if return_stats:
        return out, dict(pits=st[0], breached=st[1], left_to_fill=st[2], cells_lowered=st[3])
    return out


def breach_depressions_least_cost(z, dist, nodata=-9999.0, max_cost=None, min_dist=True, flat_increment=None, fill=True,
                                  return_stats=False):
    """wbt.breach_depressions_least_cost (Lindsay & Dhun 2015) as dep.breach_pits(method='LC') and
    dep.combined_solution(carve_method='LC') call it (tools/depressions.py:1370-1374, :996-1000): every pit is breached
    along the cheapest path (cost = elevation removed) to a lower cell or the raster's edge inside a (2 dist + 1)^2
    window, unless that costs more than max_cost; min_dist prefers the shortest breach; flat_increment is the channel's
    gradient; fill=True fills what is left (WhiteboxTools' seed rule).  One CTA per pit on the GPU
    (`hc_breach_least_cost_dev`), order-independent form, specification oracle/breach_leastcost.py -- the tool itself is
    not in the reference tree: parity unpinned.  float32 in, float32 out (numpy -> numpy, CUDA tensor -> CUDA tensor)."""
    import ctypes
    import torch
    host = not _is_torch(z)
    zt = torch.from_numpy(np.ascontiguousarray(_np2d(z, np.float32))).cuda() if host else z
    zt = _tdev(zt, torch.float32, "breach_depressions_least_cost")
    out = torch.empty_like(zt)
    st = (ctypes.c_int64 * 4)()
    with torch.cuda.device(zt.device):
        _lib.check(_lib.load().hc_breach_least_cost_dev(_lib.context(zt.device.index or 0), _tp(zt), _tp(out), zt.shape[0],
                                                       zt.shape[1], float(nodata), int(dist),
                                                       float(max_cost) if max_cost is not None else 0.0,
                                                       int(bool(min_dist)),
                                                       float(flat_increment) if flat_increment else 0.0, st, _stream(zt)),
                   "hc_breach_least_cost_dev")
    if fill:
        out = globals()["fill"](out, nodata=nodata, seed_mode=SEED_WBT)   # (the keyword shadows the module's fill)
    if host:
        out = out.cpu().numpy()
    if return_stats:
        return out, dict(pits=st[0], breached=st[1], left=st[2], cells_lowered=st[3])
    return out


def taudem_check(p, fel, flat_mask, edm, nodata=-9999.0):
    """TaudemCheck.check_basin's problem mask (models/taudem_check.py:143-147): returns (mask uint8, count).
    `p` is the uint8 D8 raster with 255 = nodata."""
    import torch
    lib = _lib.load()
    host = not _is_torch(p)
    if host:
        p = torch.from_numpy(_np2d(p, np.uint8)).cuda()
        fel = torch.from_numpy(_np2d(fel, np.float32)).cuda()
        flat_mask = torch.from_numpy(_np2d(flat_mask, np.uint8)).cuda()
        edm = torch.from_numpy(_np2d(edm, np.uint8)).cuda()
    p = _tdev(p, torch.uint8, "taudem_check")
    fel = _tdev(fel, torch.float32, "taudem_check")
    flat_mask = _tdev(flat_mask, torch.uint8, "taudem_check")
    edm = _tdev(edm, torch.uint8, "taudem_check")
    mask = torch.empty_like(p)
    n = ctypes.c_int64(0)
    with torch.cuda.device(p.device):
        _lib.check(lib.hc_taudem_check_dev(_lib.context(p.device.index or 0), _tp(p), _tp(fel), _tp(flat_mask), _tp(edm),
                                           _tp(mask), ctypes.byref(n), p.shape[0], p.shape[1], nodata, _stream(p)),
                   "hc_taudem_check_dev")
    return (mask.cpu().numpy() if host else mask), int(n.value)


def taudem_fix(dem, flat_mask, nodata=-9999.0, bump=0.10):
    """TaudemCheck.fix_basin (models/taudem_check.py:183-194): water cells (+flat mask > 0) += bump."""
    import torch
    lib = _lib.load()
    host = not _is_torch(dem)
    if host:
        dem = torch.from_numpy(_np2d(dem, np.float32)).cuda()
        flat_mask = torch.from_numpy(_np2d(flat_mask, np.uint8)).cuda()
    dem = _tdev(dem, torch.float32, "taudem_fix")
    flat_mask = _tdev(flat_mask, torch.uint8, "taudem_fix")
    out = torch.empty_like(dem)
    with torch.cuda.device(dem.device):
        _lib.check(lib.hc_taudem_fix_dev(_lib.context(dem.device.index or 0), _tp(dem), _tp(flat_mask), _tp(out),
                                         dem.shape[0], dem.shape[1], nodata, bump, _stream(dem)), "hc_taudem_fix_dev")
    return out.cpu().numpy() if host else out


def diff_round(a, b, decimals=3):
    """np.round(a - b, decimals) for float32 rasters on the GPU (hc_diff_round_dev): the difference maps of
    dep.fill_pits / dep.breach_pits (tools/depressions.py:1308, :1395).  numpy in -> numpy out, tensor in -> tensor out."""
    import torch
    lib = _lib.load()
    host = not _is_torch(a)
    if host:
        a = torch.from_numpy(_np2d(a, np.float32)).cuda()
        b = torch.from_numpy(_np2d(b, np.float32)).cuda()
    a = _tdev(a, torch.float32, "diff_round")
    b = _tdev(b, torch.float32, "diff_round")
    if a.shape != b.shape:
        raise ValueError("diff_round: shapes differ")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(lib.hc_diff_round_dev(_lib.context(a.device.index or 0), _tp(a), _tp(b), _tp(out), a.numel(), int(decimals),
                                         _stream(a)), "hc_diff_round_dev")
    return out.cpu().numpy() if host else out
