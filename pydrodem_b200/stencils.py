"""Python front of the smoothing stencils (csrc/stencil.cu) with the reference's signatures.

Mirrors `tools/derive.py` (apply_convolve_filter :66, slope_D2 :102, curvature_D2 :133, slope_D4 :176)
and `tools/build_operators.py` build_kernel :327, plus the smoothing core of
`CreateDTM.run_tile` (models/dem_noise_removal.py:397-487).  numpy in -> numpy out, torch CUDA
tensor in -> tensor out; kernels only, no CPU arithmetic on rasters.
"""
import ctypes

import numpy as np

from . import _lib
from .core import _is_torch, _np2d, _stream, _tdev, _tp


def build_kernel(weights, dtype=np.int8, normalize=False):
    """Square ring kernel: weights[0] at the centre, weights[-1] on the outer ring
    (same construction as tools/build_operators.py:327-351; host-side, a few dozen floats)."""
    f = len(weights)
    n = 2 * f - 1
    w = np.zeros((n, n), dtype=dtype)
    c = f - 1
    for ring in range(f - 1, -1, -1):  # paint from the outside in
        w[c - ring:c + ring + 1, c - ring:c + ring + 1] = weights[ring]
    if normalize:
        w = (1. / np.sum(w)) * w
    return w


def _to_dev(a, dtype):
    import torch
    if _is_torch(a):
        return a, False
    return torch.from_numpy(_np2d(a, dtype)).cuda(), True


def apply_convolve_filter(array, kernel, normalize_kernel=False, boundary='symm', mode='same', dtype=np.float32):
    """scipy.signal.convolve2d(array, kernel, boundary='symm', mode='same') in float32."""
    import torch
    if boundary != 'symm' or mode != 'same':
        raise NotImplementedError("only boundary='symm', mode='same' is on the reference path")
    kernel = np.asarray(kernel, dtype=np.float64)
    if normalize_kernel:
        kernel = (1. / np.sum(kernel)) * kernel
    k32 = np.ascontiguousarray(kernel, dtype=np.float32)
    if not _is_torch(array):
        array = np.asarray(array).astype(np.float32)
    a, host = _to_dev(array, np.float32)
    a = _tdev(a, torch.float32, "apply_convolve_filter")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().hc_conv2d_symm_dev(_lib.context(a.device.index or 0), _tp(a),
                                                  ctypes.c_void_p(k32.ctypes.data), k32.shape[0], k32.shape[1], _tp(out),
                                                  a.shape[0], a.shape[1], _stream(a)), "hc_conv2d_symm_dev")
        torch.cuda.current_stream(a.device).synchronize()  # k32 is host memory read by the async copy
    if host:
        return out.cpu().numpy().astype(dtype)
    return out


def dw_smooth4(array, kernels):
    """The four smoothings of CreateDTM.run_tile (:399-402) in one pass: `kernels` = the 9x9, 7x7, 5x5 and 3x3 kernels
    (as build_kernel returns them).  Returns [z9, z7, z5, z3]; same arithmetic as four apply_convolve_filter calls."""
    import torch
    if [k.shape for k in kernels] != [(9, 9), (7, 7), (5, 5), (3, 3)]:
        return [apply_convolve_filter(array, k) for k in kernels]
    a, host = _to_dev(array if _is_torch(array) else np.asarray(array).astype(np.float32), np.float32)
    a = _tdev(a, torch.float32, "dw_smooth4")
    k32 = np.ascontiguousarray(np.concatenate([np.asarray(k, dtype=np.float64).astype(np.float32).ravel() for k in kernels]))
    outs = [torch.empty_like(a) for _ in range(4)]
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().hc_conv2d_symm_dw4_dev(_lib.context(a.device.index or 0), _tp(a),
                                                      ctypes.c_void_p(k32.ctypes.data), *[_tp(o) for o in outs], a.shape[0],
                                                      a.shape[1], _stream(a)), "hc_conv2d_symm_dw4_dev")
        torch.cuda.current_stream(a.device).synchronize()  # k32 is host memory read by the async copy
    return [o.cpu().numpy() for o in outs] if host else outs


_DEG_THR = None
_DEG_THR_SET = set()


def degree_thresholds():
    """thr[k - 1] = the smallest float32 slope s with np.rad2deg(np.arctan(s)).astype(int8) >= k, k = 1..90, found by
    bisection over the float32 bit patterns with numpy's own float32 arctan: what tools/derive.py:128,203 computes, as
    91 comparable intervals instead of a transcendental per cell."""
    global _DEG_THR
    if _DEG_THR is None:
        def deg(bits):
            x = np.array([bits], np.uint32).view(np.float32)
            with np.errstate(all="ignore"):
                return int(np.rad2deg(np.arctan(x)).astype(np.int8)[0])
        top = int(np.array([np.inf], np.float32).view(np.uint32)[0])
        thr = np.empty(90, np.float32)
        lo_prev = 0
        for k in range(1, 91):
            lo, hi = lo_prev, top          # deg(lo) < k (or lo == 0), deg(hi) >= k
            if deg(hi) < k:
                thr[k - 1] = np.inf
                continue
            while hi - lo > 1:
                mid = (lo + hi) // 2
                if deg(mid) >= k:
                    hi = mid
                else:
                    lo = mid
            thr[k - 1] = np.array([hi], np.uint32).view(np.float32)[0]
            lo_prev = lo
        _DEG_THR = thr
    return _DEG_THR


def _ensure_thresholds(dev_index):
    if dev_index not in _DEG_THR_SET:
        t = np.ascontiguousarray(degree_thresholds(), np.float32)
        _lib.check(_lib.load().hc_slope_thresholds_set(_lib.context(dev_index), ctypes.c_void_p(t.ctypes.data), 90),
                   "hc_slope_thresholds_set")
        _DEG_THR_SET.add(dev_index)


def _slope(array, dx, dy, d4, unit):
    import torch
    a, host = _to_dev(array, np.float32)
    a = _tdev(a, torch.float32, "slope")
    if unit == 'degree':
        _ensure_thresholds(a.device.index or 0)
    deg = torch.empty(a.shape, dtype=torch.int8, device=a.device) if unit == 'degree' else None
    mag = torch.empty_like(a) if unit != 'degree' else None
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().hc_slope_dev(_lib.context(a.device.index or 0), _tp(a), _tp(deg), _tp(mag), a.shape[0],
                                            a.shape[1], float(dx), float(dy), int(d4), _stream(a)), "hc_slope_dev")
    out = deg if unit == 'degree' else mag
    return out.cpu().numpy() if host else out


def slope_D2(array, dx, dy, unit='degree'):
    return _slope(array, dx, dy, 0, unit)


def slope_D4(array, dx, dy, unit='degree'):
    return _slope(array, dx, dy, 1, unit)


def curvature_D2(array, dx, dy, geometric=True):
    import torch
    a, host = _to_dev(array, np.float32)
    a = _tdev(a, torch.float32, "curvature_D2")
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().hc_curvature_d2_dev(_lib.context(a.device.index or 0), _tp(a), _tp(out), a.shape[0],
                                                   a.shape[1], float(dx), float(dy), int(bool(geometric)), _stream(a)),
                   "hc_curvature_d2_dev")
    return out.cpu().numpy() if host else out


def mean_std(array):
    """(mean, population std) in float64 — np.std(curvature) of models/dem_noise_removal.py:413."""
    import torch
    a, _ = _to_dev(array, np.float32)
    a = _tdev(a, torch.float32, "mean_std")
    m, s = ctypes.c_double(0), ctypes.c_double(0)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().hc_mean_std_dev(_lib.context(a.device.index or 0), _tp(a), a.numel(), ctypes.byref(m),
                                               ctypes.byref(s), _stream(a)), "hc_mean_std_dev")
    return float(m.value), float(s.value)


DW_WEIGHTS = {  # CreateDTM.setfweights, filtertype 'DW' (models/dem_noise_removal.py:1265-1280)
    9: [1., 0.8, 0.6, 0.4, 0.2], 7: [1., 0.75, 0.5, 0.25], 5: [1., 0.7, 0.35], 3: [1., 0.5]}


def dtm_select(dem, z9, z7, z5, z3, slope0, curv, sigma, curvebreaks=(1.0, 2.5), slopebreaks=(0, 2, 6, 15, 30),
               swo=None, edm=None, tx=None, artifact=None):
    """Filter-selection block of CreateDTM.run_tile (:414-471) -> (Zf float32, filter_array int8)."""
    import torch
    host = not _is_torch(dem)
    f32 = [(_to_dev(x, np.float32)[0]) for x in (dem, z9, z7, z5, z3, curv)]
    s0 = _to_dev(slope0, np.int8)[0]
    masks = [None if m is None else _to_dev(np.asarray(m, np.uint8) if not _is_torch(m) else m, np.uint8)[0]
             for m in (swo, edm, tx, artifact)]
    d = f32[0]
    zf = torch.empty_like(d)
    filt = torch.empty(d.shape, dtype=torch.int8, device=d.device)
    cb = [np.float32(c) * np.float32(sigma) for c in curvebreaks]  # np.float32 scalars, as in the reference
    sb = (ctypes.c_int * 5)(*[int(v) for v in slopebreaks])
    with torch.cuda.device(d.device):
        _lib.check(_lib.load().hc_dtm_select_dev(_lib.context(d.device.index or 0), _tp(f32[0]), _tp(f32[1]), _tp(f32[2]),
                                                 _tp(f32[3]), _tp(f32[4]), _tp(s0), _tp(f32[5]), _tp(masks[0]),
                                                 _tp(masks[1]), _tp(masks[2]), _tp(masks[3]), _tp(zf), _tp(filt),
                                                 d.numel(), float(cb[0]), float(cb[1]), ctypes.cast(sb, ctypes.c_void_p),
                                                 _stream(d)), "hc_dtm_select_dev")
    if host:
        return zf.cpu().numpy(), filt.cpu().numpy()
    return zf, filt


def _dw_kernels32(fweights=None):
    fweights = fweights or [DW_WEIGHTS[9], DW_WEIGHTS[7], DW_WEIGHTS[5], DW_WEIGHTS[3]]
    ks = [build_kernel(w, dtype=np.float32, normalize=True) for w in fweights]
    if [k.shape for k in ks] != [(9, 9), (7, 7), (5, 5), (3, 3)]:
        return None, ks
    return np.ascontiguousarray(np.concatenate([np.asarray(k, dtype=np.float64).astype(np.float32).ravel() for k in ks])), ks


def dtm_pre(dem, k32, dx, dy, gdir=4, row_lo=0, row_hi=None, rasters=True):
    """First fused pass (hc_dtm_pre_dev): curvature_D2 and slope_D4 / slope_D2 of the 5x5-smoothed DEM without the
    smoothed raster, and the float64 (sum, sum of squares) of the curvature over rows [row_lo, row_hi) that
    np.std(curvature) of models/dem_noise_removal.py:413 needs.  Returns (curv, slope0, s1, s2)."""
    import torch
    _ensure_thresholds(dem.device.index or 0)
    out = (ctypes.c_double * 2)()
    ny, nx = dem.shape
    curv = torch.empty_like(dem) if rasters else None
    slope0 = torch.empty(dem.shape, dtype=torch.int8, device=dem.device) if rasters else None
    with torch.cuda.device(dem.device):
        _lib.check(_lib.load().hc_dtm_pre_dev(_lib.context(dem.device.index or 0), _tp(dem), ctypes.c_void_p(k32.ctypes.data),
                                              _tp(curv), _tp(slope0), ny, nx, float(dx), float(dy), 1 if gdir == 4 else 0,
                                              int(row_lo), int(ny if row_hi is None else row_hi), out, _stream(dem)),
                   "hc_dtm_pre_dev")
    return curv, slope0, float(out[0]), float(out[1])


def dtm_apply(dem, k32, curv, slope0, sigma, curvebreaks=(1.0, 2.5), slopebreaks=(0, 2, 6, 15, 30), swo=None, edm=None,
              tx=None, artifact=None):
    """Second fused pass (hc_dtm_apply_dev): the four smoothings and the selection block of CreateDTM.run_tile
    (:399-402, :414-471) in one tile pass -> (Zf, filter code)."""
    import torch
    masks = [None if m is None else _to_dev(np.asarray(m, np.uint8) if not _is_torch(m) else m, np.uint8)[0]
             for m in (swo, edm, tx, artifact)]
    zf = torch.empty_like(dem)
    filt = torch.empty(dem.shape, dtype=torch.int8, device=dem.device)
    cb = [np.float32(c) * np.float32(sigma) for c in curvebreaks]  # np.float32 scalars, as in the reference
    sb = (ctypes.c_int * 5)(*[int(v) for v in slopebreaks])
    with torch.cuda.device(dem.device):
        _lib.check(_lib.load().hc_dtm_apply_dev(_lib.context(dem.device.index or 0), _tp(dem), ctypes.c_void_p(k32.ctypes.data),
                                                _tp(curv), _tp(slope0), _tp(masks[0]), _tp(masks[1]), _tp(masks[2]),
                                                _tp(masks[3]), _tp(zf), _tp(filt), dem.shape[0], dem.shape[1],
                                                float(cb[0]), float(cb[1]), ctypes.cast(sb, ctypes.c_void_p), _stream(dem)),
                   "hc_dtm_apply_dev")
        torch.cuda.current_stream(dem.device).synchronize()  # k32 is host memory read by the async copy
    return zf, filt


def smooth_core_unfused(dem, dx, dy, gdir=4, fweights=None, slopebreaks=(0, 2, 6, 15, 30), curvebreaks=(1.0, 2.5), swo=None,
                        edm=None, tx=None, artifact=None):
    """The core as a chain of the single-purpose calls (every intermediate raster in HBM): the form the fused path is
    checked against, and the path for filter sets other than the four DW kernels."""
    host = not _is_torch(dem)
    d = _to_dev(dem, np.float32)[0]
    fweights = fweights or [DW_WEIGHTS[9], DW_WEIGHTS[7], DW_WEIGHTS[5], DW_WEIGHTS[3]]
    zs = dw_smooth4(d, [build_kernel(w, dtype=np.float32, normalize=True) for w in fweights])
    s0ind = 2  # 0=9x9, 1=7x7, 2=5x5, 3=3x3: the 5x5 result drives slope/curvature (:147, :405-410)
    slope0 = (slope_D4 if gdir == 4 else slope_D2)(zs[s0ind], dx, dy)
    curv = curvature_D2(zs[s0ind], dx, dy)
    _, sigma = mean_std(curv)
    sigma32 = np.float32(sigma)
    zf, filt = dtm_select(d, zs[0], zs[1], zs[2], zs[3], slope0, curv, sigma32, curvebreaks, slopebreaks, swo, edm, tx,
                          artifact)
    slope = (slope_D4 if gdir == 4 else slope_D2)(zf, dx, dy)
    if host:
        return zf.cpu().numpy(), filt.cpu().numpy(), slope.cpu().numpy()
    return zf, filt, slope


def smooth_core(dem, dx, dy, gdir=4, fweights=None, slopebreaks=(0, 2, 6, 15, 30), curvebreaks=(1.0, 2.5), swo=None,
                edm=None, tx=None, artifact=None):
    """Smoothing core of CreateDTM.run_tile (models/dem_noise_removal.py:397-476), device resident:
    four DW smoothings, slope/curvature of the 5x5 result, global sigma, filter selection, final slope -- as two fused
    tile passes over the DEM (curvature + slope of the 5x5 smoothing with the sigma sums; then the smoothings + selection)
    and the slope of Zf.
    Returns (Zf float32, filter_array int8, slope_array int8)."""
    import torch
    k32, _ = _dw_kernels32(fweights)
    if k32 is None:
        return smooth_core_unfused(dem, dx, dy, gdir, fweights, slopebreaks, curvebreaks, swo, edm, tx, artifact)
    host = not _is_torch(dem)
    d = _to_dev(dem, np.float32)[0]
    d = _tdev(d, torch.float32, "smooth_core")
    curv, slope0, s1, s2 = dtm_pre(d, k32, dx, dy, gdir)
    sigma = _sigma_from_sums(d.numel(), s1, s2)
    zf, filt = dtm_apply(d, k32, curv, slope0, np.float32(sigma), curvebreaks, slopebreaks, swo, edm, tx, artifact)
    del curv, slope0
    slope = (slope_D4 if gdir == 4 else slope_D2)(zf, dx, dy)
    if host:
        return zf.cpu().numpy(), filt.cpu().numpy(), slope.cpu().numpy()
    return zf, filt, slope


# ---- row bands (SURVEY.md §8e "smoothing stencils": bands + halo rows; sigma by all_reduce) -----------------------
SMOOTH_HALO = 8   # rows: 4 (9x9 window) + 2 (curvature of the 5x5 result) + 1 (slope of Zf) and a margin


def _sum_shifted(t, shift):
    import torch
    out = (ctypes.c_double * 2)()
    x = t.reshape(-1)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().hc_sum_shifted_dev(_lib.context(x.device.index or 0), _tp(x), x.numel(), float(shift), out,
                                                  _stream(x)), "hc_sum_shifted_dev")
    return float(out[0]), float(out[1])


class BandSmoother:
    """Smoothing core of CreateDTM.run_tile on ONE row band that carries SMOOTH_HALO halo rows where it has a
    neighbour (`top` / `bot`).  Three phases so that the caller can all_reduce the sigma sums in between:
    a = stage1() -> (n, sum x); stage2(mean) -> (sum(x-m), sum((x-m)^2)); finish(sigma) -> (Zf, filter, slope) of the
    own rows.  Everything away from the outer rows of the haloed array is identical to the whole-raster result
    (the 'symm' padding and one-sided gradients only touch rows that are cropped)."""

    def __init__(self, dem_h, top, bot, dx, dy, gdir=4, slopebreaks=(0, 2, 6, 15, 30), curvebreaks=(1.0, 2.5)):
        self.d, self.top, self.bot = dem_h, bool(top), bool(bot)
        self.dx, self.dy, self.gdir = dx, dy, gdir
        self.sb, self.cb = slopebreaks, curvebreaks
        self.lo = SMOOTH_HALO if self.top else 0
        self.hi = dem_h.shape[0] - (SMOOTH_HALO if self.bot else 0)

    def stage1(self):
        self.k32, _ = _dw_kernels32()
        self.curv, self.slope0, self.s1, self.s2 = dtm_pre(self.d, self.k32, self.dx, self.dy, self.gdir, self.lo, self.hi)
        self.n = (self.hi - self.lo) * self.d.shape[1]
        return self.n, self.s1

    def stage2(self, mean):
        # sum(x - m), sum((x - m)^2) from the unshifted float64 sums
        return self.s1 - self.n * mean, self.s2 - 2.0 * mean * self.s1 + self.n * mean * mean

    def finish(self, sigma):
        zf, filt = dtm_apply(self.d, self.k32, self.curv, self.slope0, np.float32(sigma), self.cb, self.sb)
        sl = slope_D4 if self.gdir == 4 else slope_D2
        slope = sl(zf, self.dx, self.dy)
        return zf[self.lo:self.hi], filt[self.lo:self.hi], slope[self.lo:self.hi]


def _sigma_from_sums(n, d1, d2):
    d = d1 / n
    return float(np.sqrt(max(d2 / n - d * d, 0.0)))


def smooth_core_banded(dem, nbands, dx, dy, **kw):
    """Single-process emulation: cut `dem` (torch CUDA tensor or numpy) into row bands with halos, run the
    three phases band by band, combine the sigma sums as the all_reduce would, stitch.  Same return as smooth_core."""
    import torch
    host = not _is_torch(dem)
    d = _to_dev(dem, np.float32)[0]
    ny = d.shape[0]
    base, extra = divmod(ny, nbands)
    rows = [base + (1 if i < extra else 0) for i in range(nbands)]
    r0s = np.concatenate([[0], np.cumsum(rows)])
    sm = []
    for i in range(nbands):
        a, b = int(r0s[i]), int(r0s[i + 1])
        top, bot = i > 0, i < nbands - 1
        sm.append(BandSmoother(d[max(0, a - SMOOTH_HALO) if top else a: min(ny, b + SMOOTH_HALO) if bot else b].contiguous(),
                               top, bot, dx, dy, **kw))
    s = [x.stage1() for x in sm]
    n = sum(v[0] for v in s)
    mean = sum(v[1] for v in s) / n
    s2 = [x.stage2(mean) for x in sm]
    sigma = _sigma_from_sums(n, sum(v[0] for v in s2), sum(v[1] for v in s2))
    outs = [x.finish(sigma) for x in sm]
    zf, filt, slope = (torch.cat([o[k] for o in outs]) for k in range(3))
    if host:
        return zf.cpu().numpy(), filt.cpu().numpy(), slope.cpu().numpy()
    return zf, filt, slope


def smooth_core_rank(dem_h, top, bot, dx, dy, group=None, **kw):
    """One band per rank over torch.distributed: `dem_h` = this rank's rows plus SMOOTH_HALO halo rows where a
    neighbour exists (the caller exchanges them, e.g. with bands.exchange_halos on 8-row seams); sigma by two
    all_reduces of (n, sum) and (sum, sum of squares).  Returns this band's (Zf, filter, slope)."""
    import torch
    import torch.distributed as dist
    x = BandSmoother(dem_h, top, bot, dx, dy, **kw)
    t = torch.tensor(list(x.stage1()), dtype=torch.float64, device=dem_h.device)
    dist.all_reduce(t, group=group)
    n, mean = float(t[0]), float(t[1] / t[0])
    t2 = torch.tensor(list(x.stage2(mean)), dtype=torch.float64, device=dem_h.device)
    dist.all_reduce(t2, group=group)
    return x.finish(_sigma_from_sums(n, float(t2[0]), float(t2[1])))


def clipbuffer_arr(buffered_data, buffer, set_nodata=None):
    """CreateDTM.clipbuffer_arr (models/dem_noise_removal.py:1157-1181): drop `buffer` cells from every edge, or --
    with `set_nodata` -- keep the shape and overwrite that frame (how run_tile blanks the halo of Zf / Filtermask / Slope,
    :479-484).  Works on numpy arrays and on CUDA tensors (a border assignment, no kernel needed)."""
    b = buffer
    if set_nodata is None:
        return buffered_data[b:-b, b:-b]
    out = buffered_data.clone() if _is_torch(buffered_data) else np.array(buffered_data, copy=True)
    out[:b, :] = set_nodata
    out[-b:, :] = set_nodata
    out[:, :b] = set_nodata
    out[:, -b:] = set_nodata
    return out
