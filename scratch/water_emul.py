"""numpy emulation of csrc/water.cu's per-cell logic (dev check against the reference golden; not shipped)"""
import numpy as np
from scipy import ndimage, signal
BIG = np.float32(1e6)
def mean5(x):
    k = np.full((5, 5), (1. / 25.0) * 1.0)
    return signal.convolve2d(x.astype(np.float64), k, boundary='symm', mode='same').astype(np.float32)
def river_minimize(dem, wm, swo, max_burn=5, size=3, nodata=-9999):
    nodata = np.float32(nodata); mb = np.float32(max_burn)
    t0 = np.where((wm > 2) & (wm < 8), 1, np.where((wm > 0) & (wm < 3), 2, 0)).astype(np.uint8)
    test = ndimage.maximum_filter(t0, size=3)
    inner = test == 1
    tdem = np.where(inner & (dem != nodata), dem, BIG).astype(np.float32)
    minarr = ndimage.minimum_filter(tdem, size=size)
    td = tdem.copy()
    nar = inner & (swo.astype(np.int16) > 5)
    td[nar] = minarr[nar]
    ob = (dem - td) > mb
    td[ob] = (dem - mb)[ob]
    o = np.where(inner, td, dem)
    o[o == BIG] = dem[o == BIG]
    return o
def apply(mode, dem, burn, wm, swo, apply_m, minf, nodata, cut, p0, p1):
    valid = dem != np.float32(nodata)
    if mode == 0: t = (wm > 0) & (wm < 8) & valid & (swo.astype(np.int16) < cut) & (apply_m == 1)
    elif mode == 1: t = (wm > 7) & (wm != 11) & (wm != 15) & valid & (apply_m == 1)
    else: t = (wm > 2) & (wm < 8) & valid & (swo.astype(np.int16) > cut)
    d = burn - minf
    if mode == 2: sel = t & (d >= np.float32(0.01)) & (d <= np.float32(p1)); burn[sel] = minf[sel]
    else:
        sel = t & (d <= np.float32(p0)) & (d >= -np.float32(p1)); burn[sel] = (minf + np.float32(p0))[sel]
    return burn
def big(dem, src, wm, swo, excl, nodata, lo, cut):
    m = (wm > lo) & (wm < 8) & (dem != np.float32(nodata))
    if swo is not None: m &= swo.astype(np.int16) > cut
    if excl is not None: m &= excl == 0
    return m.astype(np.uint8), np.where(m, src, BIG).astype(np.float32)
def raise_small_rivers(dem, burn, wm, swo, nodata, hi=50, lo=5, extra=0.10, mr=3.0, fs=51):
    dem = dem.copy()
    mask, tdem = big(dem, burn, wm, swo, None, nodata, 2, hi)
    minf = ndimage.minimum_filter(tdem, size=fs); ap = ndimage.maximum_filter(mask, size=fs - 16)
    return apply(0, dem, burn, wm, swo, ap, minf, nodata, lo, extra, mr)
def raise_small_streams(dem, burn, wm, nodata, extra=0.5, mr=5.0, fs=51):
    dem = dem.copy()
    sm = ((wm > 7) & (wm != 11) & (wm != 15) & (dem != np.float32(nodata))).astype(np.uint8)
    sbuf = ndimage.maximum_filter(sm, size=5)
    mask, tdem = big(dem, mean5(dem), wm, None, sbuf, nodata, 3, 0)
    minf = ndimage.minimum_filter(tdem, size=fs); ap = ndimage.maximum_filter(mask, size=fs - 16)
    return apply(1, dem, burn, wm, None, ap, minf, nodata, 0, extra, mr)
def smooth_rivers(dem, burn, wm, swo, nodata, cut=90, ml=10.0, fs=21):
    _, tdem = big(dem, mean5(burn), wm, swo, None, nodata, 2, cut)
    minf = ndimage.minimum_filter(tdem, size=fs)
    return apply(2, dem, burn, wm, swo, None, minf, nodata, cut, 0, ml)
if __name__ == "__main__":
    g = np.load("tests/golden/water_reference.npz")
    for n in "abc":
        dem, burn, wm, swo = (g[f"{n}_{k}"] for k in ("dem", "burn", "wm", "swo"))
        r = {}
        r["rm_default"] = river_minimize(burn.copy(), wm, swo)
        r["rm_call"] = river_minimize(burn.copy(), wm, swo, 2, 5)
        b = burn.copy(); r["rsr_call"] = raise_small_rivers(b, b, wm, swo, -9999, 50, 5, 0.10, 2.0, 31)
        r["rsr_default"] = raise_small_rivers(dem, burn.copy(), wm, swo, -9999)
        r["rss_call"] = raise_small_streams(dem, burn.copy(), wm, -9999, 1.0, 2.0)
        r["rss_default"] = raise_small_streams(dem, burn.copy(), wm, -9999)
        b = burn.copy(); r["rss_alias"] = raise_small_streams(b, b, wm, -9999, 0.25, 1.0)
        r["sr_default"] = smooth_rivers(dem, burn.copy(), wm, swo, -9999)
        r["sr_wide"] = smooth_rivers(dem, burn.copy(), wm, swo, -9999, 40, 3.0, 9)
        print(n, {k: int(np.count_nonzero(v != g[f"{n}_{k}"])) for k, v in r.items()})
