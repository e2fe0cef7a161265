"""The array-level water stencils of `tools/water.py` on the GPU, with the reference's signatures.

`river_minimize` (:786), `smooth_rivers` (:928), `raise_small_streams` (:960) and `raise_small_rivers` (:1010)
are chains of mask predicates, `scipy.ndimage.minimum_filter` / `maximum_filter` over square windows of up to 51
cells, a 5x5 mean and a per-cell select; csrc/water.cu runs each chain as a handful of kernels with its scratch
rasters in the context's arena.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.  Like the reference,
`DEM_burn_arr` is modified in place and returned.  Rasters are float32 (the reference's DEMs are); the water mask
may be any integer type with values 0..255 and the SWO raster any integer type that fits int16.

`minimum_filter` / `maximum_filter` (size=k, mode 'reflect') are exported too, for uint8 and float32 rasters.
`enforce_monotonic_rivers` (:1407) runs its raster-wide steps (level mask, 4-connected clumps, per-clump statistics)
on the GPU and the per-clump window tests on the host; `fill_missing_data` is the array form of the
wbt.fill_missing_data call it ends with.  Not mirrored: `enforce_monotonic_streams` / `stream_minimize` /
`raise_canals` (vector / TauDEM stream-network inputs inside the function body).
"""
import numpy as np

from .. import _lib
from ..core import _is_torch, _stream, _tp


def _dev(a, dtype, what):
    import torch
    t = a if _is_torch(a) else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    if not t.is_cuda:
        raise RuntimeError(f"{what}: CUDA tensor required (no CPU fallback)")
    if t.dim() != 2:
        raise ValueError(f"{what}: 2-D raster expected")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _call(name, dev, *args):
    import torch
    with torch.cuda.device(dev):
        _lib.check(getattr(_lib.load(), name)(_lib.context(dev.index or 0), *args), name)


def _minmax(array, size, is_max):
    import torch
    host = not _is_torch(array)
    src = array if not host else np.asarray(array)
    if (src.dtype == torch.uint8) if not host else (src.dtype in (np.uint8, np.bool_)):
        a, eb = _dev(src if not host else src.astype(np.uint8), torch.uint8, "minmax_filter"), 1
    else:
        a, eb = _dev(src, torch.float32, "minmax_filter"), 4
    out = torch.empty_like(a)
    _call("hc_minmax_filter_dev", a.device, _tp(a), _tp(out), a.shape[0], a.shape[1], int(size), int(is_max), eb,
          _stream(a))
    return out.cpu().numpy() if host else out


def minimum_filter(array, size):
    """scipy.ndimage.minimum_filter(array, size=size) for a uint8 or float32 raster"""
    return _minmax(array, size, False)


def maximum_filter(array, size):
    """scipy.ndimage.maximum_filter(array, size=size) for a uint8 or float32 raster"""
    return _minmax(array, size, True)


def _finish_inplace(given, result):
    """the reference returns the (modified) DEM_burn_arr object itself"""
    if _is_torch(given):
        if given.data_ptr() != result.data_ptr():
            given.copy_(result)
        return given
    if isinstance(given, np.ndarray) and given.dtype == np.float32 and given.flags.writeable:
        given[...] = result.cpu().numpy()
        return given
    return result.cpu().numpy()


def river_minimize(inDEM_array, water_mask_array, SWO_array, max_burn=5, river_filter_size=3, nodataval=-9999):
    """tools/water.py:786-847"""
    import torch
    host = not _is_torch(inDEM_array)
    dem = _dev(inDEM_array, torch.float32, "river_minimize")
    wm = _dev(water_mask_array, torch.uint8, "water_mask_array")
    swo = _dev(SWO_array, torch.int16, "SWO_array")
    out = torch.empty_like(dem)
    _call("hc_river_minimize_dev", dem.device, _tp(dem), _tp(wm), _tp(swo), float(nodataval), float(max_burn),
          int(river_filter_size), _tp(out), dem.shape[0], dem.shape[1], _stream(dem))
    return out.cpu().numpy() if host else out


def raise_small_rivers(inDEM_array, DEM_burn_arr, water_mask_array, SWO_array, nodata, SWO_cutoff_high=50,
                       SWO_cutoff_low=5, extra_raise=0.10, max_raise=3.0, filter_size=51):
    """tools/water.py:1010-1045"""
    import torch
    same = inDEM_array is DEM_burn_arr
    burn = _dev(DEM_burn_arr, torch.float32, "DEM_burn_arr")
    dem = burn if same else _dev(inDEM_array, torch.float32, "raise_small_rivers")
    wm = _dev(water_mask_array, torch.uint8, "water_mask_array")
    swo = _dev(SWO_array, torch.int16, "SWO_array")
    _call("hc_raise_small_rivers_dev", burn.device, _tp(dem), _tp(burn), _tp(wm), _tp(swo), float(nodata),
          int(SWO_cutoff_high), int(SWO_cutoff_low), float(np.float32(extra_raise)), float(np.float32(max_raise)),
          int(filter_size), burn.shape[0], burn.shape[1], _stream(burn))
    return _finish_inplace(DEM_burn_arr, burn)


def raise_small_streams(inDEM_array, DEM_burn_arr, water_mask_array, nodata, extra_raise=0.50, max_raise=5.0, fsize=51):
    """tools/water.py:960-1007"""
    import torch
    same = inDEM_array is DEM_burn_arr
    burn = _dev(DEM_burn_arr, torch.float32, "DEM_burn_arr")
    dem = burn if same else _dev(inDEM_array, torch.float32, "raise_small_streams")
    wm = _dev(water_mask_array, torch.uint8, "water_mask_array")
    _call("hc_raise_small_streams_dev", burn.device, _tp(dem), _tp(burn), _tp(wm), float(nodata),
          float(np.float32(extra_raise)), float(np.float32(max_raise)), int(fsize), burn.shape[0], burn.shape[1],
          _stream(burn))
    return _finish_inplace(DEM_burn_arr, burn)


def smooth_rivers(inDEM_array, DEM_burn_arr, water_mask_array, SWO_array, nodata, SWO_cut=90, max_lower=10.0,
                  filter_size=21):
    """tools/water.py:928-957"""
    import torch
    same = inDEM_array is DEM_burn_arr
    burn = _dev(DEM_burn_arr, torch.float32, "DEM_burn_arr")
    dem = burn if same else _dev(inDEM_array, torch.float32, "smooth_rivers")
    wm = _dev(water_mask_array, torch.uint8, "water_mask_array")
    swo = _dev(SWO_array, torch.int16, "SWO_array")
    _call("hc_smooth_rivers_dev", burn.device, _tp(dem), _tp(burn), _tp(wm), _tp(swo), float(nodata), int(SWO_cut),
          float(np.float32(max_lower)), int(filter_size), burn.shape[0], burn.shape[1], _stream(burn))
    return _finish_inplace(DEM_burn_arr, burn)


def halo_rows(name, **kw):
    """Rows of halo a row band needs for `name` to give the whole-raster result on its own rows (the dependency
    radius of the chain: mask dilations + 5x5 mean + moving window), for bands.apply_banded / apply_rank."""
    if name == "river_minimize":
        return 1 + int(kw.get("river_filter_size", 3)) // 2
    if name == "raise_small_rivers":
        return int(kw.get("filter_size", 51)) // 2
    if name == "raise_small_streams":
        return 2 + int(kw.get("fsize", 51)) // 2
    if name == "smooth_rivers":
        return 2 + int(kw.get("filter_size", 21)) // 2
    raise KeyError(name)


def fill_missing_data(array, nodata, filter=11, weight=2.0):
    """wbt.fill_missing_data on an array (hc_fill_missing_data_dev): IDW infill of nodata cells from the valid cells that
    border nodata within filter / 2 cells (tools/water.py:1396, :1774).  numpy in -> numpy out, tensor in -> tensor out."""
    import torch
    host = not _is_torch(array)
    z = _dev(array, torch.float32, "fill_missing_data")
    out = torch.empty_like(z)
    _call("hc_fill_missing_data_dev", z.device, _tp(z), _tp(out), z.shape[0], z.shape[1], float(nodata), int(filter),
          float(weight), _stream(z))
    return out.cpu().numpy() if host else out


# ---- enforce_monotonic_rivers (tools/water.py:1407-1779) ---------------------------------------------------------------
def _grow(mask, k):
    """binary dilation of a small window by a k x k box (what apply_convolve_filter(mask, ones((k, k))) > 0 and
    ndimage.maximum_filter(mask, size=k) > 0 decide on a 0 / positive mask; symmetric padding adds nothing new)"""
    h = k // 2
    p = np.pad(mask.astype(bool), h, mode="constant")
    out = np.zeros(mask.shape, bool)
    for dr in range(k):
        for dc in range(k):
            out |= p[dr:dr + mask.shape[0], dc:dc + mask.shape[1]]
    return out


def _label4(mask):
    """4-connected components of a small window, scipy numbering (the library's labelling kernel)"""
    from .. import core
    if mask.size == 0 or not mask.any():
        return np.zeros(mask.shape, np.int32), 0
    lab, n = core.label(np.ascontiguousarray(mask, dtype=np.uint8), 4)
    return lab, int(n)


class _MonoLevels:
    """Device side of the level loop: the candidate mask of a water level, its 4-connected clumps and their per-clump
    counts / boundary flags / bounding boxes (hc_mono_mask_dev, hc_label_dev, hc_mono_stats_dev)."""

    def __init__(self, wm0, donut, inner):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.wm = torch.from_numpy(np.ascontiguousarray(wm0, np.uint8)).to(self.dev)
        self.donut = None if donut is None else torch.from_numpy(np.ascontiguousarray(donut == 1, np.uint8)).to(self.dev)
        self.inner = None if inner is None else torch.from_numpy(np.ascontiguousarray(inner == 1, np.uint8)).to(self.dev)
        self.mask = torch.empty(self.wm.shape, dtype=torch.uint8, device=self.dev)

    def clumps(self, dem, left, h):
        """-> (labels on the host or None, stats[nlab + 1, 6] on the host)"""
        from .. import core
        torch = self.torch
        d = torch.from_numpy(dem).to(self.dev)
        lf = torch.from_numpy(np.ascontiguousarray(left != 0, np.uint8)).to(self.dev)
        ny, nx = d.shape
        _call("hc_mono_mask_dev", self.dev, _tp(d), _tp(self.wm), _tp(lf), float(np.float32(h)), _tp(self.mask), ny, nx,
              _stream(d))
        lab, n = core.label(self.mask, 4)
        n = int(n)
        if n == 0:
            return None, np.zeros((1, 6), np.int32)
        st = torch.zeros((n + 1, 6), dtype=torch.int32, device=self.dev)
        st[:, 2] = 2 ** 31 - 1
        st[:, 4] = 2 ** 31 - 1
        st[:, 3] = -1
        st[:, 5] = -1
        _call("hc_mono_stats_dev", self.dev, _tp(lab), _tp(self.donut), _tp(self.inner), n, _tp(st), ny, nx, _stream(d))
        self.labels_dev = lab
        return lab, st.cpu().numpy()


def enforce_monotonic_rivers(DEM_tif, outpath, DEM_burn_arr, SWO_array, water_mask_array, donut_array, basename,
                             SWO_cutoff=20, radius=25, strict_endcap_geom=True, nodata=-9999, wbt=None, max_r_drop=100.0,
                             inner_boundary=None, verbose_print=False):
    """tools/water.py:1407-1779: lower river reaches that stand above the water downstream ("non-monotonic" sections of
    the flattened river classes 4..7).  For every water level h of a ladder, the river cells still above h are clumped
    (4-connectivity); a clump that is large enough, stays clear of the boundary donut and passes the reference's
    geometry tests (two river end caps far enough apart, not a narrow strip, not a bank sliver) is lowered onto the
    water next to it, and the cells around it are blanked and re-interpolated (wbt.fill_missing_data) at the end.

    The raster-wide work of a level -- mask, clumps, per-clump statistics -- runs on the GPU; the per-clump tests look at
    a window of the clump's bounding box on the host, with the same window arithmetic as the reference (numpy slicing,
    percentiles).  Same signature and return (DEM_burn_arr, water_mask_array), both modified in place; `DEM_tif` /
    `outpath` / `basename` name the two GeoTIFFs the interpolation step writes, as in the reference."""
    import os
    from .. import raster_io
    dem = DEM_burn_arr
    wm = water_mask_array
    if not (isinstance(dem, np.ndarray) and dem.dtype == np.float32 and isinstance(wm, np.ndarray)):
        raise TypeError("enforce_monotonic_rivers: float32 numpy DEM_burn_arr and a numpy water mask (modified in place)")
    min_area = np.round(np.pi * ((radius / 10.0) ** 2))
    river = (wm > 3) & (wm < 8)
    if not river.any():
        return dem, wm
    if inner_boundary is not None:
        vals, cnts = np.unique(np.round(dem[river & (inner_boundary == 1)]), return_counts=True)
        levels = vals[cnts > 100]
        step = None
    else:
        rz = dem[river]
        lo = np.floor(np.maximum(-3.0, np.percentile(rz, 1) - 1.0)) - 0.5
        hi = np.floor(np.percentile(rz, 99)) + 3.5
        step = 0.5
        for span, s in ((50, 1.0), (100, 3.0), (200, 5.0)):
            if (hi - lo) > span:
                step = s
        levels = np.arange(lo, hi, step)
        if len(levels) < 2 or int(river.sum()) < 3000:
            return dem, wm
    dem0 = dem.copy()
    wm0 = wm.copy()
    blank = np.zeros_like(wm)          # 1: bank cells to re-interpolate (become class 2), 2: water cells to re-interpolate
    left = np.zeros_like(SWO_array)
    left[river] = 1
    n_river = int(river.sum())
    gpu = _MonoLevels(wm0, donut_array, inner_boundary)
    any_blank = False
    for h in levels:
        if n_river < 100:               # (left is 1 or 2 on every river cell, so the reference's count is the river's size)
            break
        lab_dev, st = gpu.clumps(dem, left, h)
        if lab_dev is None:
            continue
        ids = np.flatnonzero((st[:, 0] > min_area) & (st[:, 1] == 0))     # big enough, clear of the boundary donut
        ids = ids[ids > 0]
        if len(ids) == 0:
            continue
        lab = lab_dev.cpu().numpy()
        diff = np.zeros_like(dem)
        diff[river] = dem[river] - h
        for cid in ids:
            # the clump's window: bounding box grown by 4 cells (5 at the far end), sliced the way numpy slices
            b_r0, b_r1, b_c0, b_c1 = (int(v) for v in st[cid, 2:6])
            r_lo, r_hi, c_lo, c_hi = b_r0 - 4, b_r1 + 5, b_c0 - 4, b_c1 + 5
            win = (slice(r_lo, r_hi), slice(c_lo, c_hi))
            body = lab[win] == cid
            sr, sc = np.nonzero(lab[b_r0:b_r1 + 1, b_c0:b_c1 + 1] == cid)
            cells = (sr + b_r0, sc + b_c0)                            # raster order, like np.nonzero on the whole array
            n_cells = len(sr)
            wm_w, swo_w, diff_w = wm0[win], SWO_array[win], diff[win]
            # the clump must touch river banks (classes 1, 2), unless it is big
            ring = _grow(body, 3) & ~body
            n_bank = int((ring & (wm_w < 3)).sum())
            if n_bank == 0 and n_cells < min_area * 4:
                continue
            # 3-cell buffer: river cells (classes 3..6) next to the clump, bank cells (1, 2) next to it
            buf = _grow(body, 7)
            near_river = buf & ~body & (wm_w > 2) & (wm_w < 7)
            if n_cells < 2 * int(near_river.sum()):
                continue                                             # a narrow strip between water bodies
            near_bank = buf & ~body & (wm_w < 3) & (wm_w > 0)
            if int(near_bank.sum()) < 1.5 * int(near_river.sum()):
                if np.percentile(swo_w[body], 80) < SWO_cutoff:      # a sliver along a bank, and not mostly water
                    continue
            # end caps: the adjacent river cells that are NOT above h, as 4-connected pieces
            caps = near_river & ~(diff_w > 0)
            cap_lab, n_caps = _label4(caps)
            if n_caps < 2:
                continue
            br, bc = np.nonzero(body)
            in_r_lo, in_r_hi, in_c_lo, in_c_hi = br.min() + 10, br.max() - 10, bc.min() + 10, bc.max() - 10
            if strict_endcap_geom:
                for k in range(1, n_caps + 1):
                    piece = cap_lab == k
                    pr, pc = np.nonzero(piece)
                    if len(pr) < 3:
                        caps[piece] = False                          # too small to be a river end
                        continue
                    if in_c_lo < pc.min() and in_c_hi > pc.max() and in_r_lo < pr.min() and in_r_hi > pr.max():
                        caps[piece] = False                          # deep inside the clump's box
                        continue
                    # (the reference's third test, "the piece's surroundings lie inside the clump", only ever writes
                    #  zeros onto cells that already are zero, :1651-1658: nothing to do)
                cap_lab, n_caps = _label4(caps)
            if n_caps > 3 or n_caps < 2:
                continue
            # the caps must lie far apart: extreme cap cells in x and in y, each pair in different pieces
            cr, cc = np.nonzero(cap_lab)
            ix0, ix1, iy0, iy1 = np.argmin(cc), np.argmax(cc), np.argmin(cr), np.argmax(cr)
            x_dist = np.hypot(float(cr[ix1] - cr[ix0]), float(cc[ix1] - cc[ix0]))
            y_dist = np.hypot(float(cr[iy1] - cr[iy0]), float(cc[iy1] - cc[iy0]))
            piece_of = cap_lab[cr, cc]
            if piece_of[ix0] == piece_of[ix1]:
                x_dist = 0
            if piece_of[iy0] == piece_of[iy1]:
                y_dist = 0
            diag = np.hypot(float(r_hi - r_lo), float(c_hi - c_lo))
            far = diag / 2 if strict_endcap_geom else diag / 10
            if x_dist < far or y_dist < far:
                continue
            cap_cells = (cr + r_lo, cc + c_lo)
            z_clump = dem[cells]
            z_lo, z_hi = np.percentile(z_clump, 10), np.percentile(z_clump, 95)
            z_water = np.percentile(dem[cap_cells], 20)
            drop = z_lo - z_water
            if drop < 0.0:
                continue
            wr, wc = np.nonzero(near_river)
            water_cells = (wr + r_lo, wc + c_lo)
            if (z_hi - z_lo) < max_r_drop and left[cells].min() == 2:
                # inside a section lowered before: straight onto the water next to it; banks around it are re-interpolated
                dem[cells] = z_water
                kr, kc = np.nonzero(near_bank)
                blank[water_cells] = 2
                blank[(kr + r_lo, kc + c_lo)] = 1
                wm[cells] = 7
            else:
                blank[water_cells] = 2
                wm[cells] = 6
                left[cells] = 2
                dem[cells] = dem[cells] - drop
                over = diff[cells] < drop
                dem[(cells[0][over], cells[1][over])] = z_water
            any_blank = True
            if verbose_print:
                print(f"       {cid}-raised river section lowered {drop:.2f} m")
            # nothing may end up above where it started (only the clump's cells were touched)
            up = (dem[cells] - dem0[cells]) > 0.1
            dem[(cells[0][up], cells[1][up])] = dem0[(cells[0][up], cells[1][up])]
    if any_blank:
        holes = np.nonzero(blank)
        work = dem.copy()
        work[holes] = nodata
        gap_tif = os.path.join(outpath, f"{basename}_non_mono.tif")
        fill_tif = os.path.join(outpath, f"{basename}_non_mono_interp.tif")
        meta = raster_io.read_raster(DEM_tif)[1] if (DEM_tif is not None and os.path.exists(str(DEM_tif))) else None
        if wbt is not None and meta is not None:
            raster_io.write_raster(gap_tif, work, meta, nodata=nodata)
            wbt.fill_missing_data(gap_tif, fill_tif, filter=17, weight=2)
            filled = np.asarray(raster_io.read_raster(fill_tif)[0], np.float32)
        else:
            filled = fill_missing_data(work, nodata, 17, 2.0)
        dem[holes] = filled[holes]
        wm[blank == 1] = 2
    return dem, wm
