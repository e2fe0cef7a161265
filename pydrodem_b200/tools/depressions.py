"""Mirror of the hot-path functions of the reference's `tools/depressions.py`.

Same signatures and file protocol (`<in>_WL.tif`, `<in>_WL_DIFF.tif`, clump-ID rasters); GDAL is
replaced by pydrodem_b200.raster_io, WhiteboxTools by the `wbt` duck type (pydrodem_b200.wbt) and
scipy.ndimage.label / convolve2d by the CUDA kernels.  build_pit_catalog / select_solution /
apply_solution / fill_shallow_pits run as segmented reductions on the GPU (csrc/segment.cu);
breach_pits runs through `wbt.breach_depressions` (flow-path breach on the GPU, hc_breach_flowpath_dev).
"""
import os

import numpy as np

from .. import core, raster_io, stencils


def _diff3(a, b):
    """np.round(a - b, 3): on the GPU for float32 rasters (bit-identical, core.diff_round), numpy's own otherwise"""
    if a.dtype == np.float32 and b.dtype == np.float32 and a.shape == b.shape and a.ndim == 2:
        return core.diff_round(a, b, 3)
    return np.round((a - b), 3)


def fill_pits(inDEM, wbt, fill_method='WL', nodataval=-9999, fixflats=False, flat_increment=0.0, save_tif=True,
              downcast=True):
    """tools/depressions.py:1236-1316.  Returns (fill_array, diff_fill_array), writes `<in>_<method>.tif`
    and (save_tif) `<in>_<method>_DIFF.tif`.  `downcast` only matters for the float64 epsilon fill (flat_increment > 0)."""
    fill_tif = inDEM.replace('.tif', '_{0}.tif'.format(fill_method))
    if flat_increment > 0.0:
        fixflats = True
    if fill_method == "SF":
        wbt.fill_depressions(inDEM, fill_tif, max_depth=None, fix_flats=fixflats, flat_increment=flat_increment)
    elif fill_method == "WL":
        wbt.fill_depressions_wang_and_liu(inDEM, fill_tif, fix_flats=fixflats, flat_increment=flat_increment)
    else:
        raise ValueError("   Improper fill method specified")
    inDEM_array, meta = raster_io.read_raster(inDEM)
    fill_array, fmeta = raster_io.read_raster(fill_tif)
    if downcast and fill_array.dtype == np.float64:
        # the epsilon fill is written in float64 (as WhiteboxTools writes everything); pyct.tifdowncast narrows it (:1306)
        fill_array = fill_array.astype(np.float32)
        raster_io.write_raster(fill_tif, fill_array, fmeta, nodata=nodataval)
    diff_fill_array = _diff3(fill_array, inDEM_array)  # :1308
    if save_tif:
        diff_tif_fill = fill_tif.replace('.tif', '_DIFF.tif')
        raster_io.write_raster(diff_tif_fill, diff_fill_array.astype(np.float32), meta, nodata=nodataval)
    return fill_array, diff_fill_array


def breach_pits(inDEM_tif, radius, maxcost, wbt, min_dist=False, flat_increment=0.0, breach_increment=0.0001,
                nodataval=-9999, method='SC', fixflats=False, save_tif=True):
    """tools/depressions.py:1319-1403.  Returns (carve_array, diff_carve_array); writes `<in>_<method>_<radius>.tif`
    (re-used when it exists, :1369) and (save_tif) its `_DIFF.tif`."""
    carve_path = inDEM_tif.replace('.tif', '_{1}_{0}.tif'.format(radius, method))
    carvefill_path = inDEM_tif.replace('.tif', '_{1}_{0}_filled.tif'.format(radius, method))
    if flat_increment > 0.0:
        fixflats = True
    if not os.path.exists(carve_path):
        if method == 'LC':
            wbt.breach_depressions_least_cost(inDEM_tif, carve_path, radius, max_cost=maxcost, min_dist=min_dist,
                                              flat_increment=breach_increment, fill=False)
        else:
            wbt.breach_depressions(inDEM_tif, carve_path, max_length=radius, max_depth=None, fill_pits=True,
                                   callback=None, flat_increment=flat_increment)
    if method == 'LC':
        wbt.fill_depressions_wang_and_liu(carve_path, carvefill_path, fix_flats=fixflats, flat_increment=flat_increment)
    else:
        carvefill_path = carve_path
    inDEM_array, meta = raster_io.read_raster(inDEM_tif)
    carve_array, _ = raster_io.read_raster(carvefill_path)
    diff_carve_array = _diff3(carve_array, inDEM_array)  # :1395
    if save_tif:
        diff_tif_carve = carvefill_path.replace('.tif', '_DIFF.tif')
        raster_io.write_raster(diff_tif_carve, diff_carve_array.astype(np.float32), meta, nodata=nodataval)
    return carve_array, diff_carve_array


def _sign_mask(values, condition, extra=None):
    """uint8 device mask of a difference map's sign: 'greaterthan' (> 0), 'lessthan' (< 0, or `extra` == 1), 'nonzero'"""
    import torch
    v = values if core._is_torch(values) else torch.from_numpy(np.ascontiguousarray(values)).cuda()
    if condition == 'greaterthan':
        m = v > 0
    elif condition == 'lessthan':
        m = v < 0
        if extra is not None:
            e = extra if core._is_torch(extra) else torch.from_numpy(np.ascontiguousarray(extra)).cuda()
            m = m | (e == 1)
    elif condition == 'nonzero':
        m = v != 0
    else:
        m = torch.zeros_like(v, dtype=torch.bool)
    return m


def create_clump_array(input_array, clump_ID_tif, stencil_tif, wbt, nodataval=-9999, noflow_array=None,
                       condition='greaterthan', clumpbuffer=None, split_mask=None, save_tif=True):
    """tools/depressions.py:1593-1675: clumps of a difference map's sign, ids in scipy's raster-scan order.
    One device pipeline: sign predicate -> (clumpbuffer) square dilation as a (2b+1)-wide moving maximum
    (`hc_minmax_filter_dev`; equal to the reference's box convolution + "> 0") -> labels (`hc_label_dev`: 4-connected
    when dilated, 8-connected otherwise).  A `split_mask` is labelled on its own and numbered after the rest.  Only the
    finished id raster crosses to the host."""
    import torch
    from . import water as _water   # (the moving-maximum binding lives with the water stencils)
    mask = _sign_mask(np.asarray(input_array) if not core._is_torch(input_array) else input_array, condition, noflow_array)
    mask = mask.to(torch.uint8).contiguous()
    conn = 8
    if clumpbuffer is not None:
        mask = _water.maximum_filter(mask, 2 * int(clumpbuffer) + 1)
        conn = 4
    if split_mask is None:
        ids, _ = core.label(mask, conn)
    else:
        inside = torch.from_numpy(np.ascontiguousarray(np.asarray(split_mask) != 0)).to(mask.device)
        part = mask * inside.to(torch.uint8)
        rest = mask * (~inside).to(torch.uint8)
        ids, n_rest = core.label(rest.contiguous(), conn)
        ids_part, _ = core.label(part.contiguous(), conn)
        ids = torch.where(ids_part > 0, ids_part + n_rest, ids)
    clump_ID_array = ids.cpu().numpy()
    if save_tif:
        meta = raster_io.read_raster(stencil_tif)[1] if (stencil_tif and os.path.exists(stencil_tif)) else None
        raster_io.write_raster(clump_ID_tif, clump_ID_array, meta, nodata=0)
    return clump_ID_array


def _valid_bbox(valid):
    """(first row, last row, first col, last col) of a boolean raster's True cells, or None"""
    rows = np.flatnonzero(valid.any(axis=1))
    cols = np.flatnonzero(valid.any(axis=0))
    if rows.size == 0:
        return None
    return int(rows[0]), int(rows[-1]), int(cols[0]), int(cols[-1])


def fix_shifted_data(in_array, shift_array, nodataval, failtest='negative_diff', vprint=False):
    """tools/depressions.py:1406-1468: a WhiteboxTools output that came back displaced by whole pixels is moved back
    (the displacement is read off the valid cells' bounding boxes: lower-right corners, as upstream) and the rounded
    difference map is rebuilt.  Returns (shift_array, diff_array); `shift_array` is repaired in place.  The GPU tools
    never displace a raster, so here the displacement is (0, 0) and the call is the difference recompute plus the
    upstream sanity check (at most `|valid-count difference| + 3` cells may fail `failtest`)."""
    have, got = in_array != nodataval, shift_array != nodataval
    bb_in, bb_out = _valid_bbox(have), _valid_bbox(got)
    if bb_in is not None and bb_out is not None:
        dr, dc = bb_out[1] - bb_in[1], bb_out[3] - bb_in[3]
        if dr or dc:
            moved = np.full_like(shift_array, nodataval)
            ny, nx = shift_array.shape
            # destination (r, c) <- source (r + dr, c + dc), wherever both lie inside the raster
            r0, r1 = max(0, -dr), min(ny, ny - dr)
            c0, c1 = max(0, -dc), min(nx, nx - dc)
            moved[r0:r1, c0:c1] = shift_array[r0 + dr:r1 + dr, c0 + dc:c1 + dc]
            keep = (moved != nodataval)
            shift_array[keep] = moved[keep]
            if vprint:
                print("   shifted back by", (dr, dc))
    shift_array[~have] = nodataval
    diff_array = np.round(shift_array - in_array, 3)
    bad = diff_array < 0.0 if failtest == 'negative_diff' else diff_array > 0.0
    slack = abs(int(have.sum()) - int(got.sum())) + 3
    assert int(bad.sum()) <= slack, 'Re-mapped fill diff has too many negative values'
    return shift_array, diff_array


# ==================================================================================================
# pit catalog / solution selection (tools/depressions.py:272-810, 1140-1228, 1471-1531) on the GPU:
# every per-pit Python loop of the reference is a segmented reduction keyed by the label rasters
# (csrc/segment.cu).  Same signatures and return shapes; the per-pit index dictionaries are lazy
# views (a pit's index list is materialised only when someone subscripts it).
# ==================================================================================================
import ctypes as _ct

from .. import _lib as _L


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise _L.HydrocoreError("the pit catalog needs a CUDA device (no CPU path)")
    return torch


def _dev(a, dtype):
    """1-D contiguous device tensor of `dtype` from a numpy array / torch tensor (flattened)."""
    torch = _torch()
    if type(a).__module__.startswith("torch"):
        t = a.reshape(-1)
        return t.to(device="cuda", dtype=dtype).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.ravel(a))).to(device="cuda").to(dtype).contiguous()


def _p(t):
    return _ct.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return _ct.c_void_p(_torch().cuda.current_stream().cuda_stream)


class LazyCells:
    """dict-like `{label: uint32 index array}` computed on demand from a label raster on the device
    (the reference materialises every list up front: tools/depressions.py:463-469)."""

    def __init__(self, kind, state):
        self.kind, self.state = kind, state

    def __len__(self):
        return self.state["ncarves" if self.kind == "carve_only" else "npits"]

    def keys(self):
        return range(1, len(self) + 1)

    def __contains__(self, k):
        return 1 <= int(k) <= len(self)

    def __getitem__(self, k):
        torch = _torch()
        s, k = self.state, int(k)
        if k not in self:
            raise KeyError(k)
        if self.kind == "fill":            # fillcells_dct[pit]
            idx = torch.nonzero(s["fillID"] == k).reshape(-1)
        elif self.kind == "carve_only":    # carvecells_only_dct[carve]
            idx = torch.nonzero(s["carve_bound"] == k).reshape(-1)
        elif self.kind == "fillbehind":    # carve_fillcells_dct[pit]
            if bool(s["pit_out"][k]) or int(s["ncarvecells"][k]) == 0:
                return np.empty(0, np.uint32)
            idx = torch.nonzero((s["fillID"] == k) & (s["carvefillID"] > 0)).reshape(-1)
        else:                              # carvecells_dct[pit]: its included carves, in ascending carve id
            if bool(s["pit_out"][k]):
                return np.empty(0, np.uint32)
            pr = s["pairs"][: s["npairs"]]
            sel = ((pr >> 32) == k) & (s["included"][: s["npairs"]] != 0)
            carves = torch.sort(pr[sel] & 0xFFFFFFFF).values
            parts = [torch.nonzero(s["carve_bound"] == int(c)).reshape(-1) for c in carves.tolist()]
            idx = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64, device="cuda")
        return idx.cpu().numpy().astype(np.uint32)


def build_pit_catalog(inDEM_vec, nodataval, carve_vec, diff_vec_carve, diff_vec_fill, fillID_vec, carveID_vec,
                      carvefillID_vec, verbose_print=True, carveout=False):
    """tools/depressions.py:272-509.  Returns (DFpits, fillcells_dct, carvecells_only_dct, carvecells_dct,
    carve_fillcells_dct) like the reference; the dictionaries are LazyCells views sharing one device state.
    Float sums (volfill, volcarve, volfillbehind) are float64-accumulated: 1e-5 relative to the reference."""
    import pandas as pd
    torch = _torch()
    lib = _L.load()
    f32, i32, i64 = torch.float32, torch.int32, torch.int64
    dem, carve = _dev(inDEM_vec, f32), _dev(carve_vec, f32)
    dcarve, dfill = _dev(diff_vec_carve, f32), _dev(diff_vec_fill, f32)
    fid, cid, cfid = _dev(fillID_vec, i32), _dev(carveID_vec, i32), _dev(carvefillID_vec, i32)
    n = dem.numel()
    carveout_id = 0
    if not carveout:
        nd = torch.nonzero(dem == float(nodataval)).reshape(-1)
        if nd.numel() == 0:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # as the reference (:335)
        carveout_id = int(cid[nd[0]])
    npits, ncarves = int(fid.max()), max(int(cid.max()), 0)
    P1, C1 = npits + 1, ncarves + 1
    dv = "cuda"
    st = {"npits": npits, "ncarves": ncarves, "fillID": fid, "carvefillID": cfid,
          "carve_bound": torch.empty(n, dtype=i32, device=dv)}
    t = {k: torch.empty(P1, dtype=f32, device=dv) for k in ("maxfill", "volfill", "maxcarve", "volcarve", "volfb", "maxfb")}
    nfill = torch.empty(P1, dtype=i64, device=dv)
    ncarvecells = torch.empty(P1, dtype=i64, device=dv)
    complete = torch.empty(P1, dtype=torch.int8, device=dv)
    pit_out = torch.empty(P1, dtype=torch.uint8, device=dv)
    camax, camin = torch.empty(C1, dtype=i64, device=dv), torch.empty(C1, dtype=i64, device=dv)
    cmin = torch.empty(C1, dtype=f32, device=dv)
    csum = torch.empty(C1, dtype=torch.float64, device=dv)
    cn = torch.empty(C1, dtype=i64, device=dv)
    cap = max(1024, 4 * (npits + ncarves) + 1024)
    npairs = _ct.c_int64(0)
    ctx = _L.context(torch.cuda.current_device())
    while True:
        pairs = torch.empty(cap, dtype=i64, device=dv)
        inc = torch.zeros(cap, dtype=torch.uint8, device=dv)
        rc = lib.hc_pit_catalog_dev(ctx, _p(dem), float(nodataval), _p(carve), _p(dcarve), _p(dfill), _p(fid), _p(cid),
                                    _p(cfid), n, npits, ncarves, int(bool(carveout)), carveout_id, _p(st["carve_bound"]),
                                    _p(t["maxfill"]), _p(t["volfill"]), _p(nfill), _p(t["maxcarve"]), _p(t["volcarve"]),
                                    _p(ncarvecells), _p(t["volfb"]), _p(t["maxfb"]), _p(complete), _p(pit_out), _p(camax),
                                    _p(camin), _p(cmin), _p(csum), _p(cn), _p(pairs), cap, _p(inc), _ct.byref(npairs),
                                    _stream())
        if rc == -4 and cap < (1 << 28):   # HC_ERR_TOO_LARGE: more distinct pairs than the table holds
            cap *= 8
            continue
        _L.check(rc, "hc_pit_catalog_dev")
        break
    st.update(pairs=pairs, included=inc, npairs=int(npairs.value), pit_out=pit_out, ncarvecells=ncarvecells, nfill=nfill)
    cols = {k: t[k][1:].cpu().numpy() for k in t}
    DFpits = pd.DataFrame({'fillcarveID': np.arange(1, npits + 1), 'maxfill': cols["maxfill"], 'volfill': cols["volfill"],
                           'maxcarve': cols["maxcarve"], 'volcarve': cols["volcarve"], 'volfillbehind': cols["volfb"],
                           'maxfillbehind': cols["maxfb"], 'completecarve': complete[1:].cpu().numpy()})
    DFpits.index = DFpits['fillcarveID']
    DFpits.drop(labels=['fillcarveID'], axis=1, inplace=True)
    return (DFpits, LazyCells("fill", st), LazyCells("carve_only", st), LazyCells("carve", st),
            LazyCells("fillbehind", st))


def _col(DFpits, name, dtype):
    torch = _torch()
    a = np.concatenate([[0], DFpits[name].to_numpy()])
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(dtype).contiguous()


def _state_of(*objs):
    for o in objs:
        if isinstance(o, LazyCells):
            return o.state
    raise TypeError("select_solution / apply_solution need the dictionaries returned by this module's build_pit_catalog")


def select_solution(DFpits, fillcells_dct, carvecells_dct, carvecells_only_dct, fill_vec, carve_vec, fillID_vec,
                    carveID_vec, fill_vol_thresh, carve_vol_thresh, max_carve_depth, maxcarve_factor=1.25,
                    apply_shallow_fill=False, wall_array=None, water_mask_array=None, sink_array=None,
                    numpy_scalars=None):
    """tools/depressions.py:517-810: adds the 'solution' and 'downstream_pit' columns.  Like the reference, a
    missing water mask is an error (it raises NameError at :647).

    numpy_scalars: 'weak' evaluates the tree in float32, as numpy >= 2 does for the reference's mix of a
    float32 catalog row and Python floats; 'legacy' in float64 (numpy < 2, the reference's own era).  The two
    differ only at exact threshold hits (maxfill == float32(0.01) is "< 0.01" in float64 only).  Default: the
    rule of the numpy installed here, i.e. what the reference itself would do in this environment."""
    if numpy_scalars is None:
        numpy_scalars = "weak" if int(np.__version__.split(".")[0]) >= 2 else "legacy"
    if numpy_scalars not in ("weak", "legacy"):
        raise ValueError("numpy_scalars must be 'weak' or 'legacy'")
    if water_mask_array is None:
        raise NameError("name 'water_mask_vec' is not defined")  # the reference's latent bug, kept
    torch = _torch()
    s = _state_of(fillcells_dct, carvecells_dct, carvecells_only_dct)
    lib = _L.load()
    f32 = torch.float32
    npits, ncarves = s["npits"], s["ncarves"]
    n = s["fillID"].numel()
    wall = _dev(wall_array, f32) if wall_array is not None else None
    water = _dev(water_mask_array, f32)
    sink = _dev(sink_array, f32) if sink_array is not None else None
    sol = torch.zeros(npits + 1, dtype=torch.int32, device="cuda")
    cols = [_col(DFpits, c, f32) for c in ("maxfill", "volfill", "maxcarve", "volcarve", "volfillbehind", "maxfillbehind")]
    comp = _col(DFpits, "completecarve", torch.int8)
    _L.check(lib.hc_pit_select_dev(_L.context(torch.cuda.current_device()), _p(s["fillID"]), _p(s["carve_bound"]), n, npits,
                                   ncarves, *[_p(c) for c in cols], _p(comp), _p(s["nfill"]), _p(s["ncarvecells"]),
                                   _p(s["pit_out"]), _p(s["pairs"]), _p(s["included"]), s["npairs"], _p(wall), _p(water),
                                   _p(sink), float(fill_vol_thresh), float(carve_vol_thresh), float(max_carve_depth),
                                   float(maxcarve_factor), int(bool(apply_shallow_fill)),
                                   1 if numpy_scalars == "weak" else 0, _p(sol), _stream()),
              "hc_pit_select_dev")
    DFpits["solution"] = sol[1:].cpu().numpy().astype(np.int64)
    DFpits["downstream_pit"] = np.zeros(npits, dtype=np.int64)
    return DFpits


def apply_solution(DFpits, fillcells_dct, carvecells_dct, carve_fillcells_dct, inDEM_vec, fill_vec, carve_vec,
                   solution_mask_vec=None, excludedpits=None):
    """tools/depressions.py:1140-1228.  Returns (modDEMoutput_vec float32, solution_mask_vec int16) as numpy."""
    torch = _torch()
    s = _state_of(fillcells_dct, carvecells_dct, carve_fillcells_dct)
    lib = _L.load()
    f32 = torch.float32
    dem, fillv, carvev = _dev(inDEM_vec, f32), _dev(fill_vec, f32), _dev(carve_vec, f32)
    n = dem.numel()
    npits, ncarves = s["npits"], s["ncarves"]
    # the reference loops over DFpits.index only (callers pass DFpits[DFpits.solution < 3000]): absent pits = -1
    solh = np.full(npits + 1, -1, np.int32)
    solh[DFpits.index.to_numpy().astype(np.int64)] = DFpits["solution"].to_numpy().astype(np.int32)
    sol = torch.from_numpy(solh).cuda()
    if int(((sol == 1023)).sum()) != 0:
        raise NotImplementedError("solution 1023 (fill to downstream pit level) is dead code upstream and not built")
    mask = _dev(solution_mask_vec, torch.int16) if solution_mask_vec is not None else \
        torch.zeros(n, dtype=torch.int16, device="cuda")
    excl = None
    if excludedpits is not None:
        excl = torch.zeros(npits + 1, dtype=torch.uint8, device="cuda")
        if len(excludedpits):
            excl[torch.as_tensor(list(excludedpits), dtype=torch.int64, device="cuda")] = 1
    out = torch.empty_like(dem)
    _L.check(lib.hc_pit_apply_dev(_L.context(torch.cuda.current_device()), _p(dem), _p(fillv), _p(carvev), _p(s["fillID"]),
                                  _p(s["carve_bound"]), _p(s["carvefillID"]), n, npits, ncarves, _p(sol), _p(excl),
                                  _p(s["pit_out"]), _p(s["ncarvecells"]), _p(s["pairs"]), _p(s["included"]), s["npairs"],
                                  _p(out), _p(mask), _stream()), "hc_pit_apply_dev")
    return out.cpu().numpy(), mask.cpu().numpy()


def fill_shallow_pits(modDEM_vec, sfill, fillID_vec, diff_fill_array, fill_array, water_vec=None, verbose_print=True):
    """tools/depressions.py:1471-1531: every label (background included, as np.isin does) without a cell deeper
    than sfill and without a water-mask cell takes the filled values.  Returns the modified 1-D float32 array."""
    torch = _torch()
    lib = _L.load()
    f32 = torch.float32
    mod = _dev(modDEM_vec, f32).clone()
    fid = _dev(fillID_vec, torch.int32)
    diff, fillv = _dev(diff_fill_array, f32), _dev(fill_array, f32)
    water = _dev(water_vec, f32) if water_vec is not None else None
    npits = int(fid.max())
    nsh = _ct.c_int64(0)
    _L.check(lib.hc_fill_shallow_pits_dev(_L.context(torch.cuda.current_device()), _p(mod), float(sfill), _p(fid), _p(diff),
                                          _p(fillv), _p(water), mod.numel(), npits,
                                          _ct.byref(nsh) if verbose_print else None, _stream()), "hc_fill_shallow_pits_dev")
    out = mod.cpu().numpy()
    if isinstance(modDEM_vec, np.ndarray) and modDEM_vec.dtype == np.float32 and modDEM_vec.ndim == 1:
        modDEM_vec[:] = out            # the reference modifies its argument in place and returns it
        return modDEM_vec
    return out


def raise_lake_connected_pits(watermask, fillID_vec, modDEM_vec, inDEM_vec, starttime=None, verbose_print=True):
    """tools/depressions.py:1534-1590 as its docstring and comments describe it: every pit that holds large-waterbody
    cells (`watermask` raster == 1) is raised, where it lies below the lowest water cell inside the pit, to that level
    + 1 cm.  Returns (modDEM_vec, DEMdiff_vec = inDEM_vec - modDEM_vec).  Nothing happens with 20 or fewer lake cells.

    Upstream computes the lake/pit intersection as `np.intersect1d(np.where(water_array == 1), flat indices)`, i.e. it
    intersects the ROW and COLUMN numbers of the lake cells with flat cell indices, so its result is unrelated to the
    lake (and the option, `fill_h20_adjacent`, is off in every shipped configuration).  This is the documented
    behaviour, not that accident: one segmented minimum per pit over its lake cells (`hc_seg_stats_dev`), one
    elementwise raise.  `watermask`: a GeoTIFF path, or the 2-D mask itself."""
    torch = _torch()
    water = raster_io.read_raster(watermask)[0] if isinstance(watermask, str) else np.asarray(watermask)
    lake = _dev(water == 1, torch.int32)
    dem_in = _dev(inDEM_vec, torch.float32)
    n = dem_in.numel()
    mod_is_torch = type(modDEM_vec).__module__.startswith("torch")
    mod = _dev(modDEM_vec, torch.float32).clone()
    if int(lake.sum()) > 20:
        fid = _dev(fillID_vec, torch.int32)
        npits = int(fid.max()) if n else 0
        if npits > 0:
            lowest = torch.full((npits + 1,), float("inf"), dtype=torch.float32, device=dem_in.device)
            cnt = torch.zeros(npits + 1, dtype=torch.int64, device=dem_in.device)
            _L.check(_L.load().hc_seg_stats_dev(_L.context(dem_in.device.index or 0), _p(fid), _p(dem_in), _p(lake), n, npits, 1,
                                                None, _p(lowest), None, _p(cnt), _stream()), "hc_seg_stats_dev")
            lowest[cnt == 0] = float("-inf")          # pits without lake cells: nothing lies below
            lowest[0] = float("-inf")
            lvl = lowest[fid.long()]
            raise_it = (fid > 0) & (dem_in < lvl)
            mod = torch.where(raise_it, lvl + 0.01, mod)
    diff = dem_in - mod
    if verbose_print:
        print(f"      number of cells adjusted: {int((diff < 0).sum())} Max change: {float(diff.abs().max()) if n else 0.0:.2f}")
    if mod_is_torch:
        return mod, diff
    return mod.cpu().numpy(), diff.cpu().numpy()


# ==================================================================================================
# initial_fill_carve (tools/depressions.py:41-264): the orchestration that turns one DEM GeoTIFF into the eight
# vectors the pit catalog consumes -- two fills, four clump rasters, one breach.
# ==================================================================================================
def initial_fill_carve(inDEM_tif, basename, outpath, fill_method, maxcost, radius, min_dist, nodataval, starttime, wbt,
                       sfill=None, flat_increment=0.0, watermask=None, water_mask_array=None, carveout=False,
                       carve_method='SC', verbose_print=True, del_temp_files=False):
    """Same signature, file names and return tuple as the reference:
    (modDEM_vec, fill_vec, carve_vec, diff_vec_fill, diff_vec_carve, fillID_vec, carveID_vec, carvefillID_vec, nx, ny).
    `watermask` (path or array of the large-waterbody mask) switches on the lake raise of :163-176, see
    raise_lake_connected_pits."""
    import glob
    inDEM_array, meta = raster_io.read_raster(inDEM_tif)
    ny, nx = inDEM_array.shape
    water_mask_vec = np.ravel(water_mask_array) if water_mask_array is not None else None

    # initial fill: finds lake-draining and shallow pits (:103-117)
    fill_array, diff_fill_array = fill_pits(inDEM_tif, wbt, fill_method=fill_method, nodataval=nodataval,
                                            flat_increment=0.0)
    if np.amin(diff_fill_array) < 0.0:
        fill_array, diff_fill_array = fix_shifted_data(inDEM_array, fill_array, nodataval)
    fillID_path = os.path.join(outpath, '{0}_pitID_initial.tif'.format(basename))
    fillID_array = create_clump_array(diff_fill_array, fillID_path, inDEM_tif, wbt, nodataval=nodataval,
                                      condition='greaterthan')
    fillID_vec = fillID_array.reshape(ny * nx,)
    inDEM_vec = inDEM_array.reshape(ny * nx,)
    modDEM_vec = np.copy(inDEM_vec)
    modDEM_vec[np.where(np.isnan(modDEM_vec))] = nodataval

    if sfill is not None:                                                                     # :133-138
        modDEM_vec = fill_shallow_pits(modDEM_vec, sfill, fillID_vec, diff_fill_array, fill_array,
                                       water_vec=water_mask_vec, verbose_print=verbose_print)

    if watermask is not None:                                                                 # :163-180
        modDEM_vec, DEMdiff_lake_vec = raise_lake_connected_pits(watermask, fillID_vec, modDEM_vec, inDEM_vec,
                                                                 starttime=starttime, verbose_print=verbose_print)
        raster_io.write_raster(os.path.join(outpath, f"{basename}_lakefill_DIFF.tif"),
                               np.asarray(DEMdiff_lake_vec, np.float32).reshape(ny, nx), meta, nodata=nodataval)

    if sfill is None and watermask is None:                                                   # :183-186
        modDEM_tif = inDEM_tif
    else:                                                                                     # :187-199
        modDEM_array = np.asarray(modDEM_vec).reshape(ny, nx)
        modDEM_tif = os.path.join(outpath, f"{basename}_shallowfill.tif")
        raster_io.write_raster(modDEM_tif, modDEM_array.astype(np.float32), meta, nodata=nodataval)
        fill_array, diff_fill_array = fill_pits(modDEM_tif, wbt, fill_method=fill_method, nodataval=nodataval,
                                                flat_increment=flat_increment)
        if np.amin(diff_fill_array) < 0.0:
            fill_array, diff_fill_array = fix_shifted_data(modDEM_array, fill_array, nodataval)

    fillID_path = os.path.join(outpath, '{0}_pitID.tif'.format(basename))                     # :187-192
    fillID_array = create_clump_array(diff_fill_array, fillID_path, inDEM_tif, wbt, nodataval=nodataval,
                                      condition='greaterthan', clumpbuffer=1)
    fillID_vec = fillID_array.reshape(ny * nx,)
    fill_vec = fill_array.reshape(ny * nx,)
    diff_vec_fill = diff_fill_array.reshape(ny * nx,)

    carve_array, diff_carve_array = breach_pits(modDEM_tif, radius, maxcost, wbt, min_dist=min_dist,   # :199-205
                                                nodataval=nodataval, method=carve_method,
                                                flat_increment=flat_increment)
    carve_vec = carve_array.reshape(ny * nx,)
    diff_vec_carve = np.copy(diff_carve_array).reshape(ny * nx,)

    carveID_path = os.path.join(outpath, '{0}_carveID.tif'.format(basename))                  # :210-226
    carve_fillID_path = os.path.join(outpath, '{0}_carve_fillID.tif'.format(basename))
    if not carveout:   # nodata pixels join the carve clumps, so that carves leaving the raster can be told apart
        diff_carve_array[inDEM_array == nodataval] = -1.0
    carveID_array = create_clump_array(diff_carve_array, carveID_path, inDEM_tif, wbt, nodataval=nodataval,
                                       condition='lessthan')
    carvefillID_array = create_clump_array(diff_carve_array, carve_fillID_path, inDEM_tif, wbt, nodataval=nodataval,
                                           condition='greaterthan')
    carveID_vec = carveID_array.reshape(ny * nx,)
    carvefillID_vec = carvefillID_array.reshape(ny * nx,)

    if del_temp_files:                                                                        # :236-238 (ft.erase_files)
        raster_io.write_behind().flush()     # (queued writes of the files about to be erased must have landed)
        basename_mod = os.path.splitext(os.path.basename(modDEM_tif))[0]
        for f in glob.glob(os.path.join(outpath, '{0}_*.tif'.format(basename_mod))):
            if os.path.isfile(f):
                os.remove(f)
    return modDEM_vec, fill_vec, carve_vec, diff_vec_fill, diff_vec_carve, fillID_vec, carveID_vec, carvefillID_vec, \
        nx, ny


# ==================================================================================================
# combined_solution (tools/depressions.py:813-1137): for one big pit, try partial fills at fractions of its depth,
# re-breach each inside a window around the pit, and keep the cheapest complete result.
# ==================================================================================================
def _window_meta(meta, row_min, col_min):
    """geo-referencing of a sub-window (what pyct.clip_to_window derives from the geotransform, tools/convert.py:590-599)"""
    m = raster_io.RasterMeta(meta or {})
    geo = dict(m.get("geo", {}))
    tie, scale = geo.get(33922), geo.get(33550)
    if tie and scale:
        v = list(tie[1])
        v[3] = v[3] + col_min * scale[1][0]
        v[4] = v[4] - row_min * scale[1][1]
        geo[33922] = (tie[0], v)
    m["geo"] = geo
    return m


class _Trial:
    """one (pit, fill fraction) candidate of the combined solution: a partially filled window waiting for its breach"""
    __slots__ = ("pit", "frac", "win", "surface", "slot")

    def __init__(self, pit, frac, win, surface):
        self.pit, self.frac, self.win, self.surface, self.slot = pit, frac, win, surface, None


def _pack_shelves(shapes, max_width=4096, gutter=2):
    """greedy shelf packing of rectangles (h, w) with `gutter` free cells around each -> (positions, H, W)"""
    order = sorted(range(len(shapes)), key=lambda i: -shapes[i][0])
    W = max(max_width, max(w for _, w in shapes) + 2 * gutter)
    pos = [None] * len(shapes)
    x, y, shelf_h = gutter, gutter, 0
    for i in order:
        h, w = shapes[i]
        if x + w + gutter > W:
            x, y, shelf_h = gutter, y + shelf_h + gutter, 0
        pos[i] = (y, x)
        x += w + gutter
        shelf_h = max(shelf_h, h)
    return pos, y + shelf_h + gutter, W


def combined_solutions(pit_ids, inDEM_array, nx, ny, DFpits, fillcells_dct, carvecells_dct, carve_fillcells_dct,
                       combined_fill_interval, fillID_vec, fill_vec, carve_vec, radius, max_carve_depth, nodataval=-9999,
                       carve_method='SC', maxcost=None, min_dist=False, flat_increment=0.0):
    """The partial-fill + re-breach search of tools/depressions.py:813-1137 for MANY pits at once.

    For every pit the reference clips a window (pit bounding box + radius + 2 cells), lowers the pit's full fill to
    the fractions 0.2 ... 0.8 of its depth, breaches each partial surface again (max_length = radius, max_depth =
    max_carve_depth, no single-cell-pit fill), keeps the carves that reach the pit without running off the raster and
    without exceeding max_carve_depth, and finally picks -- among those, the plain fill and (if complete) the plain
    carve -- the candidate with the smallest absolute volume change.  Up to 4 WhiteboxTools subprocesses and 8 GeoTIFF
    round trips per pit upstream.

    Here all partial surfaces of all pits are packed into ONE mosaic raster with nodata gutters and go through ONE
    device breach and ONE device labelling: a gutter is exterior nodata, so every window's rim cells are seeds exactly
    as the rim of a stand-alone raster is, and no flow path, flat or clump can cross it -- each window's result is the
    stand-alone result.  carve_method='LC' re-breaches with the least-cost breach instead (:996-1000: dist = radius,
    max_cost = maxcost, min_dist, flat_increment, no fill).  Returns {pit_id: (solution_name, sol_df)} with the
    reference's return value per pit."""
    import pandas as pd
    import torch
    ground_full = np.asarray(inDEM_array, dtype=np.float32).reshape(ny, nx)
    fill_full = np.asarray(fill_vec).reshape(ny, nx)
    carve_full = np.asarray(carve_vec).reshape(ny, nx)
    ids_full = np.asarray(fillID_vec).reshape(ny, nx)
    nd = np.float32(nodataval)
    pad = int(radius) + 2
    fractions = np.round(np.arange(combined_fill_interval, 1, combined_fill_interval), 1)

    trials, per_pit = [], {}
    for pit in pit_ids:
        row = DFpits.loc[pit]
        cells = np.asarray(fillcells_dct[pit])
        rr, cc = np.divmod(cells, nx)
        win = (slice(int(max(0, rr.min() - pad)), int(min(ny, rr.max() + pad + 1))),
               slice(int(max(0, cc.min() - pad)), int(min(nx, cc.max() + pad + 1))))
        ground = np.ascontiguousarray(ground_full[win])
        inpit = ids_full[win] == pit
        per_pit[pit] = dict(row=row, cells=cells, win=win, ground=ground, inpit=inpit, found=[])
        for frac in fractions:
            # the full fill, lowered inside the pit by a constant (its gradient, if any, survives), never below ground
            surf = np.array(fill_full[win], copy=True)
            surf[inpit] = surf[inpit] - row.maxfill * (1 - frac)
            rise = np.round(surf - ground, 3)
            under = rise < 0
            surf[under] = ground[under]
            wet = rise > 0
            if not wet.any():
                continue
            # a partial fill that stands above the existing carve's back-fill cannot improve on that carve
            if row.completecarve == 1 and surf[wet].min() > carve_full[win][wet].min():
                continue
            trials.append(_Trial(pit, frac, win, surf.astype(np.float32)))

    if trials:
        shapes = [t.surface.shape for t in trials]
        pos, H, W = _pack_shelves(shapes)
        mosaic = np.full((H, W), nd, np.float32)
        for t, (y, x) in zip(trials, pos):
            h, w = t.surface.shape
            t.slot = (slice(y, y + h), slice(x, x + w))
            mosaic[t.slot] = t.surface
        mos_dev = torch.from_numpy(mosaic).cuda()
        if carve_method == 'LC':
            carved_dev = core.breach_depressions_least_cost(mos_dev, int(radius), nodata=float(nodataval), max_cost=maxcost,
                                                            min_dist=min_dist, flat_increment=flat_increment, fill=False)
        else:
            carved_dev = core.breach_depressions(mos_dev, nodata=float(nodataval), max_length=radius,
                                                 max_depth=max_carve_depth, fill_pits=False)
        carved = carved_dev.cpu().numpy()
        # clumps of lowered cells (plus the windows' own nodata: the raster edge a carve may not run into)
        change = np.full((H, W), 0.0, np.float32)
        own_nd = np.zeros((H, W), bool)
        for t in trials:
            g = per_pit[t.pit]["ground"]
            change[t.slot] = np.round(carved[t.slot] - g, 4)
            own_nd[t.slot] = g == nd
        clump_mask = ((change < 0) | own_nd).astype(np.uint8)
        clumps = core.label(torch.from_numpy(clump_mask).cuda(), 8)[0].cpu().numpy()
        for t in trials:
            info = per_pit[t.pit]
            g, inpit = info["ground"], info["inpit"]
            d, lab = change[t.slot], clumps[t.slot]
            cut = d < 0
            reach = np.unique(lab[inpit & cut])           # clumps whose carve enters the pit
            holes = np.flatnonzero((g == nd).ravel())
            if holes.size and lab.ravel()[holes[0]] in reach:
                continue                                   # that carve leaves through the raster edge
            member = np.isin(lab, reach) if reach.size else np.zeros(lab.shape, bool)
            if not member.any():
                continue
            region = member | inpit
            if abs(d[region].min()) > max_carve_depth:
                continue
            r0, c0 = t.win[0].start, t.win[1].start
            lr, lc = np.nonzero(region)
            info["found"].append(dict(vol=np.sum(np.abs(d[region])), elevs=list(carved[t.slot][region]),
                                      solution=int((info["row"].solution - 1000) + 10 * t.frac),
                                      footprint=((lr + r0) * nx + (lc + c0)).astype(int)))

    out = {}
    for pit, info in per_pit.items():
        row, cells = info["row"], info["cells"]
        cand = list(info["found"])
        # the plain fill, and the plain carve when it is complete, compete too
        cand.append(dict(vol=row.volfill, elevs=list(np.asarray(fill_vec)[cells]), solution=2003, footprint=cells))
        if row.completecarve == 1 and not np.isnan(row.volcarve + row.volfillbehind):
            allc = np.unique(np.concatenate([np.asarray(carvecells_dct[pit], dtype=np.int64),
                                             np.asarray(carve_fillcells_dct[pit], dtype=np.int64),
                                             np.asarray(cells, dtype=np.int64)]))
            cand.append(dict(vol=row.volcarve + row.volfillbehind, elevs=list(np.asarray(carve_vec)[allc]), solution=1003,
                             footprint=list(allc)))
        table = pd.DataFrame(cand, columns=["vol", "elevs", "solution", "footprint"])
        best = table.solution[table["vol"].astype(float).idxmin()]
        out[pit] = (best, table.loc[table.solution == best].copy())
    return out


def combined_solution(inDEM, inDEM_array, outpath, nx, ny, DFpits, fillcells_dct, carvecells_dct, carve_fillcells_dct,
                      pit_id, combined_fill_interval, fillID_vec, fill_vec, carve_vec, radius, maxcost, min_dist,
                      max_carve_depth, wbt, nodataval=-9999, fixflats=False, flat_increment=0.0, carve_method='SC',
                      water_array=None, wall_vec=None, del_temp_files=True):
    """tools/depressions.py:813-1137 for one pit: same arguments, same return value (solution_name,
    sol_df[vol, elevs, solution, footprint]).  A one-pit call of `combined_solutions` (above); the window rasters
    live in memory only -- upstream writes them under <outpath>/combined_solutions and deletes every one of them
    before returning (:1132-1135).  A graded `flat_increment` is served for carve_method='LC' only (there it is the
    channel's gradient)."""
    if carve_method != 'LC' and flat_increment and flat_increment > 0:
        raise NotImplementedError("flat_increment > 0 in the partial-fill search is served for carve_method='LC' only")
    os.makedirs(os.path.join(outpath, "combined_solutions"), exist_ok=True)
    res = combined_solutions([pit_id], inDEM_array, nx, ny, DFpits, fillcells_dct, carvecells_dct, carve_fillcells_dct,
                             combined_fill_interval, fillID_vec, fill_vec, carve_vec, radius, max_carve_depth,
                             nodataval=nodataval, carve_method=carve_method, maxcost=maxcost, min_dist=min_dist,
                             flat_increment=flat_increment)
    return res[pit_id]
