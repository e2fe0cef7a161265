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


def fill_pits(inDEM, wbt, fill_method='WL', nodataval=-9999, fixflats=False, flat_increment=0.0, save_tif=True,
              downcast=True):
    """tools/depressions.py:1236-1316.  Returns (fill_array, diff_fill_array), writes `<in>_<method>.tif`
    and (save_tif) `<in>_<method>_DIFF.tif`.  `downcast` is moot: the fill is written as float32 already."""
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
    fill_array, _ = raster_io.read_raster(fill_tif)
    diff_fill_array = np.round((fill_array - inDEM_array), 3)  # :1308
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
    diff_carve_array = np.round((carve_array - inDEM_array), 3)  # :1395
    if save_tif:
        diff_tif_carve = carvefill_path.replace('.tif', '_DIFF.tif')
        raster_io.write_raster(diff_tif_carve, diff_carve_array.astype(np.float32), meta, nodata=nodataval)
    return carve_array, diff_carve_array


def create_clump_array(input_array, clump_ID_tif, stencil_tif, wbt, nodataval=-9999, noflow_array=None,
                       condition='greaterthan', clumpbuffer=None, split_mask=None, save_tif=True):
    """tools/depressions.py:1593-1675: sign mask of a difference map -> optional (2b+1)^2 box dilation ->
    labels (8-connectivity, or 4-connectivity when buffered), ids in scipy's raster-scan order."""
    input_array = np.asarray(input_array)
    clump_array = np.zeros(input_array.shape, dtype=np.int8)
    if condition == 'greaterthan':
        clump_array[input_array > 0] = 1
    if condition == 'lessthan':
        if noflow_array is not None:
            clump_array[noflow_array == 1] = 1
        clump_array[input_array < 0] = 1
    if condition == 'nonzero':
        clump_array[input_array != 0] = 1
    if clumpbuffer is not None:
        sz = (2 * clumpbuffer) + 1
        clump_array = stencils.apply_convolve_filter(clump_array, np.ones((sz, sz)))
        clump_array = (clump_array > 0).astype(np.int8)
        conn = 4
    else:
        conn = 8
    if split_mask is not None:
        split_ind = np.nonzero(split_mask)
        split_clump_array = np.zeros_like(clump_array)
        split_clump_array[split_ind] = clump_array[split_ind]
        clump_array[split_ind] = 0
        clump_ID_array, num_feat = core.label(clump_array, conn)
        split_ID_array, _ = core.label(split_clump_array, conn)
        split_ind = np.nonzero(split_ID_array)
        split_ID_array[split_ind] += num_feat
        clump_ID_array[split_ind] = split_ID_array[split_ind]
    else:
        clump_ID_array, num_feat = core.label(clump_array, conn)
    if save_tif:
        meta = raster_io.read_raster(stencil_tif)[1] if (stencil_tif and os.path.exists(stencil_tif)) else None
        raster_io.write_raster(clump_ID_tif, clump_ID_array, meta, nodata=0)
    return clump_ID_array


def fix_shifted_data(in_array, shift_array, nodataval, failtest='negative_diff', vprint=False):
    """tools/depressions.py:1406-1468: undo a whole-raster pixel shift of a WhiteboxTools output (detected from the
    bounding boxes of the valid cells) and recompute the rounded difference map.  Returns (shift_array, diff_array).
    The GPU tools never shift a raster, so on this path the shift is (0, 0) and only the difference map is rebuilt;
    the index arithmetic is kept so that a raster produced elsewhere is repaired the same way (host numpy: a rare
    repair path over index lists, not a kernel)."""
    data_ind = np.nonzero(in_array != nodataval)
    nodata_ind = np.nonzero(in_array == nodataval)
    fill_data_ind = np.nonzero(shift_array != nodataval)
    shift_1 = fill_data_ind[1].max() - data_ind[1].max()
    shift_0 = fill_data_ind[0].max() - data_ind[0].max()
    ind_diff = np.abs(len(data_ind[0]) - len(fill_data_ind[0]))
    fill_data_ind_shift = (fill_data_ind[0] - shift_0, fill_data_ind[1] - shift_1)
    shift_array[fill_data_ind_shift] = shift_array[fill_data_ind]
    shift_array[nodata_ind] = nodataval
    diff_array = np.round((shift_array - in_array), 3)
    if failtest == 'negative_diff':
        fail_diff = np.nonzero(diff_array < 0.0)
    elif failtest == 'positive_diff':
        fail_diff = np.nonzero(diff_array > 0.0)
    num_fail = len(fail_diff[0])
    assert num_fail <= ind_diff + 3, 'Re-mapped fill diff has too many negative values'
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


# ==================================================================================================
# initial_fill_carve (tools/depressions.py:41-264): the orchestration that turns one DEM GeoTIFF into the eight
# vectors the pit catalog consumes -- two fills, four clump rasters, one breach.
# ==================================================================================================
def initial_fill_carve(inDEM_tif, basename, outpath, fill_method, maxcost, radius, min_dist, nodataval, starttime, wbt,
                       sfill=None, flat_increment=0.0, watermask=None, water_mask_array=None, carveout=False,
                       carve_method='SC', verbose_print=True, del_temp_files=False):
    """Same signature, file names and return tuple as the reference:
    (modDEM_vec, fill_vec, carve_vec, diff_vec_fill, diff_vec_carve, fillID_vec, carveID_vec, carvefillID_vec, nx, ny).
    `watermask` (the lake raise of :163-176) is refused: `raise_lake_connected_pits` is broken upstream."""
    import glob
    if watermask is not None:
        raise NotImplementedError("initial_fill_carve(watermask=...) calls raise_lake_connected_pits, which cannot run "
                                  "in the reference either (tools/depressions.py:1536-1590; DESIGN.md 7)")
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

    if sfill is None:                                                                         # :160-162
        modDEM_tif = inDEM_tif
    else:                                                                                     # :163-185
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


def combined_solution(inDEM, inDEM_array, outpath, nx, ny, DFpits, fillcells_dct, carvecells_dct, carve_fillcells_dct,
                      pit_id, combined_fill_interval, fillID_vec, fill_vec, carve_vec, radius, maxcost, min_dist,
                      max_carve_depth, wbt, nodataval=-9999, fixflats=False, flat_increment=0.0, carve_method='SC',
                      water_array=None, wall_vec=None, del_temp_files=True):
    """Same arguments and return value -- (solution_name, sol_df[vol, elevs, solution, footprint]) -- as the
    reference.  Host orchestration over a window of (pit bounding box + radius + 2) cells; the breaches go through
    `wbt.breach_depressions(max_length=radius, max_depth=max_carve_depth, fill_pits=False)` and the carve clumps
    through the GPU labelling (8-connectivity, scipy numbering)."""
    import glob
    import pandas as pd
    pit_df = DFpits.loc[pit_id].copy()
    init_sol = pit_df.solution
    footprint = np.asarray(fillcells_dct[pit_id])
    completecarve = pit_df.completecarve
    combined_sol = init_sol - 1000   # 3000 or 3020
    combined_path = os.path.join(outpath, "combined_solutions")
    os.makedirs(combined_path, exist_ok=True)

    # window = bounding box of the pit + (radius + 2) cells, clipped to the raster (pyct.clip_to_window)
    clip_buf = radius + 2
    rows, cols = np.unravel_index(footprint, (ny, nx))
    row_min = int(max(0, rows.min() - clip_buf))
    row_max = int(min(ny, rows.max() + clip_buf + 1))
    col_min = int(max(0, cols.min() - clip_buf))
    col_max = int(min(nx, cols.max() + clip_buf + 1))
    ny_c, nx_c = row_max - row_min, col_max - col_min
    rr, cc = np.mgrid[row_min:row_max, col_min:col_max]
    clipped_ind = np.ravel_multi_index((rr.ravel(), cc.ravel()), (ny, nx))
    win = (slice(row_min, row_max), slice(col_min, col_max))
    _, meta = raster_io.read_raster(inDEM)
    wmeta = _window_meta(meta, row_min, col_min)
    clippedDEM = os.path.join(combined_path, "{0}.tif".format(pit_id))
    inDEM_array_clipped = np.ascontiguousarray(np.asarray(inDEM_array)[win], dtype=np.float32)
    raster_io.write_raster(clippedDEM, inDEM_array_clipped, wmeta, nodata=nodataval)
    inDEM_vec_clipped = inDEM_array_clipped.reshape(ny_c * nx_c,)
    fill_vec_clipped = np.asarray(fill_vec).reshape(ny, nx)[win].reshape(ny_c * nx_c,)
    carve_vec_orig_clipped = np.asarray(carve_vec).reshape(ny, nx)[win].reshape(ny_c * nx_c,)
    fillID_clipped_vec = np.asarray(fillID_vec)[clipped_ind]
    footprint_clipped = np.flatnonzero(fillID_clipped_vec == pit_id)

    maxfill_value = pit_df.maxfill
    vols, elevs, sol_names, dep_indices = [], [], [], []

    fraction_list = np.round(np.arange(combined_fill_interval, 1, combined_fill_interval), 1)
    for frac in fraction_list:
        # lower the full fill by a constant (keeps any gradient of the fill), never below ground (:905-914)
        SF_vec = np.copy(fill_vec_clipped)
        fill_depth_remove = maxfill_value * (1 - frac)
        SF_vec[footprint_clipped] = SF_vec[footprint_clipped] - fill_depth_remove
        SF_diff = np.round((SF_vec - inDEM_vec_clipped), 3)
        belowground_ind = np.where(SF_diff < 0)[0]
        fill_frac_ind = np.where(SF_diff > 0)[0]
        SF_vec[belowground_ind] = inDEM_vec_clipped[belowground_ind]
        p_fill_min = (SF_vec[fill_frac_ind]).min()
        carve_fill_min = (carve_vec_orig_clipped[fill_frac_ind]).min()
        if (completecarve == 1) and (p_fill_min > carve_fill_min):   # partial fill above the carve's back-fill
            continue
        clippedDEM_SF = os.path.join(combined_path, "{0}_SF_{1}.tif".format(pit_id, frac))
        raster_io.write_raster(clippedDEM_SF, SF_vec.reshape(ny_c, nx_c).astype(np.float32), wmeta, nodata=nodataval)
        clippedDEM_carved = os.path.join(combined_path, "{0}_SF_{1}_{2}.tif".format(pit_id, frac, carve_method))
        if carve_method == 'LC':
            wbt.breach_depressions_least_cost(clippedDEM_SF, clippedDEM_carved, radius, max_cost=maxcost,
                                              min_dist=min_dist, flat_increment=flat_increment, fill=False)
            clippedDEM_filled = os.path.join(combined_path, "{0}_SF_{1}_{2}_filled.tif".format(pit_id, frac, carve_method))
            wbt.fill_depressions_wang_and_liu(clippedDEM_carved, clippedDEM_filled, fix_flats=fixflats,
                                              flat_increment=flat_increment)
        else:
            wbt.breach_depressions(clippedDEM_SF, clippedDEM_carved, max_length=radius, max_depth=max_carve_depth,
                                   fill_pits=False, flat_increment=flat_increment, callback=None)
        clippedDEM_carved_array, _ = raster_io.read_raster(clippedDEM_carved)

        # carves that touch the pit, as 8-connected clumps of (lowered or nodata) cells (:958-985)
        diff_array_clip = np.round((clippedDEM_carved_array - inDEM_array_clipped), 4)
        diff_vec_clip = diff_array_clip.reshape(ny_c * nx_c,)
        mask_carve = np.zeros(inDEM_array_clipped.shape, dtype=np.int8)
        mask_carve[diff_array_clip < 0] = 1
        mask_carve[inDEM_array_clipped == nodataval] = 1
        carveID_clip_array, _ = core.label(mask_carve, 8)
        carveID_clip_vec = np.asarray(carveID_clip_array).reshape(ny_c * nx_c,)
        carve_ind_clipped = np.flatnonzero(diff_vec_clip < 0.0)
        carve_intersect = np.intersect1d(footprint_clipped, carve_ind_clipped, assume_unique=True)
        carveIDs_pit = np.unique(carveID_clip_vec[carve_intersect])
        carvecells_add = list(np.flatnonzero(np.isin(carveID_clip_vec, carveIDs_pit)))
        nodata_clipped_ind = np.flatnonzero(inDEM_vec_clipped == nodataval)
        if len(nodata_clipped_ind > 0):        # the window holds raster edge: a carve running off it is disqualified
            carveout_ID = carveID_clip_vec[nodata_clipped_ind][0]
            if carveout_ID in list(carveIDs_pit):
                continue
        if len(carvecells_add) < 1:
            continue
        expanded_footprint_clipped = list(set(carvecells_add + list(footprint_clipped)))
        expanded_footprint = clipped_ind[expanded_footprint_clipped]
        volchange = np.sum(abs(diff_vec_clip[expanded_footprint_clipped]))
        maxcarvedepth = abs(diff_vec_clip[expanded_footprint_clipped].min())
        if maxcarvedepth <= max_carve_depth:
            elevs.append(list(clippedDEM_carved_array.reshape(ny_c * nx_c,)[expanded_footprint_clipped]))
            vols.append(volchange)
            dep_indices.append(expanded_footprint.astype(int))
            if int(frac * 10) == 0:
                if combined_sol == 3020:
                    sol_names.append(1020)
                elif combined_sol == 3000:
                    sol_names.append(1003)
            else:
                sol_names.append(int(combined_sol + (10 * frac)))
        if del_temp_files:
            for f in glob.glob(os.path.join(combined_path, '{0}_*.tif'.format(pit_id))):
                if os.path.isfile(f):
                    os.remove(f)

    # the plain fill, and the plain carve when it is complete, compete too (:1098-1117)
    sol_names.append(2003)
    vols.append(pit_df.volfill)
    dep_indices.append(footprint)
    elevs.append(list(np.asarray(fill_vec)[footprint]))
    if completecarve == 1:
        vol_total = pit_df.volcarve + pit_df.volfillbehind
        carve_cells_all = list(set(list(carvecells_dct[pit_id]) + list(carve_fillcells_dct[pit_id]) + list(footprint)))
        if not np.isnan(vol_total):
            sol_names.append(1003)
            vols.append(vol_total)
            dep_indices.append(carve_cells_all)
            elevs.append(list(np.asarray(carve_vec)[carve_cells_all]))

    df_combined = pd.DataFrame(columns=["vol", "elevs", "solution", "footprint"])
    df_combined["vol"] = vols
    df_combined["elevs"] = elevs
    df_combined["solution"] = sol_names
    df_combined["footprint"] = dep_indices
    solution_name = df_combined.solution[df_combined["vol"].idxmin()]
    sol_df = df_combined.loc[df_combined.solution == solution_name].copy()
    for f in glob.glob(os.path.join(combined_path, '*.tif')):
        if os.path.isfile(f):
            os.remove(f)
    return solution_name, sol_df
