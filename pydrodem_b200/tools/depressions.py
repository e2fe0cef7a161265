"""Mirror of the hot-path functions of the reference's `tools/depressions.py`.

Same signatures and file protocol (`<in>_WL.tif`, `<in>_WL_DIFF.tif`, clump-ID rasters); GDAL is
replaced by pydrodem_b200.raster_io, WhiteboxTools by the `wbt` duck type (pydrodem_b200.wbt) and
scipy.ndimage.label / convolve2d by the CUDA kernels.  build_pit_catalog / select_solution /
apply_solution / fill_shallow_pits run as segmented reductions on the GPU (csrc/segment.cu);
breach_pits and combined_solution (both need WhiteboxTools' breach) are not built.
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


def fix_shifted_data(in_array, shift_array, nodataval=-9999, failtest=False, vprint=False):
    """tools/depressions.py:1406-1468 repairs a one-pixel shift artefact of WhiteboxTools' GeoTIFF
    writer; the GPU path never shifts, so the arrays are returned unchanged."""
    return shift_array


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
