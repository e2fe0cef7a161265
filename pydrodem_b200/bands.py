"""Row-band sharding of the fill -> D8 -> flats -> accumulation chain (BASELINE.json
configs[2]: one large DEM over the GPUs of a box).

The reference gets its multi-process runs from TauDEM's own MPI decomposition (`mpiexec -n N
PitRemove / D8FlowDir / AreaD8`, run_aws.py:584-612): the DEM is cut into row bands, one per rank.
Here a band lives on one GPU; every pass is "band-local phase -> seam exchange -> finish phase":

  fill    local watershed labels with the seam cells as terminals, the terminal graph reduced to its
          minimum spanning forest (Barnes 2016 with a band as the tile) -> ONE all_gather of <= 5 nx
          edges per band -> every rank solves the seam graph redundantly and applies the levels;
  D8      needs the filled halo rows (one all_gather of seam rows);
  flats   the two geodesic distance fields are relaxed band-locally, seam rows exchanged, repeated
          until no halo row changes anywhere (all_gather + all_reduce per round);
  accum   local counts, flow leaving the band parked per seam cell, where inflow entering a seam cell
          leaves again (Barnes 2017) -> ONE all_gather -> seam forest solved redundantly -> final pass.

All exchanges are all_gathers of seam-sized records (a few nx words), issued through
torch.distributed (NCCL over NVLink on GPUs); the kernels are the library's hc_band_* entry points.
A process may hold several consecutive bands (world_size 1 with N bands emulates N ranks on one
GPU: that is what the single-GPU parity tests do).  Results are bit-identical to the whole-raster
chain.
"""
import ctypes

from . import _lib
from ._lib import SEED_TAUDEM, SEED_WBT, TABLE_TAUDEM  # noqa: F401


def split_rows(ny, nbands):
    """Row counts of `nbands` contiguous bands covering ny rows (first bands one row taller)."""
    base, extra = divmod(ny, nbands)
    return [base + (1 if i < extra else 0) for i in range(nbands)]


class SeamComm:
    """all_gather / any-reduce of per-band seam records over the ranks (identity for world 1)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1

    def gather(self, local):
        """local: [nlocal, ...] -> [world * nlocal, ...] in band order (ranks hold consecutive bands)."""
        if self.world == 1:
            return local
        import torch
        out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        self.dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out

    def any(self, flag, device):
        if self.world == 1:
            return bool(flag)
        import torch
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return bool(t.item())


def exchange_halos(bufs, first, nbands, comm):
    """bufs: this process's halo'd band rasters [(nb+2), nx] (consecutive global bands first, first+1, ...).
    Row 0 / row -1 of each are overwritten with the neighbouring band's last / first own row: ONE all_gather of
    the [nlocal, 2, nx] seam rows.  Works on CPU tensors over gloo as well (tests)."""
    import torch
    seams = torch.stack([torch.stack((b[1], b[-2])) for b in bufs])  # [nl, 2, nx]
    allseams = comm.gather(seams)
    for i, b in enumerate(bufs):
        g = first + i
        if g > 0:
            b[0].copy_(allseams[g - 1, 1])
        if g < nbands - 1:
            b[-1].copy_(allseams[g + 1, 0])


class BandedChain:
    """The chain on this process's bands.  `band_rows`: rows of each LOCAL band (consecutive global
    bands first_band, first_band+1, ...); `nbands`: bands in the whole DEM."""

    def __init__(self, band_rows, nx, nbands, first_band, device, comm=None, nodata=-9999.0, dx=1.0, dy=1.0,
                 seed_mode=SEED_TAUDEM, table=TABLE_TAUDEM, flats=True, want_slope=False):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.dev_index = self.device.index or 0
        self.rows = list(band_rows)
        self.nx, self.nbands, self.first = int(nx), int(nbands), int(first_band)
        self.nodata, self.dx, self.dy = float(nodata), float(dx), float(dy)
        self.seed_mode, self.table, self.flats, self.want_slope = seed_mode, table, bool(flats), bool(want_slope)
        self.comm = comm if comm is not None else SeamComm()
        nl = len(self.rows)
        if self.comm.world * nl != self.nbands:
            raise ValueError("every rank must hold nbands / world consecutive bands")
        self.cap = int(self.lib.hc_band_fill_edge_cap(self.nx))
        self.ctx = [_lib.new_context(self.dev_index) for _ in range(nl)]
        f32, u8, i32 = torch.float32, torch.uint8, torch.int32
        mk = lambda nb, dt: torch.empty((nb + 2, self.nx), dtype=dt, device=self.device)  # noqa: E731
        self.zh = [mk(nb, f32) for nb in self.rows]
        self.fh = [mk(nb, f32) for nb in self.rows]
        self.da = [mk(nb, u8) for nb in self.rows]
        self.db = [mk(nb, u8) for nb in self.rows] if self.flats else self.da
        self.acc = [torch.empty((nb, self.nx), dtype=i32, device=self.device) for nb in self.rows]
        self.slope = [torch.empty((nb, self.nx), dtype=f32, device=self.device) for nb in self.rows] \
            if self.want_slope else [None] * nl
        self.edges = torch.zeros((nl, self.cap, 3), dtype=i32, device=self.device)
        self.fseam = torch.zeros((nl, 2, 3, self.nx), dtype=i32, device=self.device)
        self.aseam = torch.zeros((nl, 2, 2, self.nx), dtype=i32, device=self.device)
        self.fchanged = torch.zeros(1, dtype=i32, device=self.device)
        self.flat_rounds = 0
        self.timing = False      # set True to collect per-phase device times of run() in self.phase_ms
        self.phase_ms = {}
        self._ev = []

    def _mark(self, name):
        if self.timing:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record(self.torch.cuda.current_stream(self.device))
            self._ev.append((name, e))

    def _collect(self):
        if not self.timing or len(self._ev) < 2:
            return
        self.torch.cuda.synchronize(self.device)
        for (n0, e0), (n1, e1) in zip(self._ev[:-1], self._ev[1:]):
            self.phase_ms[n1] = self.phase_ms.get(n1, 0.0) + e0.elapsed_time(e1)
        self._ev = []

    def close(self):
        for c in self.ctx:
            _lib.destroy_context(c)
        self.ctx = []

    # ---- helpers ---------------------------------------------------------------------------------
    def flags(self, i):
        g = self.first + i
        return (1 if g > 0 else 0) | (2 if g < self.nbands - 1 else 0)

    def z_view(self, i):
        """Own rows of local band i's input buffer (write the band's DEM here, then run())."""
        return self.zh[i][1:-1]

    def _p(self, t, row=1):
        """device pointer to own row 0 of a halo'd buffer (row=1) or to a plain tensor (row=0)."""
        return ctypes.c_void_p(t.data_ptr() + row * self.nx * t.element_size())

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def _halo_exchange(self, bufs):
        """Fill the halo rows of every local band from its neighbours' seam rows."""
        exchange_halos(bufs, self.first, self.nbands, self.comm)

    def _chk(self, rc, what):
        _lib.check(rc, what)

    def _exterior_nodata(self):
        """WBT seed rule on bands: nodata cells connected to the raster frame through nodata (8-connectivity).  Banded
        labels of the nodata mask, a per-component "touches the frame" flag (max-reduced over the ranks), and the flag
        gathered back into halo'd uint8 rasters."""
        torch, lib, nx = self.torch, self.lib, self.nx
        nl = len(self.rows)
        st = self._stream()
        masks = []
        for i in range(nl):
            z = self.z_view(i)
            masks.append(((z == self.nodata) | torch.isnan(z)).to(torch.uint8))
        any_nd = self.comm.any(any(bool(m.any()) for m in masks), self.device)
        if not any_nd:
            return [None] * nl
        rows_all = self.comm.gather(torch.tensor(self.rows, dtype=torch.int64, device=self.device).reshape(nl, 1)).reshape(-1)
        row0_all = (torch.cumsum(rows_all, 0) - rows_all).tolist()
        bl = BandedLabel(self.rows, nx, self.nbands, self.first, [row0_all[self.first + i] for i in range(nl)], self.device,
                         comm=self.comm)
        try:
            labs, total = bl.run(masks, 8)
        finally:
            bl.close()
        touch = torch.zeros(total + 1, dtype=torch.float32, device=self.device)
        for i, nb in enumerate(self.rows):
            g = self.first + i
            border = torch.zeros((nb, nx), dtype=torch.float32, device=self.device)
            border[:, 0] = 1
            border[:, -1] = 1
            if g == 0:
                border[0] = 1
            if g == self.nbands - 1:
                border[-1] = 1
            t = torch.empty(total + 1, dtype=torch.float32, device=self.device)
            self._chk(lib.hc_seg_stats_dev(self.ctx[i], ctypes.c_void_p(labs[i].data_ptr()), ctypes.c_void_p(border.data_ptr()),
                                           None, nb * nx, total, 1, ctypes.c_void_p(t.data_ptr()), None, None, None, st),
                      "hc_seg_stats_dev")
            touch = torch.maximum(touch, torch.clamp(t, min=0.0))
        if self.comm.world > 1:
            self.comm.dist.all_reduce(touch, op=self.comm.dist.ReduceOp.MAX, group=self.comm.group)
        flag = (touch > 0).to(torch.uint8)
        flag[0] = 0
        exts = []
        for i, nb in enumerate(self.rows):
            e = torch.zeros((nb + 2, nx), dtype=torch.uint8, device=self.device)
            self._chk(lib.hc_gather_flag_dev(self.ctx[i], ctypes.c_void_p(labs[i].data_ptr()), ctypes.c_void_p(flag.data_ptr()),
                                             self._p(e), nb * nx, st), "hc_gather_flag_dev")
            exts.append(e)
        self._halo_exchange(exts)
        return exts

    # ---- the chain -------------------------------------------------------------------------------
    def run(self, z_list=None):
        torch, lib, nx = self.torch, self.lib, self.nx
        nl = len(self.rows)
        with torch.cuda.device(self.device):
            st = self._stream()
            if z_list is not None:
                for i, z in enumerate(z_list):
                    self.z_view(i).copy_(z)
            self._mark("start")
            self._halo_exchange(self.zh)
            self._mark("halo_z")
            exts = self._exterior_nodata() if self.seed_mode != SEED_TAUDEM else [None] * nl
            # fill
            for i, nb in enumerate(self.rows):
                self._chk(lib.hc_band_fill_local_dev(self.ctx[i], self._p(self.zh[i]), nb, nx, self.nodata, self.seed_mode,
                                                     self.flags(i), self.first + i,
                                                     self._p(exts[i]) if exts[i] is not None else None,
                                                     ctypes.c_void_p(self.edges[i].data_ptr()), st), "hc_band_fill_local_dev")
            self._mark("fill_local")
            all_edges = self.comm.gather(self.edges)
            self._mark("gather_edges")
            for i, nb in enumerate(self.rows):
                self._chk(lib.hc_band_fill_finish_dev(self.ctx[i], ctypes.c_void_p(all_edges.data_ptr()), self.nbands,
                                                      self._p(self.fh[i]), st), "hc_band_fill_finish_dev")
            self._mark("fill_finish")
            self._halo_exchange(self.fh)
            self._mark("halo_filled")
            # D8 (+ slope)
            for i, nb in enumerate(self.rows):
                sp = ctypes.c_void_p(self.slope[i].data_ptr()) if self.want_slope else None
                self._chk(lib.hc_band_d8_dev(self.ctx[i], self._p(self.fh[i]), self._p(self.da[i]), sp, nb, nx, self.nodata,
                                             self.dx, self.dy, self.table, self.flags(i), st), "hc_band_d8_dev")
            self._mark("d8")
            self._halo_exchange(self.da)
            self._mark("halo_dir")
            # flats
            if self.flats:
                for i, nb in enumerate(self.rows):
                    self._chk(lib.hc_band_flats_begin_dev(self.ctx[i], self._p(self.fh[i]), self._p(self.da[i]),
                                                          self._p(self.db[i]), nb, nx, self.nodata, self.table,
                                                          self.flags(i), st), "hc_band_flats_begin_dev")
                self._mark("flats_begin")
                self.flat_rounds = 0
                first = True
                while True:
                    if not first:
                        for i in range(nl):
                            self._chk(lib.hc_band_flats_relax_dev(self.ctx[i], st), "hc_band_flats_relax_dev")
                        self.flat_rounds += 1
                    for i in range(nl):
                        self._chk(lib.hc_band_flats_get_seams_dev(self.ctx[i], ctypes.c_void_p(self.fseam[i].data_ptr()), st),
                                  "hc_band_flats_get_seams_dev")
                    allf = self.comm.gather(self.fseam)
                    # halo rows in, seam tile rows re-queued, "something changed" accumulated -- all on the device; the
                    # ranks then agree on convergence with ONE all_reduce + read-back per round
                    self.fchanged.zero_()
                    for i in range(nl):
                        g = self.first + i
                        top = ctypes.c_void_p(allf[g - 1, 1].data_ptr()) if g > 0 else None
                        bot = ctypes.c_void_p(allf[g + 1, 0].data_ptr()) if g < self.nbands - 1 else None
                        self._chk(lib.hc_band_flats_put_halos_flag_dev(self.ctx[i], top, bot,
                                                                       ctypes.c_void_p(self.fchanged.data_ptr()), st),
                                  "hc_band_flats_put_halos_flag_dev")
                    if self.comm.world > 1:
                        self.comm.dist.all_reduce(self.fchanged, op=self.comm.dist.ReduceOp.MAX, group=self.comm.group)
                    changed = bool(self.fchanged.item())
                    if not first and not changed:
                        break
                    first = False
                    if self.flat_rounds > 100000:
                        raise _lib.HydrocoreError("banded flats did not converge")
                self._mark("flats_relax_exchange")
                for i in range(nl):
                    self._chk(lib.hc_band_flats_finish_dev(self.ctx[i], st), "hc_band_flats_finish_dev")
                self._halo_exchange(self.db)
                self._mark("flats_finish_halo")
            # accumulation
            for i, nb in enumerate(self.rows):
                self._chk(lib.hc_band_accum_local_dev(self.ctx[i], self._p(self.db[i]), ctypes.c_void_p(self.acc[i].data_ptr()), nb, nx,
                                                      self.table, self.flags(i),
                                                      ctypes.c_void_p(self.aseam[i].data_ptr()), st),
                          "hc_band_accum_local_dev")
            self._mark("accum_local")
            alls = self.comm.gather(self.aseam)
            self._mark("gather_seams")
            for i, nb in enumerate(self.rows):
                self._chk(lib.hc_band_accum_finish_dev(self.ctx[i], ctypes.c_void_p(alls.data_ptr()), self.nbands,
                                                       self.first + i, ctypes.c_void_p(self.acc[i].data_ptr()), st),
                          "hc_band_accum_finish_dev")
            self._mark("accum_finish")
            self._collect()
        return [(self.fh[i][1:-1], self.db[i][1:-1], self.slope[i], self.acc[i]) for i in range(nl)]

    def launches(self):
        return sum(int(self.lib.hc_launch_count(c)) for c in self.ctx)

    def stats(self):
        out = []
        for c in self.ctx:
            te, mr, gr = ctypes.c_int64(0), ctypes.c_int32(0), ctypes.c_int32(0)
            self.lib.hc_band_stats(c, ctypes.byref(te), ctypes.byref(mr), ctypes.byref(gr))
            out.append({"term_edges": int(te.value), "msf_rounds": int(mr.value), "seam_rounds": int(gr.value)})
        return out


def fill_d8_accum_banded(z, nbands, nodata=-9999.0, dx=1.0, dy=1.0, table=TABLE_TAUDEM, flats=True, want_slope=True,
                         seed_mode=SEED_TAUDEM):
    """Whole-raster convenience: cut `z` (2-D torch CUDA tensor or numpy array) into `nbands` row bands on one
    device, run the banded chain, and stitch the results.  Same return as core.fill_d8_accum."""
    import numpy as np
    import torch
    host = not type(z).__module__.startswith("torch")
    zt = torch.from_numpy(np.ascontiguousarray(z, dtype=np.float32)).cuda() if host else z
    ny, nx = zt.shape
    rows = split_rows(ny, nbands)
    ch = BandedChain(rows, nx, nbands, 0, zt.device, nodata=nodata, dx=dx, dy=dy, table=table, flats=flats,
                     want_slope=want_slope, seed_mode=seed_mode)
    try:
        r0 = 0
        zs = []
        for nb in rows:
            zs.append(zt[r0:r0 + nb])
            r0 += nb
        res = ch.run(zs)
        f = torch.cat([r[0] for r in res])
        d = torch.cat([r[1] for r in res])
        s = torch.cat([r[2] for r in res]) if want_slope else None
        a = torch.cat([r[3] for r in res])
        torch.cuda.synchronize()
        ch.last_stats = ch.stats()
        fill_d8_accum_banded.last_stats = ch.last_stats
        fill_d8_accum_banded.last_flat_rounds = ch.flat_rounds
    finally:
        ch.close()
    if host:
        return (f.cpu().numpy(), d.cpu().numpy(), s.cpu().numpy() if s is not None else None,
                a.cpu().numpy().view(np.uint32))
    return f, d, s, a


class BandedLabel:
    """scipy.ndimage.label numbering of ONE mask cut into row bands (SURVEY.md §8e "Labelling").  Local union-find
    per band, min-root propagation across the seams (all_gather + relax rounds), ids = owner band's offset + local rank;
    see csrc/label.cu.  `band_rows`: rows of this process's consecutive bands; returns int32 label tensors + total."""

    def __init__(self, band_rows, nx, nbands, first_band, row0s, device, comm=None):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.rows, self.nx, self.nbands, self.first = list(band_rows), int(nx), int(nbands), int(first_band)
        self.row0s = list(row0s)          # global first row of each LOCAL band
        self.comm = comm if comm is not None else SeamComm()
        self.ctx = [_lib.new_context(self.device.index or 0) for _ in self.rows]
        self.rounds = 0

    def close(self):
        for c in self.ctx:
            _lib.destroy_context(c)
        self.ctx = []

    def run(self, masks, conn=8):
        torch, lib, nx = self.torch, self.lib, self.nx
        nl = len(self.rows)
        p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
        with torch.cuda.device(self.device):
            st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            masks = [m.to(torch.uint8).contiguous() for m in masks]
            seams = torch.empty((nl, 2, nx), dtype=torch.int32, device=self.device)
            for i, nb in enumerate(self.rows):
                _lib.check(lib.hc_band_label_local_dev(self.ctx[i], p(masks[i]), nb, nx, conn, self.row0s[i], p(seams[i]), st),
                           "hc_band_label_local_dev")
            self.rounds = 0
            while True:
                alls = self.comm.gather(seams)
                changed = False
                for i in range(nl):
                    ch = ctypes.c_int(0)
                    _lib.check(lib.hc_band_label_relax_dev(self.ctx[i], p(alls), self.nbands, self.first + i, p(seams[i]),
                                                           ctypes.byref(ch), st), "hc_band_label_relax_dev")
                    changed = changed or bool(ch.value)
                self.rounds += 1
                if not self.comm.any(changed, self.device):
                    break
                if self.rounds > 100000:
                    raise _lib.HydrocoreError("banded labelling did not converge")
            owned = torch.zeros((nl,), dtype=torch.int64, device=self.device)
            for i in range(nl):
                n = ctypes.c_int64(0)
                _lib.check(lib.hc_band_label_rank_dev(self.ctx[i], ctypes.byref(n), st), "hc_band_label_rank_dev")
                owned[i] = n.value
            counts = self.comm.gather(owned.reshape(nl, 1)).reshape(-1)
            offs = torch.cumsum(counts, 0) - counts
            total = int(counts.sum())
            pairs = torch.empty((nl, 2, nx, 2), dtype=torch.int32, device=self.device)
            for i in range(nl):
                _lib.check(lib.hc_band_label_pairs_dev(self.ctx[i], int(offs[self.first + i]), p(pairs[i]), st),
                           "hc_band_label_pairs_dev")
            allp = self.comm.gather(pairs).reshape(-1, 2)
            allp = allp[allp[:, 1] > 0]
            keys64 = allp[:, 0].to(torch.int64) & 0xFFFFFFFF
            order = torch.argsort(keys64)
            keys = allp[order, 0].contiguous()
            vals = allp[order, 1].contiguous()
            labs = []
            for i, nb in enumerate(self.rows):
                lab = torch.empty((nb, nx), dtype=torch.int32, device=self.device)
                _lib.check(lib.hc_band_label_finish_dev(self.ctx[i], int(offs[self.first + i]), p(keys), p(vals), keys.numel(),
                                                        p(lab), st), "hc_band_label_finish_dev")
                labs.append(lab)
        return labs, total


def label_banded(mask, nbands, conn=8):
    """Whole-raster convenience (single process, `nbands` local bands): returns (int32 labels, n) like core.label."""
    import numpy as np
    import torch
    host = not type(mask).__module__.startswith("torch")
    m = torch.from_numpy(np.ascontiguousarray(np.asarray(mask) != 0).astype(np.uint8)).cuda() if host else mask
    ny, nx = m.shape
    rows = split_rows(ny, nbands)
    row0s = [sum(rows[:i]) for i in range(nbands)]
    bl = BandedLabel(rows, nx, nbands, 0, row0s, m.device)
    try:
        labs, n = bl.run([m[r0:r0 + nb] for r0, nb in zip(row0s, rows)], conn)
        out = torch.cat(labs)
        torch.cuda.synchronize()
        label_banded.last_rounds = bl.rounds
    finally:
        bl.close()
    return (out.cpu().numpy(), n) if host else (out, n)


# ---- local-window functions over row bands (SURVEY.md §8e "smoothing stencils: bands + halo rows") ----------------
# Any function whose output at a cell depends only on cells within `halo` rows (water stencils, vegetation bias,
# convolutions, moving min/max) shards trivially: give every band `halo` extra rows from its neighbours, run the
# function on the halo'd rasters, keep the own rows.  At the raster's top / bottom the band edge IS the raster edge, so
# the function's own boundary rule (scipy's reflect) applies unchanged -- the stitched result is bit-identical.
def apply_banded(fn, rasters, nbands, halo, out_index=None):
    """Single-process emulation (N bands on one GPU, as the parity tests of the chain do).  `rasters`: 2-D arrays /
    tensors of equal shape, `fn(*band_rasters)` -> raster or tuple of rasters.  Returns the stitched result(s)."""
    ny = rasters[0].shape[0]
    rows = split_rows(ny, nbands)
    if nbands > 1 and min(rows) < halo:
        raise ValueError("bands must be at least `halo` rows tall")
    outs, r0 = [], 0
    for i, nb in enumerate(rows):
        a, b = max(0, r0 - halo) if i > 0 else r0, min(ny, r0 + nb + halo) if i < nbands - 1 else r0 + nb
        res = fn(*[_band_copy(x[a:b]) for x in rasters])
        res = res if isinstance(res, tuple) else (res,)
        outs.append(tuple(o[r0 - a:r0 - a + nb] for o in res))
        r0 += nb
    cat = [_cat([o[k] for o in outs]) for k in range(len(outs[0]))]
    return cat[0] if len(cat) == 1 else tuple(cat)


def _band_copy(x):
    return x.clone() if hasattr(x, "clone") else x.copy()


def _cat(parts):
    if hasattr(parts[0], "is_cuda") or hasattr(parts[0], "clone"):
        import torch
        return torch.cat(parts)
    import numpy as np
    return np.concatenate(parts)


def apply_rank(fn, own_rasters, halo, comm=None):
    """One band per rank: `own_rasters` = this rank's rows of every raster (torch tensors, CPU over gloo or CUDA over
    NCCL).  ONE all_gather per raster carries every band's first / last `halo` rows; the function runs on the halo'd
    band and the own rows of its result(s) are returned."""
    import torch
    comm = comm if comm is not None else SeamComm()
    world = comm.world
    rank = comm.dist.get_rank(comm.group) if world > 1 else 0
    nb = own_rasters[0].shape[0]
    if world > 1 and nb < halo:
        raise ValueError("bands must be at least `halo` rows tall")
    haloed = []
    for x in own_rasters:
        if world == 1:
            haloed.append(x)
            continue
        edge = torch.stack((x[:halo], x[nb - halo:])).unsqueeze(0).contiguous()   # [1, 2, halo, nx]
        alle = comm.gather(edge)
        parts = ([alle[rank - 1, 1]] if rank > 0 else []) + [x] + ([alle[rank + 1, 0]] if rank < world - 1 else [])
        haloed.append(torch.cat(parts))
    res = fn(*haloed)
    top = halo if rank > 0 else 0
    crop = lambda o: o[top:top + nb]  # noqa: E731
    return tuple(crop(o) for o in res) if isinstance(res, tuple) else crop(res)


# ---- pit catalog / selection / apply over row bands (SURVEY.md §8e "Catalog / select / apply": AllReduce of per-label
# ---- partial max / sum).  The per-cell reductions run band by band (hc_band_catalog_local_dev, hc_band_select_partials_dev),
# ---- the per-label tables are reduced across the bands (torch.distributed all_reduce over NCCL; plain tensor ops between the
# ---- bands one process holds), the small table logic runs redundantly on every rank, and apply_solution is per cell.
def u64_order_flip(t):
    """int64 view of uint64 keys <-> int64 values whose SIGNED order is the keys' UNSIGNED order (flip the sign bit;
    the map is its own inverse): lets a signed MAX / MIN all_reduce stand in for the unsigned one"""
    return t ^ (-(1 << 63))


class BandedCatalog:
    """dep.build_pit_catalog / select_solution / apply_solution (tools/depressions.py:272-810, :1140-1228) for one raster
    cut into row bands.  `bands`: this process's consecutive bands, each a dict of flat device tensors
    dem, carve, dcarve, dfill (float32) and fid, cid, cfid (int32 label rasters with WHOLE-RASTER ids, e.g. from
    BandedLabel) plus `cell0`, the whole-raster index of the band's first cell.  Results equal the whole-raster calls
    (counts, maxima, codes and the applied surface exactly; the float64-accumulated volume columns to the last ulp of
    their float32 value, since partial sums are added in a different order)."""

    def __init__(self, bands, nodataval, carveout=False, comm=None):
        import torch
        self.torch = torch
        self.comm = comm if comm is not None else SeamComm()
        self.bands = bands
        self.nodata = float(nodataval)
        self.carveout = bool(carveout)
        self.lib = _lib.load()
        self.dev = bands[0]["dem"].device
        self.ctx = _lib.context(self.dev.index or 0)
        self.state = None

    # -- reductions: between this process's bands by tensor ops, between the ranks by all_reduce ----------------------
    def _reduce(self, parts, op):
        torch, dist = self.torch, self.comm.dist
        fn = {"max": torch.maximum, "min": torch.minimum, "sum": torch.add}[op]
        out = parts[0].clone()
        for p in parts[1:]:
            out = fn(out, p)
        if self.comm.world > 1:
            rop = {"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}[op]
            dist.all_reduce(out, op=rop, group=self.comm.group)
        return out

    def _reduce_u64(self, parts, op):
        """unsigned 64-bit max / min of keys held in int64 tensors (torch.distributed has no unsigned reduction)"""
        return u64_order_flip(self._reduce([u64_order_flip(p) for p in parts], op))

    def _p(self, t):
        return ctypes.c_void_p(t.data_ptr()) if t is not None else None

    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.dev).cuda_stream)

    def build(self):
        """-> DFpits (the reference's columns), replicated on every rank"""
        import numpy as np
        import pandas as pd
        torch = self.torch
        f32, f64, i32, i64 = torch.float32, torch.float64, torch.int32, torch.int64
        dv = self.dev
        big = torch.iinfo(i64).max
        # label counts and the carve that holds the first nodata cell of the WHOLE raster (:335)
        loc = []
        for b in self.bands:
            nd = torch.nonzero(b["dem"] == self.nodata).reshape(-1)
            first = big if nd.numel() == 0 else ((int(b["cell0"]) + int(nd[0])) << 32) | max(int(b["cid"][nd[0]]), 0)
            loc.append(torch.tensor([int(b["fid"].max()), max(int(b["cid"].max()), 0), -first], dtype=i64, device=dv))
        g = self._reduce(loc, "max")
        npits, ncarves, first = int(g[0]), int(g[1]), -int(g[2])
        carveout_id = 0
        if not self.carveout:
            if first == big:
                raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # as the reference (:335)
            carveout_id = first & 0xFFFFFFFF
        P1, C1 = npits + 1, ncarves + 1
        names = (("maxfill", P1, f32, "max"), ("volfill64", P1, f64, "sum"), ("nfill", P1, i64, "sum"), ("fb_max", P1, f32, "max"),
                 ("fb_sum", P1, f64, "sum"), ("fb_n", P1, i64, "sum"), ("carve_min", C1, f32, "min"), ("carve_sum", C1, f64, "sum"),
                 ("carve_n", C1, i64, "sum"), ("kmax", C1, i64, "umax"), ("kmin", C1, i64, "umin"))
        parts = {k: [] for k, *_ in names}
        pair_parts = []
        with torch.cuda.device(dv):
            for b in self.bands:
                n = b["dem"].numel()
                b["carve_bound"] = torch.empty(n, dtype=i32, device=dv)
                t = {k: torch.empty(m, dtype=dt, device=dv) for k, m, dt, _ in names}
                cap = max(1024, 4 * (npits + ncarves) + 1024)
                npairs = ctypes.c_int64(0)
                while True:
                    pairs = torch.empty(cap, dtype=i64, device=dv)
                    rc = self.lib.hc_band_catalog_local_dev(
                        self.ctx, self._p(b["dem"]), self.nodata, self._p(b["carve"]), self._p(b["dcarve"]), self._p(b["dfill"]),
                        self._p(b["fid"]), self._p(b["cid"]), self._p(b["cfid"]), n, int(b["cell0"]), npits, ncarves,
                        self._p(b["carve_bound"]), *[self._p(t[k]) for k, *_ in names], self._p(pairs), cap,
                        ctypes.byref(npairs), self._stream())
                    if rc == -4 and cap < (1 << 28):   # HC_ERR_TOO_LARGE: more distinct pairs than the table holds
                        cap *= 8
                        continue
                    _lib.check(rc, "hc_band_catalog_local_dev")
                    break
                for k, *_ in names:
                    parts[k].append(t[k])
                pair_parts.append(pairs[: int(npairs.value)])
            red = {}
            for k, _, _, op in names:
                red[k] = self._reduce_u64(parts[k], op[1:]) if op[0] == "u" else self._reduce(parts[k], op)
            # the distinct pairs of all bands
            mine = torch.cat(pair_parts) if pair_parts else torch.empty(0, dtype=i64, device=dv)
            if self.comm.world > 1:
                cnt = torch.tensor([mine.numel()], dtype=i64, device=dv)
                allc = self.comm.gather(cnt)
                width = int(allc.max())
                padded = torch.zeros((1, max(width, 1)), dtype=i64, device=dv)
                padded[0, : mine.numel()] = mine
                mine = self.comm.gather(padded).reshape(-1)
                mine = mine[mine != 0]
            pairs = torch.unique(mine)
            np_ = int(pairs.numel())
            pairs = pairs.contiguous() if np_ else torch.zeros(1, dtype=i64, device=dv)
            # pit ids at the carves' extreme cells: every band answers for the cells it owns
            fa, fb = [], []
            for b in self.bands:
                a = torch.empty(C1, dtype=i32, device=dv)
                c = torch.empty(C1, dtype=i32, device=dv)
                _lib.check(self.lib.hc_band_catalog_owner_ids_dev(self.ctx, self._p(b["fid"]), b["fid"].numel(), int(b["cell0"]),
                                                                   ncarves, self._p(red["kmax"]), self._p(red["kmin"]),
                                                                   self._p(a), self._p(c), self._stream()),
                           "hc_band_catalog_owner_ids_dev")
                fa.append(a)
                fb.append(c)
            fid_amax, fid_amin = self._reduce(fa, "max"), self._reduce(fb, "max")
            out = {k: torch.empty(P1, dtype=f32, device=dv) for k in ("volfill", "maxcarve", "volcarve", "volfb", "maxfb")}
            ncarvecells = torch.empty(P1, dtype=i64, device=dv)
            complete = torch.empty(P1, dtype=torch.int8, device=dv)
            pit_out = torch.empty(P1, dtype=torch.uint8, device=dv)
            included = torch.zeros(max(np_, 1), dtype=torch.uint8, device=dv)
            _lib.check(self.lib.hc_catalog_finish_dev(
                self.ctx, npits, ncarves, self._p(pairs), np_, self._p(fid_amax), self._p(fid_amin), int(self.carveout),
                carveout_id, self._p(red["volfill64"]), self._p(red["fb_max"]), self._p(red["fb_sum"]), self._p(red["fb_n"]),
                self._p(red["carve_min"]), self._p(red["carve_sum"]), self._p(red["carve_n"]), self._p(out["volfill"]),
                self._p(out["maxcarve"]), self._p(out["volcarve"]), self._p(ncarvecells), self._p(out["volfb"]),
                self._p(out["maxfb"]), self._p(complete), self._p(pit_out), self._p(included), self._stream()),
                "hc_catalog_finish_dev")
        self.state = dict(npits=npits, ncarves=ncarves, pairs=pairs, npairs=np_, included=included, pit_out=pit_out,
                          ncarvecells=ncarvecells, nfill=red["nfill"], maxfill=red["maxfill"], complete=complete, **out)
        cols = {k: v[1:].cpu().numpy() for k, v in out.items()}
        DFpits = pd.DataFrame({'fillcarveID': np.arange(1, npits + 1), 'maxfill': red["maxfill"][1:].cpu().numpy(),
                               'volfill': cols["volfill"], 'maxcarve': cols["maxcarve"], 'volcarve': cols["volcarve"],
                               'volfillbehind': cols["volfb"], 'maxfillbehind': cols["maxfb"],
                               'completecarve': complete[1:].cpu().numpy()})
        DFpits.index = DFpits['fillcarveID']
        DFpits.drop(labels=['fillcarveID'], axis=1, inplace=True)
        return DFpits

    def select(self, DFpits, fill_vol_thresh, carve_vol_thresh, max_carve_depth, maxcarve_factor=1.25, apply_shallow_fill=False,
               wall=None, water=None, sink=None, numpy_scalars=None):
        """dep.select_solution: `wall` / `water` / `sink` are lists with one flat float32 device tensor per local band
        (wall / sink may be None); adds the 'solution' and 'downstream_pit' columns."""
        import numpy as np
        torch = self.torch
        if water is None:
            raise NameError("name 'water_mask_vec' is not defined")  # the reference's latent bug, kept
        if numpy_scalars is None:
            numpy_scalars = "weak" if int(np.__version__.split(".")[0]) >= 2 else "legacy"
        s = self.state
        npits, ncarves = s["npits"], s["ncarves"]
        P1, C1 = npits + 1, ncarves + 1
        dv = self.dev
        wm, sx, cw = [], [], []
        with torch.cuda.device(dv):
            for i, b in enumerate(self.bands):
                w_ = torch.empty(P1, dtype=torch.float32, device=dv) if wall is not None else None
                s_ = torch.empty(P1, dtype=torch.float32, device=dv) if sink is not None else None
                c_ = torch.empty(C1, dtype=torch.int64, device=dv)
                _lib.check(self.lib.hc_band_select_partials_dev(
                    self.ctx, self._p(b["fid"]), self._p(b["carve_bound"]), b["fid"].numel(), npits, ncarves,
                    self._p(wall[i]) if wall is not None else None, self._p(water[i]),
                    self._p(sink[i]) if sink is not None else None, self._p(w_), self._p(s_), self._p(c_), self._stream()),
                    "hc_band_select_partials_dev")
                wm.append(w_)
                sx.append(s_)
                cw.append(c_)
            wall_min = self._reduce(wm, "min") if wall is not None else None
            sink_max = self._reduce(sx, "max") if sink is not None else None
            carve_water = self._reduce(cw, "sum")
            sol = torch.zeros(P1, dtype=torch.int32, device=dv)
            _lib.check(self.lib.hc_pit_select_tables_dev(
                self.ctx, npits, self._p(s["maxfill"]), self._p(s["volfill"]), self._p(s["maxcarve"]), self._p(s["volcarve"]),
                self._p(s["volfb"]), self._p(s["maxfb"]), self._p(s["complete"]), self._p(s["nfill"]), self._p(s["ncarvecells"]),
                self._p(s["pit_out"]), self._p(s["pairs"]), self._p(s["included"]), s["npairs"], self._p(wall_min),
                self._p(sink_max), self._p(carve_water), float(fill_vol_thresh), float(carve_vol_thresh), float(max_carve_depth),
                float(maxcarve_factor), int(bool(apply_shallow_fill)), 1 if numpy_scalars == "weak" else 0, self._p(sol),
                self._stream()), "hc_pit_select_tables_dev")
        s["solution"] = sol
        DFpits["solution"] = sol[1:].cpu().numpy().astype(np.int64)
        DFpits["downstream_pit"] = np.zeros(npits, dtype=np.int64)
        return DFpits

    def apply(self, DFpits, fill, carve=None, excludedpits=None):
        """dep.apply_solution on every local band: `fill` = list of the bands' filled surfaces (flat float32 device
        tensors), `carve` defaults to the bands' own carve rasters.  Pits absent from DFpits are left alone (the
        reference loops over DFpits.index only).  -> (list of modified DEM bands, list of int16 solution masks)"""
        import numpy as np
        torch = self.torch
        s = self.state
        npits, ncarves = s["npits"], s["ncarves"]
        dv = self.dev
        solh = np.full(npits + 1, -1, np.int32)
        solh[DFpits.index.to_numpy().astype(np.int64)] = DFpits["solution"].to_numpy().astype(np.int32)
        sol = torch.from_numpy(solh).to(dv)
        excl = None
        if excludedpits is not None:
            excl = torch.zeros(npits + 1, dtype=torch.uint8, device=dv)
            if len(excludedpits):
                excl[torch.as_tensor(list(excludedpits), dtype=torch.int64, device=dv)] = 1
        outs, masks = [], []
        with torch.cuda.device(dv):
            for i, b in enumerate(self.bands):
                n = b["dem"].numel()
                out = torch.empty_like(b["dem"])
                mask = torch.zeros(n, dtype=torch.int16, device=dv)
                cv = b["carve"] if carve is None else carve[i]
                _lib.check(self.lib.hc_pit_apply_dev(
                    self.ctx, self._p(b["dem"]), self._p(fill[i]), self._p(cv), self._p(b["fid"]), self._p(b["carve_bound"]),
                    self._p(b["cfid"]), n, npits, ncarves, self._p(sol), self._p(excl), self._p(s["pit_out"]),
                    self._p(s["ncarvecells"]), self._p(s["pairs"]), self._p(s["included"]), s["npairs"], self._p(out),
                    self._p(mask), self._stream()), "hc_pit_apply_dev")
                outs.append(out)
                masks.append(mask)
        return outs, masks
