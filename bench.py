#!/usr/bin/env python
"""bench.py — Mcells/s for fill + D8 (+flat resolution) + accumulation (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload ...]

One "step" = one pass of the hot path (PitRemove -> D8FlowDir -> AreaD8 semantics, run_aws.py:580-617) over one batch of
synthetic input that is already resident in HBM.

  N = 1 (default)  workload C2 = BASELINE.json configs[1]: one synthetic 10801 x 10801 float32 DEM with random pits, through
                   hc_fill_d8_accum_ld_dev on a pitched device raster (core.empty_raster: the TMA / vector path).
  N > 1 (torchrun) workload C3 = configs[2]: ONE synthetic 40000 x 40000 DEM row-band-sharded over the N GPUs, halo and
                   seam-graph exchange over NCCL (strong scaling; the headline line).  The same line carries, under
                   config.weak_bands, the weak-scaling number (one 10801 x 10801 band per GPU of one 10801 x 10801 N DEM)
                   and, under config.checks, whether the gathered bands equal the whole-raster chain bit for bit.
  --workload c1    configs[0]: the per-unit depression pipeline (DepHandleCore.run_depression_core) on a 3601^2 DEM.
  --workload c4 / c5 / bands / c3 / c2: the other configurations as extra measured lines.

`e2e` times the same work through the host-buffer entry points (pinned host arrays in, results out, copies inside the timed
region).  `roofline` describes the dominant FULL-GRID kernel (one launch sweeps the raster once); kernels that run many
sparse launches per step are reported with their total time per step in `kernels`, never as a per-launch fraction.

--impl reference times the CPU path on the host cores: the oracle port of the WhiteboxTools / TauDEM algorithms (the
binaries pydrodem shells out to are not installable here), farmed over min(cores, 14) independent 3601^2 units the way
the reference farms units over a process pool (models/depression_handling.py:1625-1635).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mcells/s fill+D8+accum"
UNIT = "Mcells/s"
BYTES_PER_CELL = {"fill": 8, "d8": 5, "flats": 0, "accum": 5}  # 18 B/cell in total
KERNEL_PASS = {
    "k_ptr": "fill", "k_perim": "fill", "k_flatten": "fill", "k_minedge": "fill", "k_tab_init": "fill",
    "k_hook": "fill", "k_mutual": "fill", "k_jump": "fill", "k_merge": "fill", "k_repflat": "fill",
    "k_levels": "fill", "k_apply": "fill", "k_nodata_mask": "fill", "k_lab_tile": "label", "k_lab_border": "label",
    "k_lab_compress": "fill", "k_ext_flag_border": "fill", "k_ext_out": "fill",
    "k_d8": "d8", "k_flat_init": "flats", "k_flat_relax": "flats", "k_flat_dirs": "flats",
    "k_acc_init": "accum", "k_acc_walk": "accum", "k_acc_out": "accum",
    "k_minedge_emit": "fill", "k_minedge_list": "fill", "k_term_edges": "fill", "k_seam_seed_edges": "fill",
    "k_pick_both": "fill", "k_hook_idx": "fill", "k_compact_marked": "fill", "k_cross_edges": "fill",
    "k_unpack_edges": "fill", "k_level_final": "fill",
    "k_flat_fast": "flats", "k_flat_dilate": "flats", "k_flat_get_seams": "flats", "k_flat_put_halo": "flats",
    "k_acc_tile_local": "accum", "k_acc_tile_final": "accum", "k_acc_link_count": "accum", "k_acc_link_mark": "accum",
    "k_acc_link_walk": "accum", "k_acc_seam_exit": "accum", "k_seam_init": "accum", "k_seam_count": "accum",
    "k_seam_mark": "accum", "k_seam_walk": "accum", "k_acc_seam_apply": "accum", "k_acc_seam_inject": "accum",
    "k_edges": "fill", "k_bor_persistent": "fill", "k_apply_d8": "fill+d8", "k_flat_list": "flats",
    "k_flat_fast_sparse": "flats", "k_flat_queue_tiles": "flats", "k_pad_nodata": "fill",
}
# kernels whose ONE launch sweeps the whole raster once: the only candidates for the per-kernel roofline line, with the
# algorithmic bytes per cell of what that launch reads and writes (z 4, filled 4, dir 1, acc 4; the
# watershed-label raster P is internal and counted at its 4 B)
FULL_GRID = {
    "k_ptr": 8,             # z in, P out
    "k_edges": 8,           # z + P in
    "k_minedge": 8, "k_minedge_emit": 8,
    "k_apply": 12,          # z + P in, filled out
    "k_apply_d8": 14,       # z + P in, filled + 2 x dir out
    "k_d8": 5,              # filled in, dir out
    "k_flat_fast": 5, "k_flat_list": 1,
    "k_acc_tile_local": 5,  # dir in, acc out
    "k_acc_tile_final": 9,  # dir + acc in, acc out
    "k_term_edges": 8,
}
BYTES_PER_CELL["fill+d8"] = 13


def dominant_roofline(prof, steps, cells, step_ms, peak, peak_src, traffic_of=None):
    """roofline line for the dominant full-grid kernel: achieved = its algorithmic bytes per launch / its average launch
    time (launches per step is 1 for these kernels).  prof: name -> (launch count, total ms, max ms)."""
    cand = {k: v for k, v in prof.items() if k in FULL_GRID and v[0] > 0}
    if not cand:
        return None
    name, (cnt, tot, mx) = max(cand.items(), key=lambda kv: kv[1][1])
    avg_ms = tot / cnt
    alg = FULL_GRID[name] * cells
    ach = alg / (avg_ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": name, "pass": KERNEL_PASS.get(name, "other"), "achieved": ach, "peak": peak,
            "unit": "GB/s", "frac": ach / peak, "traffic": traffic_of(name) if traffic_of else None,
            "peak_source": peak_src, "launches_per_step": cnt / steps, "avg_launch_ms": avg_ms,
            "algorithmic_bytes_per_launch": alg, "algorithmic_bytes_per_cell": FULL_GRID[name],
            "step_share": tot / steps / step_ms,
            "note": "dominant FULL-GRID kernel; multi-launch sparse kernels (k_flat_relax, k_bor_persistent) are listed "
                    "with their total time per step under `kernels` and have no per-launch fraction"}


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture of this workload (profiles/ncu_traffic.json),
    or None.  Only meaningful for the C2 single-GPU workload the capture was taken on."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            return float(json.load(f)["kernels"][kernel]["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        if int(os.environ.get("RANK", "0")) != 0:
            return   # one sampler per job: eight nvidia-smi pollers contend with the ranks for the driver and the cores
        try:
            # the sampler must not share the rank's pinned cores (a child inherits the affinity mask)
            def _unpin():
                try:
                    os.sched_setaffinity(0, set(range(os.cpu_count() or 1)))
                except (AttributeError, OSError):
                    pass
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"],
                                      stdout=self.f, stderr=subprocess.DEVNULL, preexec_fn=_unpin)
        except OSError:
            self.p = None

    def stop(self, window=None):
        """window = (t0, t1) in time.time() seconds: the timed region.  Samples are taken every 20 ms from start() on (the
        sampler is started before the warm-up so that it is up when the timed region begins); `sm_mhz` is the median over
        the samples INSIDE the window when there are any (else over all samples under the same load), `samples` says how
        many those were and `samples_total` how many were taken in all."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "samples_total": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                sm, mx = float(parts[2]), float(parts[3])
            except ValueError:
                continue
            rows.append((ts, sm, mx, [nm for nm, val in zip(names, parts[6:10]) if val.lower().startswith("active")]))
        self.f.close()
        os.unlink(self.f.name)
        if not rows:
            return out
        inside = [r for r in rows if window and window[0] - 0.02 <= r[0] <= window[1] + 0.02]
        use = inside if inside else rows
        sms = sorted(r[1] for r in use)
        reasons = set()
        for r in use:
            reasons.update(r[3])
        out.update(sm_mhz=sms[len(sms) // 2], sm_max_mhz=max(r[2] for r in use), reasons=sorted(reasons),
                   samples=len(inside), samples_total=len(rows),
                   window="timed region" if inside else "warm-up + timed region (no sample fell inside the timed region)")
        return out


def cpu_chain_rate(z, threads=1):
    """Oracle fill+D8+flats+accum on numpy DEM(s); returns (Mcells/s, seconds)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import hydro_oracle as O
    O.build()
    zs = z if isinstance(z, list) else [z]
    t0 = time.perf_counter()
    if threads == 1:
        for a in zs:
            O.fill_d8_accum(a, dx=30.0, dy=30.0, seed_mode=O.SEED_TAUDEM, table=O.TABLE_TAUDEM)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda a: O.fill_d8_accum(a, dx=30.0, dy=30.0, seed_mode=O.SEED_TAUDEM,
                                                  table=O.TABLE_TAUDEM), zs))
    dt = time.perf_counter() - t0
    cells = sum(a.size for a in zs)
    return cells / dt / 1e6, dt


def run_reference(args):
    """CPU arm: unit farm over the host cores, 3601^2 units (BASELINE.json configs[0] size)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from pydrodem_b200 import synth
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 14))
    side = args.ref_size
    units = [synth.make_dem_numpy(side, side, 1002 + i) for i in range(threads)]
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu_chain_rate(units, threads)
    tot_cells, tot_t = 0, 0.0
    steps = args.steps
    for _ in range(steps):
        r, dt = cpu_chain_rate(units, threads)
        tot_cells += sum(u.size for u in units)
        tot_t += dt
    value = tot_cells / tot_t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": tot_t / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"C2 recipe (fBm + random pits + bowls, 1 cm), {threads} units of {side}x{side} per step, "
                               "fill+D8+flats+accum (TauDEM semantics)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{threads} x {side}^2 units per step over a {threads}-thread pool "
                                   f"({cores} logical cores on the box); oracle port of WBT/TauDEM algorithms"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def raster_hash(t, chunk_rows=512):
    """order-sensitive 64-bit hash of a 2-D device raster (int64 wrap-around arithmetic), computed in row chunks"""
    import torch
    h = torch.zeros((), dtype=torch.int64, device=t.device)
    ny, nx = t.shape
    base = 0
    for r0 in range(0, ny, chunk_rows):
        x = t[r0:r0 + chunk_rows].contiguous()
        v = x.view(torch.int32) if x.element_size() == 4 else x.to(torch.int32)
        v = v.reshape(-1).to(torch.int64)
        w = (torch.arange(base, base + v.numel(), device=t.device, dtype=torch.int64) % 1000003) + 1
        h = h + (v * w).sum()
        base += v.numel()
    return h


def measure_banded(args, world, rank, local, dev, workload, steps, warmup, want_e2e, want_check):
    """One DEM cut into row bands, a band per GPU (BASELINE.json configs[2]); seam exchange over NCCL.  Returns the bench
    line (a dict) on every rank; only rank 0's is printed."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pydrodem_b200 import _lib, bands, synth
    import pydrodem_b200 as hc

    nl = args.local_bands
    nbands = world * nl
    if workload == "c3":
        nx, ny_total = 40000, 40000
        rows_all = bands.split_rows(ny_total, nbands)
        scaling = "strong"
        wl = f"C3: one synthetic {ny_total}x{nx} float32 DEM row-band-sharded over {nbands} band(s)"
    else:
        nx = args.size
        rows_all = [args.size] * nbands
        ny_total = args.size * nbands
        scaling = "weak"
        wl = (f"C2 recipe as ONE {ny_total}x{nx} float32 DEM row-band-sharded, a {args.size}x{nx} band per GPU "
              f"({nbands} band(s))")
    first = rank * nl
    my_rows = rows_all[first:first + nl]
    row0s = [sum(rows_all[:first + i]) for i in range(nl)]
    seams = [sum(rows_all[:i]) for i in range(1, nbands)]
    comm = bands.SeamComm()
    chain = bands.BandedChain(my_rows, nx, nbands, first, dev, comm=comm, dx=30.0, dy=30.0, flats=bool(args.flats),
                              want_slope=False)
    n_bowls = max(200, int(200 * ny_total * nx / (10801.0 * 10801.0)))
    dem_kw = dict(n_bowls=n_bowls, seam_rows=seams, plateau_rows=seams[:1])
    for i, nb in enumerate(my_rows):
        zv = chain.z_view(i)
        for r0 in range(0, nb, 512):
            nr = min(512, nb - r0)
            zv[r0:r0 + nr] = synth.make_dem(ny_total, nx, 1003, row0=row0s[i] + r0, nrows=nr, device=dev, **dem_kw)
    torch.cuda.synchronize()
    cells_local = sum(my_rows) * nx
    cells_total = ny_total * nx

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    for _ in range(warmup):
        out = chain.run()
    barrier()
    for c in chain.ctx:
        _lib.profile_enable(True, ctx=c)
    launches0 = chain.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.time()
    e0.record()
    for _ in range(steps):
        out = chain.run()
    e1.record()
    barrier()
    tw1 = time.time()
    clocks = sampler.stop((tw0, tw1))
    ms = e0.elapsed_time(e1)
    launches = chain.launches() - launches0
    prof = {}
    for c in chain.ctx:
        for k, (cnt, tot, mx) in _lib.profile_read(ctx=c).items():
            a = prof.get(k, (0, 0.0, 0.0))
            prof[k] = (a[0] + cnt, a[1] + tot, max(a[2], mx))
        _lib.profile_enable(False, ctx=c)
    # per-phase device time (includes the host read-backs each phase waits on), 2 extra untimed steps
    chain.timing = True
    for _ in range(2):
        chain.run()
    chain.timing = False
    phases = {k: v / 2 for k, v in chain.phase_ms.items()}
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    per_rank = [ms / steps]
    if world > 1:
        allms = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allms, t)
        per_rank = [float(x.item()) / steps for x in allms]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = cells_total * steps / (ms_max * 1e-3) / 1e6
    stats = chain.stats()
    flat_rounds = chain.flat_rounds

    # size-independent checks on the full-size result (tests/run_bands_nccl.py checks the flow identity too)
    tot_in = torch.zeros(3, dtype=torch.float64, device=dev)
    for i, (f, d, s_, a) in enumerate(out):
        z = chain.z_view(i)
        tot_in[0] += (f < z).sum()                    # fill never lowers a cell
        tot_in[1] += (d == 0).sum()                   # cells left without a direction (undrainable flats)
        tot_in[2] += ((d != 255) & (a < 1)).sum()     # every cell with a direction counts itself
    if world > 1:
        dist.all_reduce(tot_in)
    chk = {"cells_below_input": int(tot_in[0].item()), "noflow_cells": int(tot_in[1].item()),
           "valid_cells_with_zero_count": int(tot_in[2].item())}

    # bit-equality with the WHOLE-RASTER chain: every rank hashes its bands' results; rank 0 runs the whole DEM through
    # hc_fill_d8_accum on its own GPU and hashes the same row ranges
    if want_check:
        hb = torch.stack([torch.stack([raster_hash(f), raster_hash(d), raster_hash(a)]) for (f, d, s_, a) in out])  # [nl, 3]
        if world > 1:
            allh = [torch.zeros_like(hb) for _ in range(world)]
            dist.all_gather(allh, hb)
            hb_all = torch.cat(allh)
        else:
            hb_all = hb
        equal = None
        if rank == 0:
            try:
                zf = hc.empty_raster(ny_total, nx, torch.float32, dev)
                for r0 in range(0, ny_total, 512):
                    nr = min(512, ny_total - r0)
                    zf[r0:r0 + nr] = synth.make_dem(ny_total, nx, 1003, row0=r0, nrows=nr, device=dev, **dem_kw)
                wf, wd, ws_, wa = hc.fill_d8_accum(zf, dx=30.0, dy=30.0, seed_mode=hc.SEED_TAUDEM, table=hc.TABLE_TAUDEM,
                                                   flats=bool(args.flats), want_slope=False)
                torch.cuda.synchronize()
                equal = True
                r0 = 0
                for g, nb in enumerate(rows_all):
                    want = torch.stack([raster_hash(wf[r0:r0 + nb]), raster_hash(wd[r0:r0 + nb]), raster_hash(wa[r0:r0 + nb])])
                    if not torch.equal(want, hb_all[g].to(want.device)):
                        equal = False
                    r0 += nb
                del zf, wf, wd, wa
                torch.cuda.empty_cache()
            except torch.cuda.OutOfMemoryError:
                equal = "skipped: the whole raster did not fit beside the band on rank 0"
        chk["bands_equal_whole_raster_chain"] = equal
        if world > 1:
            dist.barrier()

    # ---- end to end: pinned host band in, filled + dir + acc out, copies overlapped with the next unit's compute ----
    e2e = None
    if want_e2e:
        hz = [torch.empty((nb, nx), dtype=torch.float32, pin_memory=True) for nb in my_rows]
        # two result sets (host pinned + device staging): unit k's results cross PCIe while unit k + 1 is computed
        hf = [[torch.empty((nb, nx), dtype=torch.float32, pin_memory=True) for nb in my_rows] for _ in range(2)]
        hd = [[torch.empty((nb, nx), dtype=torch.uint8, pin_memory=True) for nb in my_rows] for _ in range(2)]
        ha = [[torch.empty((nb, nx), dtype=torch.int32, pin_memory=True) for nb in my_rows] for _ in range(2)]
        sf = [[torch.empty((nb, nx), dtype=torch.float32, device=dev) for nb in my_rows] for _ in range(2)]
        sd = [[torch.empty((nb, nx), dtype=torch.uint8, device=dev) for nb in my_rows] for _ in range(2)]
        sa = [[torch.empty((nb, nx), dtype=torch.int32, device=dev) for nb in my_rows] for _ in range(2)]
        sz = [[torch.empty((nb, nx), dtype=torch.float32, device=dev) for nb in my_rows] for _ in range(2)]
        for i in range(nl):
            hz[i].copy_(chain.z_view(i))
        torch.cuda.synchronize()
        main = torch.cuda.current_stream(dev)
        s_up, s_down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        ev_up = [torch.cuda.Event() for _ in range(2)]
        ev_stage = [torch.cuda.Event() for _ in range(2)]
        ev_down = [torch.cuda.Event() for _ in range(2)]
        ev_zfree = [torch.cuda.Event() for _ in range(2)]

        def upload(k):
            with torch.cuda.stream(s_up):
                s_up.wait_event(ev_zfree[k & 1])
                for i in range(nl):
                    sz[k & 1][i].copy_(hz[i], non_blocking=True)
                ev_up[k & 1].record(s_up)

        def compute(k):
            main.wait_event(ev_up[k & 1])
            for i in range(nl):
                chain.z_view(i).copy_(sz[k & 1][i], non_blocking=True)
            ev_zfree[k & 1].record(main)
            res = chain.run()
            main.wait_event(ev_down[k & 1])       # the staging set's previous download has finished
            for i, (f, d, s_, a) in enumerate(res):
                sf[k & 1][i].copy_(f, non_blocking=True)
                sd[k & 1][i].copy_(d, non_blocking=True)
                sa[k & 1][i].copy_(a, non_blocking=True)
            ev_stage[k & 1].record(main)

        def download(k):
            with torch.cuda.stream(s_down):
                s_down.wait_event(ev_stage[k & 1])
                for i in range(nl):
                    hf[k & 1][i].copy_(sf[k & 1][i], non_blocking=True)
                    hd[k & 1][i].copy_(sd[k & 1][i], non_blocking=True)
                    ha[k & 1][i].copy_(sa[k & 1][i], non_blocking=True)
                ev_down[k & 1].record(s_down)

        def stream_units(n):
            for e in ev_zfree + ev_down:
                e.record(main)
            upload(0)
            for k in range(n):
                if k + 1 < n:
                    upload(k + 1)
                compute(k)
                download(k)
            torch.cuda.synchronize()
        stream_units(2)
        ksteps = max(3, min(steps, 6))
        barrier()
        t0 = time.perf_counter()
        stream_units(ksteps)
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        same = all(torch.equal(hf[(ksteps - 1) & 1][i], out[i][0].cpu()) and torch.equal(ha[(ksteps - 1) & 1][i], out[i][3].cpu())
                   for i in range(nl))
        e2e = {"value": cells_total * ksteps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": cells_total * 4,
               "d2h_bytes_per_step": cells_total * 9, "steps": ksteps, "host_results_equal_device_results": bool(same),
               "api": "bands.BandedChain.run on a stream of units: pinned host band in, filled+dir+acc out per rank, uploads "
                      "and downloads of neighbouring units overlap the chain (two staging sets)"}
        del hz, hf, hd, ha, sf, sd, sa, sz

    peak, peak_src = peaks()
    cells = cells_local  # rank 0's share, for the per-kernel roofline
    pass_ms = {}
    for k, (cnt, tot, mx) in prof.items():
        pass_ms[KERNEL_PASS.get(k, "other")] = pass_ms.get(KERNEL_PASS.get(k, "other"), 0.0) + tot / steps
    roof = dominant_roofline(prof, steps * nl, cells // nl, ms_max / steps, peak, peak_src)
    kernels = {k: {"launches_per_step": c / steps, "ms_per_step": t_ / steps} for k, (c, t_, m) in
               sorted(prof.items(), key=lambda kv: -kv[1][1])}
    passes = {p_: {"ms_per_step": v,
                   "hbm_frac_algorithmic": (BYTES_PER_CELL.get(p_, 0) * cells / (v * 1e-3) / 1e9 / peak) if v > 0 else None}
              for p_, v in pass_ms.items()}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl + "; fill + D8 + flat resolution + accumulation, TauDEM semantics; bowls centred on "
                                    "every seam and one plateau across the first seam; halo + seam-graph exchange by "
                                    "all_gather over NCCL",
                   "cells_total": cells_total, "cells_per_gpu": cells_local, "bands": nbands,
                   "l2": "every band's rasters exceed the 126 MB L2; no explicit flush", "flats": bool(args.flats),
                   "flat_exchange_rounds": flat_rounds, "band_stats_rank0": stats, "phase_ms_rank0": phases, "ms_per_step_per_rank": per_rank,
                   "algorithmic_bytes_per_cell": 18, "checks": chk},
        "roofline": roof, "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "hbm_frac_algorithmic_whole_step": 18 * cells_local / (ms_max / steps * 1e-3) / 1e9 / peak,
        "passes": passes, "kernels": kernels,
    }
    chain.close()
    del chain, out
    torch.cuda.empty_cache()
    return line


def run_banded(args, world, rank, local, dev):
    """N > 1 (or --workload bands / c3): the headline line is C3 (40000^2, strong); workload 'auto' adds the weak-bands
    measurement of the same chain under config.weak_bands."""
    import torch.distributed as dist
    auto = args.workload == "auto_n"
    wl = "c3" if auto else args.workload
    line = measure_banded(args, world, rank, local, dev, wl, args.steps, args.warmup, not args.no_e2e, True)
    if auto:
        weak = measure_banded(args, world, rank, local, dev, "bands", max(3, min(args.steps, 10)), 3, False, False)
        line["config"]["weak_bands"] = {k: weak[k] for k in ("value", "unit", "ms_per_step", "scaling", "steps")}
        line["config"]["weak_bands"].update(workload=weak["config"]["workload"], cells_total=weak["config"]["cells_total"],
                                            ms_per_step_per_rank=weak["config"]["ms_per_step_per_rank"],
                                            phase_ms_rank0=weak["config"]["phase_ms_rank0"], clocks=weak["clocks"],
                                            checks=weak["config"]["checks"])
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_c1(args, world, rank, local, dev):
    """BASELINE.json configs[0]: the per-unit depression pipeline (DepHandle.depression_handling's core + the final fill,
    models/depression_handling.py:717-830, 470-484) on a synthetic 3601^2 tile, GeoTIFF in, conditioned DEM out, the
    reference's file protocol between the stages (all-zero water mask: the only configuration the reference completes).
    One unit per GPU.  Extra measured line; value = cells / unit time."""
    import shutil
    import numpy as np
    import torch
    import torch.distributed as dist
    from pydrodem_b200 import _lib, raster_io, synth
    from pydrodem_b200.models.depression_handling import DepHandleCore
    import pydrodem_b200.tools.depressions as dep
    n = 3601 if args.size == 10801 else args.size
    z = synth.make_dem_numpy(n, n, 1001 + rank, nodata_frame=True) if n <= 1200 else \
        synth.make_dem(n, n, 1001 + rank, nodata_frame=True, device=dev).cpu().numpy()
    stage = {}
    for name in ("fill_pits", "create_clump_array", "fill_shallow_pits", "breach_pits", "build_pit_catalog",
                 "select_solution", "apply_solution", "combined_solutions"):
        fn = getattr(dep, name)

        def wrap(*a, __fn=fn, __name=name, **k):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            try:
                return __fn(*a, **k)
            finally:
                torch.cuda.synchronize()
                stage[__name] = stage.get(__name, 0.0) + time.perf_counter() - t0
        setattr(dep, name, wrap)
    times, codes, pits = [], {}, 0
    launches0 = _lib.launch_count(local)
    reps = max(1, min(args.steps, 3))
    sampler = ClockSampler(local)
    for it in range(1 + reps):
        tmp = tempfile.mkdtemp(prefix="hc_c1_")
        try:
            tif = os.path.join(tmp, "DEM.tif")
            raster_io.write_raster(tif, z, None, nodata=-9999.0)
            core = DepHandleCore(basename="DEM", del_temp_files=False)
            if it == 1:
                stage.clear()
                sampler.start()
                launches0 = _lib.launch_count(local)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ffill, mod, DFpits, solmask = core.run_depression_core(tif, tmp, water_mask_array=np.zeros((n, n), np.uint8))
            torch.cuda.synchronize()
            if it >= 1:
                times.append(time.perf_counter() - t0)
            pits = int(len(DFpits))
            vals, cnts = np.unique(DFpits.solution.values, return_counts=True)
            codes = {str(int(v)): int(c) for v, c in zip(vals, cnts)}
        finally:
            shutil.rmtree(tmp, ignore_errors=True)
    clocks = sampler.stop()
    launches = _lib.launch_count(local) - launches0
    dt = sum(times) / len(times)
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    line = {"metric": "Mcells/s depression_handling unit (configs[0])", "value": world * n * n / dt / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": reps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C1: {n}x{n} float32 1-arcsec tile through DepHandleCore.run_depression_core (initial_fill_carve -> "
                                   "catalog -> select -> apply -> combined solutions -> final fill), GeoTIFF files between the "
                                   "stages as in the reference, reference ini parameters, all-zero water mask; one unit per GPU",
                       "pits": pits, "solution_codes": codes,
                       "stage_seconds": {k: round(v / reps, 4) for k, v in sorted(stage.items(), key=lambda kv: -kv[1])}},
            "roofline": None, "cpu_baseline": None, "e2e": {"value": world * n * n / dt / 1e6, "unit": UNIT,
                                                             "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                                                             "api": "the line itself is end to end: files in, files out"},
            "gpu_launches": launches // reps, "clocks": clocks}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_label(args, world, rank, local, dev):
    """Connected-component labelling (scipy.ndimage.label's numbering; dep.create_clump_array, tools/depressions.py:
    1646-1669) on a 10801^2 random mask at the 8-connected site-percolation threshold (p = 0.407: components of every size
    up to raster-spanning ones).  Algorithmic bytes: 1 (mask in) + 4 (int32 labels out) = 5 B/cell."""
    import torch
    import pydrodem_b200 as hc
    from pydrodem_b200 import _lib
    n = args.size
    g = torch.Generator(device=dev).manual_seed(4077 + rank)
    mask = (torch.rand((n, n), device=dev, generator=g) < 0.407).to(torch.uint8)
    for _ in range(args.warmup):
        lab, nlab = hc.label(mask, 8)
    torch.cuda.synchronize()
    _lib.profile_enable(True, local)
    launches0 = _lib.launch_count(local)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lab, nlab = hc.label(mask, 8)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _lib.launch_count(local) - launches0
    prof = _lib.profile_read(local)
    _lib.profile_enable(False, local)
    peak, peak_src = peaks()
    cells = n * n
    kernels = {k: {"launches_per_step": c / args.steps, "ms_per_step": t_ / args.steps} for k, (c, t_, m) in
               sorted(prof.items(), key=lambda kv: -kv[1][1])}
    name, (cnt, tot, mx) = max(prof.items(), key=lambda kv: kv[1][1])
    per_launch_bytes = {"k_lab_tile": 5, "k_lab_compress": 8, "k_lab_out": 12, "k_root_rank": 8, "k_root_count": 4, "k_lab_border": 5}
    b = per_launch_bytes.get(name, 5) * cells
    roof = {"bound": "hbm", "kernel": name, "pass": "label", "achieved": b / (tot / cnt * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": b / (tot / cnt * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "avg_launch_ms": tot / cnt,
            "algorithmic_bytes_per_launch": b, "launches_per_step": cnt / args.steps}
    line = {"metric": "Mcells/s connected-component labelling", "value": cells / (ms * 1e-3) / 1e6, "unit": UNIT, "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8->i32", "data": "synthetic",
            "config": {"workload": f"label: {n}x{n} uint8 random mask, density 0.407 (8-connected site-percolation threshold), "
                                   "8-connectivity, scipy.ndimage.label numbering", "components": int(nlab),
                       "algorithmic_bytes_per_cell": 5},
            "roofline": roof, "cpu_baseline": None, "e2e": None, "gpu_launches": launches // args.steps, "clocks": clocks,
            "hbm_frac_algorithmic_whole_step": 5 * cells / (ms * 1e-3) / 1e9 / peak, "kernels": kernels}
    if args.cpu_sample > 0:
        from scipy import ndimage
        import numpy as np
        s_ = min(3000, n)
        crop = mask[:s_, :s_].cpu().numpy()
        t0 = time.perf_counter()
        ndimage.label(crop, structure=np.ones((3, 3)))
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": crop.size / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "reference", "seconds": dt,
                                "sample": f"scipy.ndimage.label (the reference's own call) on the top-left {s_}x{s_} crop"}
    print(json.dumps(line))
    return 0


def run_c5(args, world, rank, local, dev):
    """BASELINE.json configs[4]: dem_noise_removal smoothing core (4 DW smoothings 9/7/5/3, slope_D4 + curvature_D2 of
    the 5x5 result, global sigma, filter selection, final slope; models/dem_noise_removal.py:397-476) on a 20000^2 DEM,
    all-zero masks (pure bandwidth).  One unit per GPU.  Not the driver's headline metric: an extra measured line."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pydrodem_b200 import _lib, stencils, synth
    n = 20000 if args.size == 10801 else args.size
    cells = n * n
    z = torch.empty((n, n), dtype=torch.float32, device=dev)
    for r0 in range(0, n, 1024):
        nr = min(1024, n - r0)
        z[r0:r0 + nr] = synth.make_dem(n, n, 1005 + rank, row0=r0, nrows=nr, device=dev, quantise=False)
    z += torch.randn(z.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(1005 + rank)) * 1.5
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        out = stencils.smooth_core(z, 30.0, 30.0)
    barrier()
    _lib.profile_enable(True, local)
    launches0 = _lib.launch_count(local)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = stencils.smooth_core(z, 30.0, 30.0)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count(local) - launches0
    prof = _lib.profile_read(local)
    _lib.profile_enable(False, local)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * cells * args.steps / (ms_max * 1e-3) / 1e6
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peak, peak_src = peaks()
    dom = max(prof.items(), key=lambda kv: kv[1][1]) if prof else None
    roof = None
    if dom:
        name, (cnt, tot, mx) = dom
        avg_ms = tot / cnt
        alg = 8 * cells  # one k x k smoothing pass: 4 B in + 4 B out per cell
        ach = alg / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "pass": "smoothing", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src, "launches_per_step": cnt / args.steps,
                "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg,
                "step_share": tot / args.steps / (ms_max / args.steps)}
    cpu = None
    if world == 1 and args.cpu_sample > 0:
        from scipy.signal import convolve2d
        s = min(2000, n)
        crop = np.ascontiguousarray(z[:s, :s].cpu().numpy())
        t0 = time.perf_counter()
        for w in (9, 7, 5, 3):
            k = stencils.build_kernel(stencils.DW_WEIGHTS[w], dtype=np.float32, normalize=True)
            convolve2d(crop, k, boundary="symm", mode="same")   # the reference's own call (tools/derive.py:94)
        gy, gx = np.gradient(crop, 30.0, 30.0)
        dt = time.perf_counter() - t0
        cpu = {"value": crop.size / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
               "sample": f"{s}x{s} crop: the four scipy.signal.convolve2d(boundary='symm') smoothings + one np.gradient, the "
                         "library calls tools/derive.py makes (selection / slopes not included: an upper bound on the CPU rate)"}
    kernels = {k: {"launches_per_step": c / args.steps, "ms_per_step": t_ / args.steps} for k, (c, t_, m) in
               sorted(prof.items(), key=lambda kv: -kv[1][1])}
    line = {"metric": "Mcells/s DEM-NR smoothing core (configs[4])", "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C5: {n}x{n} float32 DEM, DW 9/7/5/3 smoothing + slope_D4 + curvature_D2 + sigma + filter "
                                   "selection + final slope (CreateDTM.run_tile smoothing core), all-zero masks, one unit per GPU",
                       "cells_per_gpu": cells, "l2": "1.6 GB rasters exceed the 126 MB L2; no explicit flush",
                       "algorithmic_bytes_per_cell_fused_pipeline": 12},
            "roofline": roof, "cpu_baseline": cpu, "e2e": None, "gpu_launches": launches, "clocks": clocks,
            "hbm_frac_algorithmic_whole_step": 12 * cells / (ms_max / args.steps * 1e-3) / 1e9 / peak, "kernels": kernels}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=10801, help="DEM edge (default: configs[1] = 10801)")
    ap.add_argument("--ref-size", type=int, default=3601)
    ap.add_argument("--cpu-sample", type=int, default=5400, help="edge of the crop timed on the CPU (0 = skip)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--flats", type=int, default=1)
    ap.add_argument("--workload", default="auto", choices=["auto", "c1", "c2", "bands", "c3", "c4", "c5", "label"],
                    help="auto: c2 (one 10801^2 DEM) at N=1; at N>1 c3 (the 40000^2 DEM row-band-sharded over the N GPUs, "
                         "strong scaling) as the headline with the weak-scaling bands run (one 10801 x 10801*N DEM, a band "
                         "per GPU) reported inside the same line; c1: the per-unit depression pipeline on 3601^2")
    ap.add_argument("--dense", action="store_true", help="C2 / C4 on a dense (unpitched) device raster: the scalar-loader path")
    ap.add_argument("--local-bands", type=int, default=1, help="bands per rank (N=1 with >1 emulates the exchange)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    import pydrodem_b200 as hc
    from pydrodem_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # one rank per GPU, each with its own pair of host cores: every seam exchange waits for the slowest rank, so a
        # rank that gets descheduled at one of its host read-backs stalls all of them
        try:
            ncpu = os.cpu_count() or 1
            per = max(1, ncpu // world)
            os.sched_setaffinity(0, set(range(local * per, min(ncpu, (local + 1) * per))))
        except (AttributeError, OSError):
            pass
    if world > 1:
        # keep stdout to the ONE JSON line: NCCL prints its version banner to stdout when the first communicator comes up,
        # so file descriptor 1 points at stderr until that has happened
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    if args.workload == "auto":
        args.workload = "c2" if world == 1 and args.local_bands == 1 else ("auto_n" if world > 1 else "bands")
    if args.workload in ("bands", "c3", "auto_n"):
        return run_banded(args, world, rank, local, dev)
    if args.workload == "c1":
        return run_c1(args, world, rank, local, dev)
    if args.workload == "label":
        return run_label(args, world, rank, local, dev)
    if args.workload == "c5":
        return run_c5(args, world, rank, local, dev)

    ny = nx = args.size
    cells = ny * nx
    # one unit per GPU (weak scaling): same recipe, per-rank seed
    # pitched device raster (rows padded to 64 elements): TMA-staged tile windows and aligned vector loads; --dense
    # keeps the caller-style dense raster, which takes the kernels' scalar loaders
    z = torch.empty((ny, nx), dtype=torch.float32, device=dev) if args.dense else hc.empty_raster(ny, nx, torch.float32, dev)
    band = 1024
    if args.workload == "c4":
        # configs[3]: flattened lakes + burned rivers (large exact flats: flat resolution heavy)
        z0, _water = synth.make_c4(ny, nx, 1004 + rank, device=dev)
        z.copy_(z0)
        del z0, _water
    else:
        for r0 in range(0, ny, band):
            nr = min(band, ny - r0)
            z[r0:r0 + nr] = synth.make_dem(ny, nx, 1002 + rank, row0=r0, nrows=nr, device=dev)
    torch.cuda.synchronize()

    kw = dict(dx=30.0, dy=30.0, seed_mode=hc.SEED_TAUDEM, table=hc.TABLE_TAUDEM, flats=bool(args.flats),
              want_slope=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)        # nvidia-smi needs a moment to come up; it then samples every 20 ms
    for _ in range(args.warmup):
        out = hc.fill_d8_accum(z, **kw)
    barrier()
    launches0 = _lib.launch_count(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    tw0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = hc.fill_d8_accum(z, **kw)
    e1.record()
    barrier()
    tw1 = time.time()
    clocks = sampler.stop((tw0, tw1))
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count(local) - launches0
    # per-kernel times: a profiled pass of the same steps right after the timed region (the library brackets every launch
    # with CUDA events on the launching stream; the chain runs its kernels in sequence on one stream, so each kernel's
    # events time that kernel alone)
    _lib.profile_enable(True, local)
    psteps = max(3, min(args.steps, 10))
    for _ in range(psteps):
        hc.fill_d8_accum(z, **kw)
    torch.cuda.synchronize()
    prof = {k: (c * args.steps / psteps, t_ * args.steps / psteps, m) for k, (c, t_, m) in _lib.profile_read(local).items()}
    _lib.profile_enable(False, local)
    basins, rounds = hc.core.fill_stats(local)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * cells * args.steps / (ms_max * 1e-3) / 1e6

    # ---- end to end through the host-buffer C ABI ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        import ctypes
        lib = _lib.load()
        ctx = _lib.context(local)
        hz = torch.empty((ny, nx), dtype=torch.float32, pin_memory=True)
        hz.copy_(z)
        # two sets of pinned result buffers: a stream of units is processed with two in flight (begin / end API), so
        # one unit's results cross PCIe while the next unit is uploaded and computed
        hf = [torch.empty((ny, nx), dtype=torch.float32, pin_memory=True) for _ in range(2)]
        hd = [torch.empty((ny, nx), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
        ha = [torch.empty((ny, nx), dtype=torch.int32, pin_memory=True) for _ in range(2)]
        torch.cuda.synchronize()

        def host_call():
            _lib.check(lib.hc_fill_d8_accum_host(ctx, hz.data_ptr(), hf[0].data_ptr(), hd[0].data_ptr(), None,
                                                 ha[0].data_ptr(), ny, nx, -9999.0, 30.0, 30.0, hc.SEED_TAUDEM,
                                                 hc.TABLE_TAUDEM, int(bool(args.flats))), "hc_fill_d8_accum_host")

        def begin(slot):
            _lib.check(lib.hc_fill_d8_accum_host_begin(ctx, slot, hz.data_ptr(), hf[slot].data_ptr(), hd[slot].data_ptr(),
                                                       None, ha[slot].data_ptr(), ny, nx, -9999.0, 30.0, 30.0,
                                                       hc.SEED_TAUDEM, hc.TABLE_TAUDEM, int(bool(args.flats))),
                       "hc_fill_d8_accum_host_begin")

        def end(slot):
            _lib.check(lib.hc_fill_d8_accum_host_end(ctx, slot), "hc_fill_d8_accum_host_end")
        host_call()
        begin(1)
        end(1)
        ksteps = max(2, min(args.steps, 5))
        # (a) one synchronous call per unit
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            host_call()
        torch.cuda.synchronize()
        dt_sync = time.perf_counter() - t0
        # (b) the same units as a stream, two in flight
        ksteps = max(4, min(args.steps, 10))
        hf[0].zero_(), hf[1].zero_(), ha[0].zero_(), ha[1].zero_()
        barrier()
        t0 = time.perf_counter()
        begin(0)
        for i in range(1, ksteps):
            begin(i & 1)
            end((i - 1) & 1)
        end((ksteps - 1) & 1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, dt_sync], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt, dt_sync = float(tt[0].item()), float(tt[1].item())
        e2e = {"value": world * cells * ksteps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": cells * 4,
               "d2h_bytes_per_step": cells * 9, "steps": ksteps,
               "api": "hc_fill_d8_accum_host_begin/_end, two units in flight (pinned host buffers in, filled+dir+acc "
                      "out, every unit uploaded and downloaded inside the timed region)",
               "single_call_value": world * cells * max(2, min(args.steps, 5)) / dt_sync / 1e6,
               "single_call_api": "hc_fill_d8_accum_host (one synchronous call per unit)"}
        hf, hd, ha = hf[1], hd[1], ha[1]
        # the host path must give the same answer as the resident path
        assert torch.equal(hf, out[0].cpu()) and torch.equal(hd, out[1].cpu()) and torch.equal(ha, out[3].cpu())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel + per-pass table ---------------------------------------
    peak, peak_src = peaks()
    pass_ms = {}
    for k, (cnt, tot, mx) in prof.items():
        pass_ms[KERNEL_PASS.get(k, "other")] = pass_ms.get(KERNEL_PASS.get(k, "other"), 0.0) + tot / args.steps
    roof = dominant_roofline(prof, args.steps, cells, ms_max / args.steps, peak, peak_src,
                             traffic_of=ncu_traffic if ny == 10801 else None)
    if roof is not None:
        roof["traffic_source"] = "profiles/ncu_traffic.json (ncu dram__bytes_read+write per launch)"
    kernels = {k: {"launches_per_step": c / args.steps, "ms_per_step": t_ / args.steps} for k, (c, t_, m) in
               sorted(prof.items(), key=lambda kv: -kv[1][1])}
    passes = {p: {"ms_per_step": v,
                  "hbm_frac_algorithmic": (BYTES_PER_CELL.get(p, 0) * cells / (v * 1e-3) / 1e9 / peak) if v > 0 else None}
              for p, v in pass_ms.items()}

    cpu = None
    if world == 1 and args.cpu_sample > 0:
        s = min(args.cpu_sample, ny)
        crop = np.ascontiguousarray(z[:s, :s].cpu().numpy())
        rate, secs = cpu_chain_rate(crop, 1)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port", "seconds": secs,
               "sample": f"top-left {s}x{s} crop of the same DEM, same chain, single-thread oracle port "
                         f"(WBT/TauDEM binaries unavailable); host has {os.cpu_count()} logical cores"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"C4: synthetic {ny}x{nx} float32 DEM with flattened lakes and burned rivers (large exact "
                                "flats) on the C2 terrain" if args.workload == "c4" else
                                f"C2: synthetic {ny}x{nx} float32 DEM with random pits (0.5% cells + 200 bowls, 1 cm "
                                "quantised)") + ", fill + D8 + flat resolution + accumulation, TauDEM semantics "
                               "(PitRemove/D8FlowDir/AreaD8), one unit per GPU",
                   "cells_per_gpu": cells, "l2": "inputs (467 MB z + outputs) exceed the 126 MB L2; no explicit flush",
                   "device_raster": "dense" if args.dense else "pitched (core.empty_raster, hc_fill_d8_accum_ld_dev): TMA path",
                   "flats": bool(args.flats), "basins": basins, "boruvka_rounds": rounds,
                   "kernel_times": "profiled pass right after the timed region (same inputs, per-launch CUDA events); "
                                   "the timed region itself runs unprofiled",
                   "algorithmic_bytes_per_cell": 18},
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "hbm_frac_algorithmic_whole_step": 18 * cells / (ms_max / args.steps * 1e-3) / 1e9 / peak,
        "passes": passes, "kernels": kernels,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
