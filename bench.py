#!/usr/bin/env python
"""bench.py — Mcells/s for fill + D8 (+flat resolution) + accumulation on the C2 workload.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path (hc_fill_d8_accum: PitRemove -> D8FlowDir -> AreaD8 semantics,
run_aws.py:580-617) over one synthetic 10801 x 10801 float32 DEM with random pits (BASELINE.json
configs[1]) that is already resident in HBM.  `e2e` times the same call through the HOST-buffer C
ABI (hc_fill_d8_accum_host) with pinned host arrays, H2D/D2H inside the timed region.  N > 1 runs
one such unit per GPU (weak scaling, no data-path collective), timed as max over ranks.

--impl reference times the CPU path on the host cores: the oracle port of the WhiteboxTools /
TauDEM algorithms (the binaries pydrodem shells out to are not installable here), farmed over
min(cores, 14) independent 3601^2 units the way the reference farms units over a process pool
(models/depression_handling.py:1625-1635).
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
    "k_levels": "fill", "k_apply": "fill", "k_nodata_mask": "fill", "k_lab_init": "fill", "k_lab_merge": "fill",
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
}


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
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
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
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL, preexec_fn=_unpin)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            s = sorted(sm)
            out.update(sm_mhz=s[len(s) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
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


def run_banded(args, world, rank, local, dev):
    """One DEM cut into row bands, a band per GPU (BASELINE.json configs[2]); seam exchange over NCCL."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from pydrodem_b200 import _lib, bands, synth
    import pydrodem_b200 as hc

    nl = args.local_bands
    nbands = world * nl
    if args.workload == "c3":
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
    for i, nb in enumerate(my_rows):
        zv = chain.z_view(i)
        for r0 in range(0, nb, 512):
            nr = min(512, nb - r0)
            zv[r0:r0 + nr] = synth.make_dem(ny_total, nx, 1003, row0=row0s[i] + r0, nrows=nr, device=dev,
                                            n_bowls=n_bowls, seam_rows=seams, plateau_rows=seams[:1])
    torch.cuda.synchronize()
    cells_local = sum(my_rows) * nx
    cells_total = ny_total * nx

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    out = None
    for _ in range(args.warmup):
        out = chain.run()
    barrier()
    for c in chain.ctx:
        _lib.profile_enable(True, ctx=c)
    launches0 = chain.launches()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = chain.run()
    e1.record()
    barrier()
    clocks = sampler.stop()
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
    per_rank = [ms / args.steps]
    if world > 1:
        allms = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allms, t)
        per_rank = [float(x.item()) / args.steps for x in allms]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = cells_total * args.steps / (ms_max * 1e-3) / 1e6
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

    # ---- end to end: pinned host band in, filled + dir + acc out, through the same public call ----
    e2e = None
    if not args.no_e2e:
        hz = [torch.empty((nb, nx), dtype=torch.float32, pin_memory=True) for nb in my_rows]
        hf = [torch.empty((nb, nx), dtype=torch.float32, pin_memory=True) for nb in my_rows]
        hd = [torch.empty((nb, nx), dtype=torch.uint8, pin_memory=True) for nb in my_rows]
        ha = [torch.empty((nb, nx), dtype=torch.int32, pin_memory=True) for nb in my_rows]
        for i in range(nl):
            hz[i].copy_(chain.z_view(i))
        torch.cuda.synchronize()

        def host_step():
            for i in range(nl):
                chain.z_view(i).copy_(hz[i], non_blocking=True)
            res = chain.run()
            for i, (f, d, s_, a) in enumerate(res):
                hf[i].copy_(f, non_blocking=True)
                hd[i].copy_(d, non_blocking=True)
                ha[i].copy_(a, non_blocking=True)
            torch.cuda.synchronize()
        host_step()
        ksteps = max(2, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            host_step()
        if world > 1:
            dist.barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": cells_total * ksteps / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": cells_total * 4,
               "d2h_bytes_per_step": cells_total * 9, "steps": ksteps,
               "api": "bands.BandedChain.run (pinned host band in, filled+dir+acc out per rank)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = peaks()
    cells = cells_local  # rank 0's share, for the per-kernel roofline
    pass_ms = {}
    for k, (cnt, tot, mx) in prof.items():
        pass_ms[KERNEL_PASS.get(k, "other")] = pass_ms.get(KERNEL_PASS.get(k, "other"), 0.0) + tot / args.steps
    dom = max(prof.items(), key=lambda kv: kv[1][1]) if prof else None
    roof = None
    if dom:
        name, (cnt, tot, mx) = dom
        pname = KERNEL_PASS.get(name, "other")
        avg_ms = tot / cnt
        alg_bytes = BYTES_PER_CELL.get(pname, 0) * (cells // nl)
        if pname == "flats":
            alg_bytes = BYTES_PER_CELL["d8"] * (cells // nl)
        ach = alg_bytes / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "pass": pname, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": peak_src, "launches_per_step": cnt / args.steps,
                "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "step_share": tot / args.steps / (ms_max / args.steps)}
    kernels = {k: {"launches_per_step": c / args.steps, "ms_per_step": t_ / args.steps} for k, (c, t_, m) in
               sorted(prof.items(), key=lambda kv: -kv[1][1])}
    passes = {p_: {"ms_per_step": v,
                   "hbm_frac_algorithmic": (BYTES_PER_CELL.get(p_, 0) * cells / (v * 1e-3) / 1e9 / peak) if v > 0 else None}
              for p_, v in pass_ms.items()}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl + "; fill + D8 + flat resolution + accumulation, TauDEM semantics; bowls centred on "
                                    "every seam and one plateau across the first seam; halo + seam-graph exchange by "
                                    "all_gather over NCCL",
                   "cells_total": cells_total, "cells_per_gpu": cells_local, "bands": nbands,
                   "l2": "every band's rasters exceed the 126 MB L2; no explicit flush", "flats": bool(args.flats),
                   "flat_exchange_rounds": flat_rounds, "band_stats_rank0": stats, "phase_ms_rank0": phases, "ms_per_step_per_rank": per_rank,
                   "algorithmic_bytes_per_cell": 18, "checks": chk},
        "roofline": roof, "cpu_baseline": None, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "hbm_frac_algorithmic_whole_step": 18 * cells_local / (ms_max / args.steps * 1e-3) / 1e9 / peak,
        "passes": passes, "kernels": kernels,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
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
    ap.add_argument("--workload", default="auto", choices=["auto", "c2", "bands", "c3", "c4", "c5"],
                    help="auto: c2 (one 10801^2 DEM) at N=1, bands (one 10801 x 10801*N DEM, a band per GPU, seam "
                         "exchange over NCCL) at N>1; c3: the 40000^2 DEM row-band-sharded over the N GPUs")
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
        # keep stdout to the one JSON line: NCCL's version banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    if args.workload == "auto":
        args.workload = "c2" if world == 1 and args.local_bands == 1 else "bands"
    if args.workload in ("bands", "c3"):
        return run_banded(args, world, rank, local, dev)
    if args.workload == "c5":
        return run_c5(args, world, rank, local, dev)

    ny = nx = args.size
    cells = ny * nx
    # one unit per GPU (weak scaling): same recipe, per-rank seed
    z = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    band = 1024
    if args.workload == "c4":
        # configs[3]: flattened lakes + burned rivers (large exact flats: flat resolution heavy)
        z, _water = synth.make_c4(ny, nx, 1004 + rank, device=dev)
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
    for _ in range(args.warmup):
        out = hc.fill_d8_accum(z, **kw)
    barrier()
    _lib.profile_enable(True, local)
    launches0 = _lib.launch_count(local)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = hc.fill_d8_accum(z, **kw)
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count(local) - launches0
    prof = _lib.profile_read(local)
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
    dom = max(prof.items(), key=lambda kv: kv[1][1]) if prof else None
    roof = None
    if dom:
        name, (cnt, tot, mx) = dom
        pname = KERNEL_PASS.get(name, "other")
        avg_ms = tot / cnt
        alg_bytes = BYTES_PER_CELL.get(pname, 0) * cells
        if pname == "flats":
            alg_bytes = BYTES_PER_CELL["d8"] * cells  # flat resolution is part of producing the D8 raster
        ach = alg_bytes / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": name, "pass": pname, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": ncu_traffic(name) if ny == 10801 else None, "peak_source": peak_src,
                "launches_per_step": cnt / args.steps, "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "step_share": tot / args.steps / (ms_max / args.steps),
                "traffic_source": "profiles/ncu_traffic.json (ncu dram__bytes_read+write per launch)"}
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
                   "flats": bool(args.flats), "basins": basins, "boruvka_rounds": rounds,
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
