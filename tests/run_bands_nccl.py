#!/usr/bin/env python
"""Multi-process check of the row-band path over NCCL (one band per GPU).  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tests/run_bands_nccl.py --ny 3000 --nx 2000 [--full]

Every rank builds the same seeded DEM, keeps its band, runs bands.BandedChain with the NCCL SeamComm;
rank 0 gathers the bands, runs the whole-raster chain on the full DEM and compares bit for bit.
--full adds the size-independent flow identity  acc(c) = 1 + sum of acc over the cells draining into c
evaluated band-wise with exchanged halo rows (what a 40000^2 run is checked with, no oracle needed).
Prints BANDS_NCCL_OK on success.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DR = [0, -1, -1, -1, 0, 1, 1, 1]   # TauDEM table, codes 1..8
DC = [1, 1, 0, -1, -1, -1, 0, 1]


def flow_identity_violations(chain, res):
    """Cells of this rank's bands where acc != 1 + inflow (or != 0 on 255 cells)."""
    import torch
    bad = 0
    nx = chain.nx
    accs, dirs = [], []
    for i, (f, d, s, a) in enumerate(res):
        nb = d.shape[0]
        ap = torch.zeros((nb + 2, nx), dtype=torch.int32, device=d.device)
        dp = torch.full((nb + 2, nx), 255, dtype=torch.uint8, device=d.device)
        ap[1:-1] = a
        dp[1:-1] = d
        accs.append(ap)
        dirs.append(dp)
    chain._halo_exchange(accs)
    chain._halo_exchange(dirs)
    for i, (f, d, s, a) in enumerate(res):
        nb = d.shape[0]
        ap = torch.nn.functional.pad(accs[i].to(torch.int64), (1, 1))
        dp = torch.nn.functional.pad(dirs[i], (1, 1), value=255)
        inflow = torch.zeros((nb, nx), dtype=torch.int64, device=d.device)
        for k in range(8):
            r0, c0 = 1 - DR[k], 1 - DC[k]
            src = dp[r0:r0 + nb, c0:c0 + nx] == (k + 1)
            inflow += torch.where(src, ap[r0:r0 + nb, c0:c0 + nx], torch.zeros((), dtype=torch.int64, device=d.device))
        a64 = a.to(torch.int64)
        valid = d != 255
        bad += int(((a64 != 1 + inflow) & valid).sum().item()) + int(((a64 != 0) & ~valid).sum().item())
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ny", type=int, default=3000)
    ap.add_argument("--nx", type=int, default=2000)
    ap.add_argument("--seed", type=int, default=1003)
    ap.add_argument("--full", action="store_true")
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import pydrodem_b200 as hc
    from pydrodem_b200 import bands, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rows = bands.split_rows(args.ny, world)
    row0 = sum(rows[:rank])
    seams = [sum(rows[:i]) for i in range(1, world)]
    z = synth.make_dem(args.ny, args.nx, args.seed, device=dev, seam_rows=seams, plateau_rows=seams[:1],
                       n_bowls=max(20, args.ny * args.nx // 500000))
    chain = bands.BandedChain([rows[rank]], args.nx, world, rank, dev, comm=bands.SeamComm(), dx=30.0, dy=30.0,
                              want_slope=True)
    res = chain.run([z[row0:row0 + rows[rank]]])
    torch.cuda.synchronize()
    bad = flow_identity_violations(chain, res) if args.full else 0
    # gather the bands on rank 0 (padded to the tallest band)
    mx = max(rows)
    ok = True
    parts = []
    for t, fillv in ((res[0][0], 0.0), (res[0][1], 0), (res[0][2], 0.0), (res[0][3], 0)):
        pad = torch.full((mx, args.nx), fillv, dtype=t.dtype, device=dev)
        pad[:rows[rank]] = t
        if world > 1:
            outs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
            dist.gather(pad, outs, dst=0)
        else:
            outs = [pad]
        if rank == 0:
            parts.append(torch.cat([o[:rows[i]] for i, o in enumerate(outs)]))
    if rank == 0:
        f, d, s, a = hc.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=hc.SEED_TAUDEM, table=hc.TABLE_TAUDEM)
        torch.cuda.synchronize()
        for name, got, want in (("fill", parts[0], f), ("dir", parts[1], d), ("slope", parts[2], s), ("acc", parts[3], a)):
            n = int((got != want).sum().item())
            if n:
                ok = False
                print(f"MISMATCH {name}: {n} cells differ", flush=True)
    tb = torch.tensor([bad], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tb)
    if rank == 0:
        if int(tb.item()):
            ok = False
            print(f"flow identity violated at {int(tb.item())} cells", flush=True)
        print(f"bands={world} dem={args.ny}x{args.nx} flat_rounds={chain.flat_rounds} stats={chain.stats()}", flush=True)
        print("BANDS_NCCL_OK" if ok else "BANDS_NCCL_FAIL", flush=True)
    chain.close()
    if world > 1:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
