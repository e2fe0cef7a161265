#!/usr/bin/env python
"""Multi-process check of the banded pit catalog / selection / apply over NCCL (one band per GPU).  Launch:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29534 \
        tests/run_catalog_nccl.py

Every rank loads the golden scene the REFERENCE's own tools/depressions.py produced (tests/golden/catalog_reference.npz),
keeps its row band, runs bands.BandedCatalog with the NCCL all_reduces; every rank compares the (replicated) catalog and
solution codes with the golden vectors, rank 0 gathers the applied surface and compares it too.  Prints CATALOG_NCCL_OK."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    from pydrodem_b200 import bands
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    g = np.load(os.path.join(ROOT, "tests", "golden", "catalog_reference.npz"))
    ok = True
    for t in ("a", "b"):
        ny, nx = g[f"{t}_z"].shape
        rows = bands.split_rows(ny, world)
        r0 = sum(rows[:rank])
        nb = rows[rank]

        def dev(name, dt):
            return torch.from_numpy(np.ascontiguousarray(np.asarray(g[f"{t}_{name}"])[r0:r0 + nb]).reshape(-1)).cuda().to(dt).contiguous()
        b = dict(dem=dev("z", torch.float32), carve=dev("carve", torch.float32), dcarve=dev("diff_carve", torch.float32),
                 dfill=dev("diff_fill", torch.float32), fid=dev("fillID", torch.int32), cid=dev("carveID", torch.int32),
                 cfid=dev("carvefillID", torch.int32), cell0=r0 * nx)
        cat = bands.BandedCatalog([b], -9999.0)
        DF = cat.build()
        for col in ("maxfill", "maxcarve", "maxfillbehind", "completecarve"):
            ok &= bool(np.array_equal(DF[col].to_numpy(), g[f"{t}_cat_{col}"]))
        for col in ("volfill", "volcarve", "volfillbehind"):
            ok &= bool(np.allclose(DF[col].to_numpy(), g[f"{t}_cat_{col}"], rtol=1e-5, atol=1e-6))
        thr = g[f"{t}_thresholds"]
        D2 = cat.select(DF, float(thr[0]), float(thr[1]), float(thr[2]), wall=[dev("wall", torch.float32)],
                        water=[dev("water", torch.float32)], sink=[dev("sink", torch.float32)])
        ok &= bool(np.array_equal(D2["solution"].to_numpy(), g[f"{t}_sel_solution"]))
        outs, masks = cat.apply(D2[D2.solution < 3000], [dev("fill", torch.float32)])
        want = torch.from_numpy(np.ascontiguousarray(g[f"{t}_sel_mod"].reshape(ny, nx)[r0:r0 + nb]).reshape(-1)).cuda()
        ok &= bool(torch.equal(outs[0], want))
        wantm = torch.from_numpy(np.ascontiguousarray(g[f"{t}_sel_mask"].reshape(ny, nx)[r0:r0 + nb]).reshape(-1)).cuda()
        ok &= bool(torch.equal(masks[0], wantm))
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("CATALOG_NCCL_OK" if int(flag.item()) == 1 else "CATALOG_NCCL_MISMATCH", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
