"""Pit catalog / selection / apply over ROW BANDS (per-label partial max / sum reduced across the bands,
pydrodem_b200/bands.py BandedCatalog) against the golden vectors the REFERENCE's own tools/depressions.py produced
(tests/golden/make_catalog_golden.py) -- the same vectors the whole-raster calls are held to (tests/test_gpu_catalog.py).
N bands on one GPU emulate N ranks: the reductions between a process's bands and between ranks are the same operations."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "catalog_reference.npz")
NODATA = -9999.0


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _bands(g, t, nbands, extra=()):
    import torch
    from pydrodem_b200 import bands
    ny, nx = g[f"{t}_z"].shape
    rows = bands.split_rows(ny, nbands)
    r0s = [sum(rows[:i]) for i in range(nbands)]
    keys = dict(dem="z", carve="carve", dcarve="diff_carve", dfill="diff_fill", fid="fillID", cid="carveID", cfid="carvefillID")

    def dev(a, r0, nb, dt):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(a)[r0:r0 + nb]).reshape(-1)).cuda().to(dt).contiguous()
    out, ex = [], {e: [] for e in extra}
    for r0, nb in zip(r0s, rows):
        b = {k: dev(g[f"{t}_{v}"], r0, nb, torch.int32 if k in ("fid", "cid", "cfid") else torch.float32) for k, v in keys.items()}
        b["cell0"] = r0 * nx
        out.append(b)
        for e in extra:
            ex[e].append(dev(g[f"{t}_{e}"], r0, nb, torch.float32))
    return out, ex


@pytest.mark.parametrize("t", ["a", "b"])
@pytest.mark.parametrize("nbands", [1, 2, 3, 7])
def test_banded_catalog_equals_the_reference(gold, t, nbands, cuda):
    from pydrodem_b200 import bands
    bl, _ = _bands(gold, t, nbands)
    DF = bands.BandedCatalog(bl, NODATA).build()
    assert list(DF.columns) == ["maxfill", "volfill", "maxcarve", "volcarve", "volfillbehind", "maxfillbehind", "completecarve"]
    assert len(DF) == int(gold[f"{t}_fillID"].max())
    for col in ("maxfill", "maxcarve", "maxfillbehind", "completecarve"):
        assert np.array_equal(DF[col].to_numpy(), gold[f"{t}_cat_{col}"]), col
    for col in ("volfill", "volcarve", "volfillbehind"):
        np.testing.assert_allclose(DF[col].to_numpy(), gold[f"{t}_cat_{col}"], rtol=1e-5, atol=1e-6, err_msg=col)


@pytest.mark.parametrize("t", ["a", "b"])
@pytest.mark.parametrize("variant,nbands", [("sel", 3), ("selshallow", 2), ("sel2", 5), ("sel3", 4)])
def test_banded_select_and_apply_equal_the_reference(gold, t, variant, nbands, cuda):
    import torch
    from pydrodem_b200 import bands
    use_masks = variant == "sel"
    bl, ex = _bands(gold, t, nbands, extra=("wall", "water", "sink", "fill"))
    cat = bands.BandedCatalog(bl, NODATA)
    DF = cat.build()
    thr = gold[f"{t}_thresholds" if variant in ("sel", "selshallow") else f"{t}_thresholds_{variant}"]
    D2 = cat.select(DF, float(thr[0]), float(thr[1]), float(thr[2]), maxcarve_factor=1.25,
                    apply_shallow_fill=(variant == "selshallow"), wall=ex["wall"] if use_masks else None, water=ex["water"],
                    sink=ex["sink"] if use_masks else None)
    want = gold[f"{t}_{variant}_solution"]
    assert np.array_equal(D2["solution"].to_numpy(), want), np.flatnonzero(D2["solution"].to_numpy() != want)[:10]
    D3 = D2[D2.solution < 3000]
    outs, masks = cat.apply(D3, ex["fill"])
    mod = torch.cat(outs).cpu().numpy()
    mask = torch.cat(masks).cpu().numpy()
    assert np.array_equal(mask, gold[f"{t}_{variant}_mask"])
    assert np.array_equal(mod, gold[f"{t}_{variant}_mod"]), f"{(mod != gold[f'{t}_{variant}_mod']).sum()} cells differ"
    if variant == "sel":
        exl = [int(p) for p in gold[f"{t}_excluded"]]
        outs2, _ = cat.apply(D3, ex["fill"], excludedpits=exl)
        assert np.array_equal(torch.cat(outs2).cpu().numpy(), gold[f"{t}_sel_mod_excl"])


def test_banded_catalog_over_nccl_two_ranks(cuda):
    """One band per GPU with the real NCCL all_reduces (needs >= 2 GPUs; the single-GPU box skips)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests/test_gpu_catalog_bands.py -k nccl)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", os.path.join(root, "tests", "run_catalog_nccl.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "CATALOG_NCCL_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
