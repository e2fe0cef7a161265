"""Pit catalog / solution selection on the GPU vs golden vectors produced by the REFERENCE's own
tools/depressions.py (tests/golden/make_catalog_golden.py; parity pinned).  Integer outputs (solution codes,
masks, completecarve, index lists) and elevations must be identical; the float32 volume columns are
float64-accumulated here vs float32 pairwise sums in numpy: rtol 1e-5 (stated in include/hydrocore.h)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "catalog_reference.npz")
NODATA = -9999.0


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _catalog(g, t):
    from pydrodem_b200.tools import depressions as dep
    v = {k: np.ravel(g[f"{t}_{k}"]) for k in ("z", "fill", "carve", "diff_fill", "diff_carve", "fillID", "carveID", "carvefillID")}
    res = dep.build_pit_catalog(v["z"], NODATA, v["carve"], v["diff_carve"], v["diff_fill"], v["fillID"], v["carveID"],
                                v["carvefillID"], verbose_print=False, carveout=False)
    return dep, v, res


@pytest.mark.parametrize("t", ["a", "b"])
def test_build_pit_catalog(gold, t, cuda):
    dep, v, (DF, fillcells, carveonly, carvecells, carvefill) = _catalog(gold, t)
    assert list(DF.columns) == ["maxfill", "volfill", "maxcarve", "volcarve", "volfillbehind", "maxfillbehind", "completecarve"]
    assert DF.index[0] == 1 and len(DF) == int(gold[f"{t}_fillID"].max())
    for col in ("maxfill", "maxcarve", "maxfillbehind", "completecarve"):
        assert np.array_equal(DF[col].to_numpy(), gold[f"{t}_cat_{col}"]), col
    for col in ("volfill", "volcarve", "volfillbehind"):
        np.testing.assert_allclose(DF[col].to_numpy(), gold[f"{t}_cat_{col}"], rtol=1e-5, atol=1e-6, err_msg=col)
    assert DF.maxfill.dtype == np.float32 and DF.completecarve.dtype == np.int8
    for p in gold[f"{t}_probe"]:
        p = int(p)
        assert np.array_equal(fillcells[p], gold[f"{t}_fillcells_{p}"])
        assert np.array_equal(np.sort(carvecells[p]), np.sort(gold[f"{t}_carvecells_{p}"]))
        assert np.array_equal(carvefill[p], gold[f"{t}_carvefill_{p}"])


@pytest.mark.parametrize("t", ["a", "b"])
@pytest.mark.parametrize("variant", ["sel", "selshallow", "sel2", "sel3"])
def test_select_and_apply_solution(gold, t, variant, cuda):
    dep, v, (DF, fillcells, carveonly, carvecells, carvefill) = _catalog(gold, t)
    thr = gold[f"{t}_thresholds" if variant in ("sel", "selshallow") else f"{t}_thresholds_{variant}"]
    use_masks = variant == "sel"
    shape = gold[f"{t}_z"].shape
    D2 = dep.select_solution(DF, fillcells, carvecells, carveonly, v["fill"], v["carve"], v["fillID"], v["carveID"],
                             float(thr[0]), float(thr[1]), float(thr[2]), maxcarve_factor=1.25,
                             apply_shallow_fill=(variant == "selshallow"),
                             wall_array=gold[f"{t}_wall"] if use_masks else None, water_mask_array=gold[f"{t}_water"],
                             sink_array=gold[f"{t}_sink"] if use_masks else None)
    want = gold[f"{t}_{variant}_solution"]
    assert np.array_equal(D2["solution"].to_numpy(), want), np.flatnonzero(D2["solution"].to_numpy() != want)[:10]
    assert (D2["downstream_pit"] == 0).all()
    D3 = D2[D2.solution < 3000]
    mod, mask = dep.apply_solution(D3, fillcells, carvecells, carvefill, v["z"], v["fill"], v["carve"])
    assert mod.dtype == np.float32 and mask.dtype == np.int16
    assert np.array_equal(mask, gold[f"{t}_{variant}_mask"])
    assert np.array_equal(mod, gold[f"{t}_{variant}_mod"]), f"{(mod != gold[f'{t}_{variant}_mod']).sum()} cells differ"
    if variant == "sel":
        ex = [int(p) for p in gold[f"{t}_excluded"]]
        mod2, mask2 = dep.apply_solution(D3, fillcells, carvecells, carvefill, v["z"], v["fill"], v["carve"],
                                         solution_mask_vec=mask.copy(), excludedpits=ex)
        assert np.array_equal(mask2, gold[f"{t}_sel_mask_excl"])
        assert np.array_equal(mod2, gold[f"{t}_sel_mod_excl"])
    assert shape[0] * shape[1] == mod.size


@pytest.mark.parametrize("t", ["a", "b"])
def test_fill_shallow_pits(gold, t, cuda):
    from pydrodem_b200.tools import depressions as dep
    z, fid = np.ravel(gold[f"{t}_z"]), np.ravel(gold[f"{t}_fillID"])
    for sfill, use_water in ((0.3, True), (1.5, False)):
        got = dep.fill_shallow_pits(z.copy(), sfill, fid, gold[f"{t}_diff_fill"], gold[f"{t}_fill"],
                                    water_vec=np.ravel(gold[f"{t}_water"]) if use_water else None, verbose_print=False)
        assert np.array_equal(got, gold[f"{t}_shallow_{int(sfill*10)}_{int(use_water)}"])


def test_select_without_water_mask_raises_like_the_reference(gold, cuda):
    dep, v, (DF, fillcells, carveonly, carvecells, carvefill) = _catalog(gold, "a")
    with pytest.raises(NameError):
        dep.select_solution(DF, fillcells, carvecells, carveonly, v["fill"], v["carve"], v["fillID"], v["carveID"], 1.0, 1.0, 1.0)


def test_seg_stats_large_random(cuda):
    """hc_seg_stats on 4M cells / 5000 labels vs numpy (exact for max/min/count, 1e-9 for the float64 sums)."""
    import ctypes
    import torch
    from pydrodem_b200 import _lib
    rng = np.random.default_rng(5)
    n, nseg = 4_000_000, 5000
    ids = np.repeat(rng.integers(0, nseg + 1, n // 50), 50).astype(np.int32)   # runs of 50 like real labels
    vals = rng.normal(0, 10, n).astype(np.float32)
    ti, tv = torch.from_numpy(ids).cuda(), torch.from_numpy(vals).cuda()
    mx = torch.empty(nseg + 1, dtype=torch.float32, device="cuda")
    mn = torch.empty_like(mx)
    sm = torch.empty(nseg + 1, dtype=torch.float64, device="cuda")
    ct = torch.empty(nseg + 1, dtype=torch.int64, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    _lib.check(_lib.load().hc_seg_stats_dev(_lib.context(0), p(ti), p(tv), None, n, nseg, 0, p(mx), p(mn), p(sm), p(ct), None),
               "hc_seg_stats_dev")
    torch.cuda.synchronize()
    cnt = np.bincount(ids, minlength=nseg + 1)
    assert np.array_equal(ct.cpu().numpy(), cnt)
    wmx = np.full(nseg + 1, -np.inf, np.float32)
    wmn = np.full(nseg + 1, np.inf, np.float32)
    np.maximum.at(wmx, ids, vals)
    np.minimum.at(wmn, ids, vals)
    assert np.array_equal(mx.cpu().numpy(), wmx) and np.array_equal(mn.cpu().numpy(), wmn)
    np.testing.assert_allclose(sm.cpu().numpy(), np.bincount(ids, weights=vals.astype(np.float64), minlength=nseg + 1), rtol=1e-9, atol=1e-6)
