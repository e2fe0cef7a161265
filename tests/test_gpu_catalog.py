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


@pytest.mark.gpu
def test_raise_lake_connected_pits_documented_semantics(tmp_path):
    """tools/depressions.py:1534-1590 as documented: pits holding lake cells are raised to (lowest lake cell in the pit)
    + 1 cm where they lie below it; pits without lake cells and cells outside pits stay; nothing happens with <= 20 lake
    cells.  (Upstream's own intersection is accidental -- see the function's docstring -- so the check is against the
    definition, written out in numpy.)"""
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools import depressions as dep
    rng = np.random.default_rng(3)
    ny, nx = 90, 120
    dem = (rng.random((ny, nx)) * 5 + 100).astype(np.float32)
    fid = np.zeros((ny, nx), np.int32)
    fid[10:30, 10:40] = 1
    fid[40:70, 20:60] = 2
    fid[75:85, 90:110] = 3                      # no lake inside
    water = np.zeros((ny, nx), np.uint8)
    water[15:22, 15:30] = 1                     # inside pit 1
    water[50:60, 30:50] = 1                     # inside pit 2
    water[2:5, 100:110] = 1                     # outside every pit
    water[60, 5] = 2                            # a river code: not a lake
    dem[water == 1] += np.float32(2.5)          # water surfaces stand above much of the pits' ground
    mod0 = dem.copy()
    mod0[12, 12] += 50.0                        # an earlier stage already raised this cell: must be kept
    want = mod0.ravel().copy()
    d = dem.ravel()
    for pit in (1, 2, 3):
        cells = np.flatnonzero(fid.ravel() == pit)
        lake_cells = cells[water.ravel()[cells] == 1]
        if lake_cells.size == 0:
            continue
        lvl = d[lake_cells].min()
        below = cells[d[cells] < lvl]
        want[below] = lvl + np.float32(0.01)
    got, diff = dep.raise_lake_connected_pits(water, fid.ravel(), mod0.ravel().copy(), dem.ravel(), verbose_print=False)
    assert np.array_equal(got, want) and np.array_equal(diff, dem.ravel() - want)
    assert (got != mod0.ravel()).sum() > 50 and got[12 * nx + 12] in (mod0[12, 12], want[12 * nx + 12])
    assert np.array_equal(got.reshape(ny, nx)[75:85, 90:110], mod0[75:85, 90:110])
    # the same through a GeoTIFF path, as initial_fill_carve passes it
    p = str(tmp_path / "lakes.tif")
    raster_io.write_raster(p, water, None, nodata=0)
    got2, _ = dep.raise_lake_connected_pits(p, fid.ravel(), mod0.ravel().copy(), dem.ravel(), verbose_print=False)
    assert np.array_equal(got2, want)
    few = np.zeros((ny, nx), np.uint8)
    few[15:17, 15:20] = 1                       # 10 lake cells: below the threshold of 20
    got3, diff3 = dep.raise_lake_connected_pits(few, fid.ravel(), mod0.ravel().copy(), dem.ravel(), verbose_print=False)
    assert np.array_equal(got3, mod0.ravel())


GOLD_C4 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "catalog_c4_reference.npz")


@pytest.mark.gpu
def test_c4_recipe_catalog_selection_with_water_mask_matches_reference(oracle, cuda):
    """BASELINE.json configs[3] at 1201^2 (synth.make_c4: flattened lakes, burned rivers and their water mask): the
    reference's own build_pit_catalog / select_solution / apply_solution / fill_shallow_pits outputs
    (tests/golden/make_catalog_c4_golden.py, 5254 pits, two threshold sets) against the GPU segmented reductions."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_catalog_c4_golden import derive
    from pydrodem_b200.tools import depressions as dep
    g = np.load(GOLD_C4)
    z, water = g["z"], g["water"]
    carve = z.copy().ravel()
    carve[g["carve_idx"]] = g["carve_val"]
    carve = carve.reshape(z.shape)
    d = derive(z, carve)                       # oracle fill + scipy labels: what initial_fill_carve derives
    v = {k: np.ravel(a) for k, a in d.items()}
    DF, fillcells, carveonly, carvecells, carvefill = dep.build_pit_catalog(
        z.ravel(), NODATA, carve.ravel(), v["diff_carve"], v["diff_fill"], v["fillID"], v["carveID"], v["carvefillID"],
        verbose_print=False, carveout=False)
    assert len(DF) == len(g["cat_maxfill"]) == 5254
    for col in ("maxfill", "maxcarve", "maxfillbehind", "completecarve"):
        assert np.array_equal(DF[col].to_numpy(), g[f"cat_{col}"]), col
    for col in ("volfill", "volcarve", "volfillbehind"):
        np.testing.assert_allclose(DF[col].to_numpy(), g[f"cat_{col}"], rtol=1e-5, atol=1e-5, err_msg=col)
    for tag in ("a", "b"):
        fvt, cvt, mcd = (float(x) for x in g[f"{tag}_thresholds"])
        D2 = dep.select_solution(DF.copy(), fillcells, carvecells, carveonly, v["fill"], carve.ravel(), v["fillID"], v["carveID"],
                                 fvt, cvt, mcd, maxcarve_factor=1.25, apply_shallow_fill=False, wall_array=None,
                                 water_mask_array=water, sink_array=None)
        got = D2["solution"].to_numpy()
        want = g[f"{tag}_solution"]
        assert np.array_equal(got, want), (tag, np.flatnonzero(got != want)[:10], got[got != want][:10], want[got != want][:10])
        D3 = D2[D2.solution < 3000]
        mod, mask = dep.apply_solution(D3, fillcells, carvecells, carvefill, z.ravel().copy(), v["fill"], carve.ravel())
        want_mod = z.ravel().copy()
        want_mod[g[f"{tag}_mod_idx"]] = g[f"{tag}_mod_val"]
        want_mask = np.zeros(z.size, np.int16)
        want_mask[g[f"{tag}_mask_idx"]] = g[f"{tag}_mask_val"]
        assert np.array_equal(np.asarray(mod), want_mod) and np.array_equal(np.asarray(mask), want_mask), tag
    shallow = dep.fill_shallow_pits(z.ravel().copy(), 1.0, v["fillID"], d["diff_fill"], d["fill"], water_vec=water.ravel(),
                                    verbose_print=False)
    want_sh = z.ravel().copy()
    want_sh[g["shallow_idx"]] = g["shallow_val"]
    assert np.array_equal(np.asarray(shallow), want_sh)
