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
