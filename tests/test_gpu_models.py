"""Model-level mirrors (`DepHandle`, `TaudemCheck`, `FindSinks`) with the reference's constructor signatures, unit
directory layout, output file names and `-c config -l log` entry points (models/depression_handling.py:1419-1646,
models/taudem_check.py:204-247, models/find_sinks.py:146-247)."""
import json
import os

import numpy as np
import pytest

from pydrodem_b200 import synth

ND = -9999.0

INI = """
[region]
base_unit = tile
[paths]
WBT_path = /nowhere
SWO_vrt = {root}/swo.vrt
OSM_dir = {root}/osm
output_dir_cond = {root}/cond
[outputs]
projectname = demo
dtm_mosaic = {root}/mosaic.tif
dh_work_dir = {root}/dh
[job-options]
run_parallel = False
run_multiprocess = False
[processes-depression-handling]
overwrite = False
verbose_print = False
del_temp_files = False
burn_rivers = False
enforce_monotonicity = True
reservoir_barrier = False
fill_h20_adjacent = False
enforce_endo = False
noflow_walls = False
osm_coast_clip = False
[parameters-depression-handling]
basin_levels = [6]
buffer = 100
sfill = 0.5
fill_method = WL
carve_method = SC
min_dist = False
flat_increment = 0.0
maxcost = 200
radius = 12
fill_vol_thresh = 60
carve_vol_thresh = 25
max_carve_depth = 6
combined_fill_interval = .2
SWO_cutoff = 20
initial_burn = 2.0
SWO_burn = 2.0
burn_cutoff = 6.0
gap_dist = 240
[paths-taudem]
taudem_dir = {root}/taudem
[paths-DH]
output_dir = {root}/dhb
[parameters-DH-basin]
basin_levels = [6]
"""


def _unit(ny, nx, seed):
    rng = np.random.default_rng(seed)
    z = (synth.make_dem_numpy(ny, nx, seed, quantise=False) * 0.05).astype(np.float32)
    z += rng.normal(0, 0.3, z.shape).astype(np.float32)
    yy, xx = np.mgrid[0:ny, 0:nx]
    for _ in range(10):
        cy, cx, rad, d = rng.integers(15, ny - 15), rng.integers(15, nx - 15), rng.integers(3, 12), rng.uniform(1, 8)
        z -= (d * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2.0 * rad * rad))).astype(np.float32)
    z = np.round(z, 2).astype(np.float32)
    z[0, :] = ND
    return z


@pytest.mark.gpu
def test_dephandle_main_config_driven_units(tmp_path):
    import pydrodem_b200 as hc
    from pydrodem_b200 import raster_io
    from pydrodem_b200.models import depression_handling as dh
    root = str(tmp_path)
    cfg = os.path.join(root, "demo.ini")
    with open(cfg, "w") as f:
        f.write(INI.format(root=root))
    mosaic = _unit(300, 340, 3)
    raster_io.write_raster(os.path.join(root, "mosaic.tif"), mosaic, None, nodata=ND)
    os.makedirs(os.path.join(root, "dh", "tile_A1"))
    os.makedirs(os.path.join(root, "dh", "tile_B2"))
    os.makedirs(os.path.join(root, "dh", "tile_C3"))
    raster_io.write_raster(os.path.join(root, "dh", "tile_A1", "inDEM.tif"), mosaic[:180, :200].copy(), None, nodata=ND)
    with open(os.path.join(root, "dh", "tile_B2", "window.json"), "w") as f:
        json.dump({"row0": 0, "row1": 200, "col0": 120, "col1": 340}, f)   # (row 0 is the nodata frame: a unit needs nodata cells)
    with open(os.path.join(root, "dh", "tile_C3", "DEMcompleted.txt"), "w") as f:      # finished earlier: skipped
        f.write("done")
    log = os.path.join(root, "run.log")
    assert dh.main(["-c", cfg, "-l", log]) == 0
    assert "Depression handling COMPLETE" in open(log).read()
    assert not os.path.exists(os.path.join(root, "dh", "tile_C3", "conditionedDEM.tif"))
    for unit, src in (("tile_A1", mosaic[:180, :200]), ("tile_B2", mosaic[0:200, 120:340])):
        d = os.path.join(root, "dh", unit)
        for name in ("conditionedDEM.tif", "conditionedDEM_DIFF.tif", "Water_mask.tif", "DEMcompleted.txt", "DFpits.csv",
                     "inDEM_solution_mask.tif", "inDEM_WL.tif"):
            assert os.path.exists(os.path.join(d, name)), (unit, name)
        assert not os.path.exists(os.path.join(d, "error_log_main.txt"))
        cond, meta = raster_io.read_raster(os.path.join(d, "conditionedDEM.tif"))
        assert cond.shape == src.shape and meta.nodata == ND
        assert np.array_equal(hc.fill(cond, nodata=ND, seed_mode=hc.SEED_WBT), cond)       # drains everywhere
        diff, _ = raster_io.read_raster(os.path.join(d, "conditionedDEM_DIFF.tif"))
        assert np.array_equal(diff, cond - src)
        # the same unit through the array-level core with the ini's parameters
        core = dh.DepHandleCore(basename="inDEM", radius=12, sfill=0.5, fill_vol_thresh=60, carve_vol_thresh=25,
                                max_carve_depth=6, combined_fill_interval=0.2)
        again = tmp_path / ("again_" + unit)
        again.mkdir()
        raster_io.write_raster(str(again / "inDEM.tif"), np.ascontiguousarray(src), None, nodata=ND)
        ff, _, _, _ = core.run_depression_core(str(again / "inDEM.tif"), str(again), np.zeros(src.shape, np.uint8))
        assert np.array_equal(ff, cond)
    # a second run finds nothing left to do
    assert dh.main(["-c", cfg, "-l", log]) == 0
    with pytest.raises(NotImplementedError):
        dh.DepHandle(os.path.join(root, "mosaic.tif"), root, None, "demo", [6], burn_rivers=True)


@pytest.mark.gpu
def test_taudem_check_class_and_main(tmp_path):
    import pydrodem_b200 as hc
    from pydrodem_b200 import raster_io
    from pydrodem_b200.models import taudem_check as tc
    root = str(tmp_path)
    cfg = os.path.join(root, "demo.ini")
    with open(cfg, "w") as f:
        f.write(INI.format(root=root))
    z = _unit(160, 180, 5)
    z[70:80, 60:75] = ND                                    # a hole: its rim is "contaminated" (d8 nodata) in TauDEM's rule
    fel, p, _, _ = hc.fill_d8_accum(z, nodata=ND, dx=30.0, dy=30.0, seed_mode=hc.SEED_TAUDEM, table=hc.TABLE_TAUDEM)
    pdir = os.path.join(root, "taudem", "demo")
    os.makedirs(os.path.join(pdir, "fel.vrtd"))
    os.makedirs(os.path.join(pdir, "p.vrtd"))
    # the reference reads fel / p through GDAL .vrt mosaics; same-grid GeoTIFFs stand in under the same names
    p16 = p.astype(np.int16)
    p16[p == 255] = -32768
    bdir = os.path.join(root, "dhb", "demo", "basins_lvl06")
    flat = np.zeros(z.shape, np.uint8)
    flat[60:90, 50:85] = 3                                  # a flattened lake around the hole
    for bid, with_sink in ((6001, False), (6002, True), (6003, False)):
        d = os.path.join(bdir, f"basin_{bid}")
        os.makedirs(d)
        raster_io.write_raster(os.path.join(d, "taudem_DEM.tif"), fel, None, nodata=ND)
        d8 = p.astype(np.int16)
        d8[p == 255] = 99
        raster_io.write_raster(os.path.join(d, "taudem_d8_flowdir.tif"), d8, None, nodata=99)
        raster_io.write_raster(os.path.join(d, "Water_flatten_mask.tif"), flat if bid != 6003 else np.zeros_like(flat), None, nodata=0)
        if with_sink:
            raster_io.write_raster(os.path.join(d, "endo_sink.tif"), np.zeros(z.shape, np.uint8), None, nodata=0)
    log = os.path.join(root, "tc.log")
    assert tc.main(["-c", cfg, "-l", log]) == 0
    assert "TAUDEM CHECK COMPLETE" in open(log).read()
    want_mask = (p == 255) & (fel != ND) & (flat > 1) & (flat < 6)
    assert want_mask.sum() > 0
    adj = os.path.join(bdir, "basin_6001", "conditionedDEM_taudem_adjust.tif")
    assert os.path.exists(adj)
    out, meta = raster_io.read_raster(adj)
    want = fel.copy()
    water = (flat > 1) & (flat < 6)
    want[water] = want[water] + np.float32(0.10)
    want[fel == ND] = np.nan
    assert np.array_equal(out, want, equal_nan=True)
    assert not os.path.exists(os.path.join(bdir, "basin_6002", "conditionedDEM_taudem_adjust.tif"))   # endorheic: left alone
    assert not os.path.exists(os.path.join(bdir, "basin_6003", "conditionedDEM_taudem_adjust.tif"))   # no water: no problem
    assert not os.path.exists(os.path.join(bdir, "basin_6003", "taudem_DEM.tif"))                     # temporaries erased
    t = tc.TaudemCheck(os.path.join(root, "dhb", "demo"), bdir, os.path.join(root, "taudem"), "demo")
    assert t.filled_DEM_vrt.endswith(os.path.join("demo", "fel.vrtd", "orig.vrt"))


@pytest.mark.gpu
def test_find_sinks_class_and_main(tmp_path):
    from pydrodem_b200 import raster_io
    from pydrodem_b200.models import find_sinks as fsk
    root = str(tmp_path)
    cfg = os.path.join(root, "demo.ini")
    with open(cfg, "w") as f:
        f.write(INI.format(root=root))
    ny, nx = 140, 160
    yy, xx = np.mgrid[0:ny, 0:nx]
    z = (0.02 * ((yy - 90) ** 2 + (xx - 40) ** 2)).astype(np.float32) + 100.0       # a bowl, lowest at (90, 40)
    z[yy + xx < 60] = 99999.0                                                        # outside the basin polygon
    z[20, 120] = 90.0                                                                # a single-cell pit: must not win
    d = os.path.join(root, "cond", "demo", "endo_basins", "basin_4120012345")
    os.makedirs(d)
    raster_io.write_raster(os.path.join(d, "DEM_endo.tif"), z, None, nodata=99999.0)
    log = os.path.join(root, "fs.log")
    assert fsk.main(["-c", cfg, "-l", log]) == 0
    assert "ENDO SINK FINDING COMPLETE" in open(log).read()
    gj = json.load(open(os.path.join(d, "sink.geojson")))
    props = gj["features"][0]["properties"]
    assert (props["row"], props["col"]) == (90, 40) and props["basin_id"] == 4120012345
    mask, _ = raster_io.read_raster(os.path.join(d, "endo_sink.tif"))
    assert mask.sum() == 1 and mask[90, 40] == 1
    assert not os.path.exists(os.path.join(d, "DEM_endo_sc.tif"))
    # finished basins are skipped unless overwrite
    os.remove(os.path.join(d, "endo_sink.tif"))
    assert fsk.main(["-c", cfg, "-l", log]) == 0
    assert not os.path.exists(os.path.join(d, "endo_sink.tif"))
