"""The file-path boundary (WhiteboxToolsGPU duck type, tools.depressions mirror, TauDEM CLI) and the
small kernels around the fill, checked against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = -9999.0


def _dem(seed=3, shape=(180, 140)):
    from pydrodem_b200 import synth
    z = synth.random_dem(shape[0], shape[1], seed, "smooth", nodata_holes=2)
    rng = np.random.default_rng(seed)
    pits = rng.random(z.shape) < 0.03
    z[pits & (z != N)] -= 3.0
    return z


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_single_cell_pit_kernels(oracle, cuda, seed):
    import pydrodem_b200 as hc
    z = _dem(seed)
    assert np.array_equal(hc.fill_single_cell_pits(z, N), oracle.fill_single_cell_pits(z, N))
    assert np.array_equal(hc.breach_single_cell_pits(z, N), oracle.breach_single_cell_pits(z, N))
    noise = np.random.default_rng(seed).random((64, 70)).astype(np.float32)  # dense pits: many bridge conflicts
    assert np.array_equal(hc.breach_single_cell_pits(noise, N), oracle.breach_single_cell_pits(noise, N))
    assert np.array_equal(hc.fill_single_cell_pits(noise, N), oracle.fill_single_cell_pits(noise, N))


def test_taudem_check_mask(oracle, cuda):
    import pydrodem_b200 as hc
    z = _dem(5)
    f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=1, table=1, flats=False)
    rng = np.random.default_rng(0)
    flat = rng.integers(0, 8, z.shape).astype(np.uint8)
    edm = rng.integers(0, 8, z.shape).astype(np.uint8)
    want = (d == 255) & (f != np.float32(N)) & (flat > 1) & (flat < 6) & (edm < 4)  # models/taudem_check.py:143-147
    got, n = hc.taudem_check(d, f, flat, edm, N)
    assert np.array_equal(got.astype(bool), want) and n == int(want.sum()) and n > 0
    fixed = hc.taudem_fix(f, flat, N, 0.10)
    want_fix = np.where((f != np.float32(N)) & (flat > 0), f + np.float32(0.10), f)
    assert np.array_equal(fixed, want_fix.astype(np.float32))


def test_wbt_duck_type_fill_and_d8(tmp_path, oracle, cuda):
    from pydrodem_b200 import raster_io
    from pydrodem_b200.wbt import WhiteboxToolsGPU
    z = _dem(7)
    geo = {33550: (12, [30.0, 30.0, 0.0]), 33922: (12, [0.0, 0.0, 0.0, 500000.0, 4100000.0, 0.0])}
    dem_tif = str(tmp_path / "inDEM.tif")
    raster_io.write_raster(dem_tif, z, {"geo": geo}, nodata=N)
    wbt = WhiteboxToolsGPU()
    wbt.set_verbose_mode(False)
    wbt.set_whitebox_dir("/nonexistent")
    fill_tif = dem_tif.replace(".tif", "_WL.tif")
    assert wbt.fill_depressions_wang_and_liu(dem_tif, fill_tif, fix_flats=False, flat_increment=0.0) == 0
    got, meta = raster_io.read_raster(fill_tif)
    assert meta.nodata == N and meta["geo"][33550][1][:2] == [30.0, 30.0]
    assert np.array_equal(got, oracle.fill(z, N, oracle.SEED_WBT))
    sf_tif = dem_tif.replace(".tif", "_SF.tif")
    wbt.fill_depressions(dem_tif, sf_tif, max_depth=None, fix_flats=False, flat_increment=0.0)
    assert np.array_equal(raster_io.read_raster(sf_tif)[0], got)
    eps_tif = dem_tif.replace(".tif", "_WLeps.tif")
    assert wbt.fill_depressions_wang_and_liu(dem_tif, eps_tif, fix_flats=True, flat_increment=1e-5) == 0
    eps_fill = raster_io.read_raster(eps_tif)[0]
    assert eps_fill.dtype == np.float64 and np.array_equal(eps_fill[z != N], oracle.fill_eps(z, 1e-5, N, oracle.SEED_WBT)[z != N])
    with pytest.raises(NotImplementedError):
        wbt.fill_depressions_wang_and_liu(dem_tif, fill_tif, fix_flats=True, flat_increment=None)
    d8_tif, acc_tif = str(tmp_path / "d8.tif"), str(tmp_path / "acc.tif")
    wbt.d8_pointer(fill_tif, d8_tif)
    wbt.d8_flow_accumulation(d8_tif, acc_tif, pntr=True)
    p = raster_io.read_raster(d8_tif)[0]
    d_want = oracle.d8(got, N, 30.0, 30.0, oracle.TABLE_WBT, want_slope=False)
    assert p.dtype == np.int16 and np.array_equal(np.where(p == -32768, 255, p).astype(np.uint8), d_want)
    acc = raster_io.read_raster(acc_tif)[0]
    a_want = oracle.accum(d_want, oracle.TABLE_WBT)
    assert np.array_equal(np.where(acc < 0, 0, acc).astype(np.uint32), a_want)
    scp = str(tmp_path / "scp.tif")
    wbt.fill_single_cell_pits(dem_tif, scp)
    assert np.array_equal(raster_io.read_raster(scp)[0], oracle.fill_single_cell_pits(z, N))


def test_tools_depressions_mirror(tmp_path, oracle, cuda):
    """dep.fill_pits / dep.create_clump_array with the reference's signatures and file naming."""
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools import depressions as dep
    from pydrodem_b200.wbt import WhiteboxToolsGPU
    z = _dem(9)
    dem_tif = str(tmp_path / "inDEM.tif")
    raster_io.write_raster(dem_tif, z, nodata=N)
    wbt = WhiteboxToolsGPU()
    fill_array, diff = dep.fill_pits(dem_tif, wbt, fill_method="WL", nodataval=-9999, fixflats=False, flat_increment=0.0)
    want = oracle.fill(z, N, oracle.SEED_WBT)
    assert np.array_equal(fill_array, want)
    assert np.array_equal(diff, np.round(want - z, 3))  # tools/depressions.py:1308
    assert os.path.exists(dem_tif.replace(".tif", "_WL.tif")) and os.path.exists(dem_tif.replace(".tif", "_WL_DIFF.tif"))
    ids = dep.create_clump_array(diff, str(tmp_path / "pitID.tif"), dem_tif, wbt, condition="greaterthan")
    lab, n = oracle.label(diff > 0, 8)
    assert np.array_equal(ids, lab)
    ids_b = dep.create_clump_array(diff, str(tmp_path / "pitIDb.tif"), dem_tif, wbt, condition="greaterthan", clumpbuffer=1)
    from scipy import ndimage
    dil = ndimage.binary_dilation(diff > 0, structure=np.ones((3, 3)))
    assert np.array_equal(ids_b, oracle.label(dil, 4)[0])  # buffered clumps use 4-connectivity (:1646-1650)
    dneg = np.where(z == np.float32(N), -1.0, 0.0).astype(np.float32)
    ids_c = dep.create_clump_array(dneg, str(tmp_path / "c.tif"), dem_tif, wbt, condition="lessthan", save_tif=False)
    assert np.array_equal(ids_c, oracle.label(dneg < 0, 8)[0])


def test_taudem_cli(tmp_path, oracle, cuda):
    """PitRemove / D8FlowDir / AreaD8 with TauDEM's own flag names (run_aws.py:580-617)."""
    from pydrodem_b200 import raster_io
    z = _dem(11)
    dem = str(tmp_path / "dem.tif")
    raster_io.write_raster(dem, z, {"geo": {33550: (12, [30.0, 30.0, 0.0])}}, nodata=N)
    fel, p, sd8, ad8 = (str(tmp_path / n) for n in ("fel.tif", "p.tif", "sd8.tif", "ad8.tif"))
    run = lambda *a: subprocess.run([sys.executable, "-m", "pydrodem_b200.taudem", *a], check=True, cwd=ROOT)  # noqa: E731
    run("PitRemove", "-z", dem, "-fel", fel)
    run("D8FlowDir", "-fel", fel, "-p", p, "-sd8", sd8)
    run("AreaD8", "-nc", "-p", p, "-ad8", ad8)
    f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=1, table=1)
    assert np.array_equal(raster_io.read_raster(fel)[0], f)
    pp = raster_io.read_raster(p)[0]
    assert pp.dtype == np.int16 and np.array_equal(np.where(pp == -32768, 255, pp).astype(np.uint8), d)
    assert np.array_equal(raster_io.read_raster(sd8)[0], s)
    aa = raster_io.read_raster(ad8)[0]
    assert np.array_equal(np.where(aa < 0, 0, aa).astype(np.uint32), a)


def test_diff_round_equals_numpy(cuda):
    """hc_diff_round_dev = np.round(a - b, 3) on float32, bit for bit (ties, negatives, nodata-sized values, NaN)"""
    import pydrodem_b200 as hc
    rng = np.random.default_rng(5)
    a = (rng.normal(0, 30, (257, 301))).astype(np.float32)
    b = (a + rng.normal(0, 2, a.shape)).astype(np.float32)
    b.ravel()[::7] = a.ravel()[::7]                       # zeros
    a.ravel()[3::11] = b.ravel()[3::11] + np.float32(0.0005)   # ties at the third decimal
    a[0, 0], b[0, 0] = -9999.0, -9999.0
    a[0, 1], b[0, 1] = np.nan, 1.0
    a[0, 2], b[0, 2] = 3.0e6, -9999.0
    want = np.round((a - b), 3)
    got = hc.core.diff_round(a, b, 3)
    assert got.dtype == np.float32
    assert np.array_equal(got, want, equal_nan=True)
    assert np.array_equal(np.signbit(got), np.signbit(want))
