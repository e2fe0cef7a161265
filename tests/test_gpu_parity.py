"""GPU parity: libhydrocore (through the C ABI) vs the CPU oracle, bit for bit, on seeded inputs.

Bar (BASELINE.json north_star): filled float32 elevations, D8 codes and integer accumulation are
bit-exact; the float32 slope written beside D8 is bit-exact too (same arithmetic, no reductions).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # (ny, nx, kind, seed, holes, levels)
    (1, 1, "noise", 1, 0, None),
    (1, 7, "noise", 2, 0, None),
    (9, 1, "noise", 3, 0, None),
    (2, 2, "noise", 4, 0, None),
    (5, 5, "noise", 5, 0, None),
    (17, 33, "noise", 6, 0, 8),
    (64, 64, "noise", 7, 0, None),
    (65, 63, "noise", 8, 2, None),
    (130, 97, "noise", 9, 3, 16),
    (200, 300, "cone", 10, 2, 40),
    (200, 300, "bowl", 11, 2, 40),
    (257, 511, "smooth", 12, 3, None),
    (300, 200, "noise", 13, 0, 3),
    (640, 640, "smooth", 14, 4, None),
    (1000, 777, "bowl", 15, 0, 200),
]


def _dem(case):
    from pydrodem_b200 import synth
    ny, nx, kind, seed, holes, levels = case
    if min(ny, nx) < 8 and kind != "noise":
        kind = "noise"
    z = synth.random_dem(ny, nx, seed, kind, nodata_holes=holes if min(ny, nx) >= 8 else 0, levels=levels)
    return z


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}-{c[2]}-{c[3]}")
@pytest.mark.parametrize("seed_mode", [0, 1], ids=["wbt", "taudem"])
def test_fill_bit_exact(case, seed_mode, oracle, cuda):
    import pydrodem_b200 as hc
    z = _dem(case)
    want = oracle.fill(z, seed_mode=seed_mode)
    got = hc.fill(z, seed_mode=seed_mode)
    assert got.dtype == np.float32 and got.shape == z.shape
    assert np.array_equal(got, want), f"{(got != want).sum()} cells differ"


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}-{c[2]}-{c[3]}")
@pytest.mark.parametrize("table", [0, 1], ids=["wbt", "taudem"])
def test_d8_flats_accum_bit_exact(case, table, oracle, cuda):
    import pydrodem_b200 as hc
    z = _dem(case)
    f = oracle.fill(z, seed_mode=table)
    d_want, s_want = oracle.d8(f, dx=30.0, dy=20.0, table=table)
    d_got, s_got = hc.d8(f, dx=30.0, dy=20.0, table=table)
    assert np.array_equal(d_got, d_want), f"d8: {(d_got != d_want).sum()} cells differ"
    assert np.array_equal(s_got, s_want), "slope differs"
    r_want = oracle.resolve_flats(f, d_want, table=table)
    r_got = hc.resolve_flats(f, d_want, table=table)
    assert np.array_equal(r_got, r_want), f"flats: {(r_got != r_want).sum()} cells differ"
    a_want = oracle.accum(r_want, table=table)
    a_got = hc.accum(r_want, table=table)
    assert np.array_equal(a_got, a_want), f"accum: {(a_got != a_want).sum()} cells differ"
    # accumulation of the unresolved pointer too (WBT's own chain has no flat resolution)
    assert np.array_equal(hc.accum(d_want, table=table), oracle.accum(d_want, table=table))


@pytest.mark.parametrize("case", CASES[4:], ids=lambda c: f"{c[0]}x{c[1]}-{c[2]}-{c[3]}")
def test_chain_host_and_device(case, oracle, cuda):
    import torch
    import pydrodem_b200 as hc
    z = _dem(case)
    for seed_mode, table in ((1, 1), (0, 0)):
        f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table)
        gf, gd, gs, ga = hc.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table)
        assert np.array_equal(gf, f) and np.array_equal(gd, d) and np.array_equal(gs, s) and np.array_equal(ga, a)
        zt = torch.from_numpy(z).to(cuda)
        tf, td, ts, ta = hc.fill_d8_accum(zt, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table)
        torch.cuda.synchronize()
        assert np.array_equal(tf.cpu().numpy(), f)
        assert np.array_equal(td.cpu().numpy(), d)
        assert np.array_equal(ta.cpu().numpy().view(np.uint32), a)


@pytest.mark.parametrize("conn", [4, 8])
@pytest.mark.parametrize("shape,p,seed", [((1, 1), 0.5, 1), ((7, 9), 0.5, 2), ((64, 64), 0.6, 3), ((130, 257), 0.45, 4),
                                          ((500, 333), 0.55, 5), ((1025, 1023), 0.59, 6), ((300, 300), 1.0, 7),
                                          ((300, 300), 0.0, 8)])
def test_label_matches_scipy(shape, p, seed, conn, oracle, cuda):
    import pydrodem_b200 as hc
    rng = np.random.default_rng(seed)
    m = (rng.random(shape) < p).astype(np.uint8)
    want, n_want = oracle.label(m, conn)
    got, n_got = hc.label(m, conn)
    assert n_got == n_want
    assert np.array_equal(got, want)


def test_medium_synthetic_tile(oracle, cuda):
    """A 1201^2 crop-sized version of configuration C2 (pits + bowls, 1 cm quantised)."""
    import pydrodem_b200 as hc
    from pydrodem_b200 import synth
    z = synth.make_dem_numpy(1201, 1201, 1002, nodata_frame=True)
    for seed_mode, table in ((1, 1), (0, 0)):
        f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table)
        gf, gd, gs, ga = hc.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table)
        assert np.array_equal(gf, f)
        assert np.array_equal(gd, d)
        assert np.array_equal(gs, s)
        assert np.array_equal(ga, a)


def test_c4_flattened_lakes_and_rivers(oracle, cuda):
    """configs[3] recipe at 900 x 1100: lakes and river reaches set to exact constants (flat resolution heavy:
    flats hundreds of cells across go through the slow tier), whole raster and 3 bands, vs the oracle."""
    import pydrodem_b200 as hc
    from pydrodem_b200 import bands, synth
    z, mask = synth.make_c4(900, 1100, 1004)
    z = z.numpy()
    assert (mask.numpy() == 3).sum() > 500 and (mask.numpy() == 1).sum() > 2000
    f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=1, table=1)
    gf, gd, gs, ga = hc.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=1, table=1)
    assert np.array_equal(gf, f) and np.array_equal(gd, d) and np.array_equal(gs, s) and np.array_equal(ga, a)
    bf, bd, bs, ba = bands.fill_d8_accum_banded(z, 3, dx=30.0, dy=30.0, table=1)
    assert np.array_equal(bf, f) and np.array_equal(bd, d) and np.array_equal(ba, a)


@pytest.mark.parametrize("seed,levels", [(41, 30), (42, 80), (43, 12)])
def test_flats_spanning_many_tiles(seed, levels, oracle, cuda):
    """Terraced terrain (elevation quantised to a few dozen levels) gives flats that wind across many 64x64 tiles,
    with their outlets in tiles that hold no pending cell themselves: the slow tier's lazy start (only tiles that
    hold part of an open flat, woken transitively) must still reach every one of them."""
    import pydrodem_b200 as hc
    from pydrodem_b200 import synth
    z = synth.random_dem(1100, 1300, seed, "smooth", nodata_holes=2, levels=levels)
    f = oracle.fill(z, seed_mode=1)
    d0, _ = oracle.d8(f, dx=30.0, dy=30.0, table=1)
    want = oracle.resolve_flats(f, d0, table=1)
    got = hc.resolve_flats(f, d0, table=1)
    assert np.array_equal(got, want), f"{(got != want).sum()} cells differ"
    assert np.array_equal(hc.accum(got, table=1), oracle.accum(want, table=1))


@pytest.mark.parametrize("case", [CASES[7], CASES[9], CASES[11], CASES[14]], ids=lambda c: f"{c[0]}x{c[1]}-{c[2]}-{c[3]}")
@pytest.mark.parametrize("seed_mode,table", [(1, 1), (0, 0)], ids=["taudem", "wbt"])
def test_chain_pitched_rasters_equal_dense(case, seed_mode, table, oracle, cuda):
    """hc_fill_d8_accum_ld_dev (pitched rasters: rows padded to 64 elements, padding = nodata, TMA-staged tiles) must
    give the dense call's / the oracle's result in the real columns, for widths that are not multiples of 4."""
    import torch
    import pydrodem_b200 as hc
    z = _dem(case)
    f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table)
    zp = hc.empty_raster(z.shape[0], z.shape[1], torch.float32, cuda)
    zp.copy_(torch.from_numpy(z))
    gf, gd, gs, ga = hc.fill_d8_accum(zp, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table)
    assert gf.stride(0) % 64 == 0 and gf.shape == z.shape
    assert np.array_equal(gf.cpu().numpy(), f)
    assert np.array_equal(gd.cpu().numpy(), d)
    assert np.array_equal(gs.cpu().numpy(), s)
    assert np.array_equal(ga.cpu().numpy().view(np.uint32), a)


@pytest.mark.parametrize("case", [CASES[7], CASES[9], CASES[10], CASES[11]], ids=lambda c: f"{c[0]}x{c[1]}-{c[2]}-{c[3]}")
@pytest.mark.parametrize("seed_mode", [0, 1], ids=["wbt", "taudem"])
def test_fill_eps_equals_priority_flood_plus_epsilon(case, seed_mode, oracle, cuda):
    """flat_increment > 0: hc_fill_eps_dev against the oracle's serial Priority-Flood + epsilon (float64).  The surface is
    the unique solution of W = max(z, min_n W(n) + eps), and both make one float64 addition per step of a chain, so the
    comparison is exact here; the stated tolerance towards WhiteboxTools itself (unknown heap order, parity unpinned) is
    eps times the longest chain."""
    import pydrodem_b200 as hc
    z = _dem(case)
    for eps in (1e-5, 1e-3):
        want = oracle.fill_eps(z, eps, seed_mode=seed_mode)
        got, sweeps = hc.fill_eps(z, eps, seed_mode=seed_mode, return_sweeps=True)
        assert got.dtype == np.float64 and got.shape == z.shape and sweeps >= 8
        valid = z != np.float32(-9999.0)
        assert np.array_equal(got[valid], want[valid]), f"eps={eps}: max |diff| {np.abs(got[valid] - want[valid]).max()}"
        assert np.all(got[valid] >= hc.fill(z, seed_mode=seed_mode)[valid])          # never below the eps = 0 surface
    assert np.array_equal(hc.fill_eps(z, 0.0, seed_mode=seed_mode)[valid], hc.fill(z, seed_mode=seed_mode)[valid].astype(np.float64))


@pytest.mark.gpu
def test_wbt_fill_with_flat_increment_writes_float64_and_fill_pits_downcasts(tmp_path, oracle, cuda):
    import pydrodem_b200 as hc
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools import depressions as dep
    from pydrodem_b200.wbt import WhiteboxToolsGPU
    z = _dem(CASES[11])
    dem = str(tmp_path / "u.tif")
    raster_io.write_raster(dem, z, None, nodata=-9999.0)
    wbt = WhiteboxToolsGPU()
    fill32, diff = dep.fill_pits(dem, wbt, flat_increment=1e-4)
    assert fill32.dtype == np.float32
    want = oracle.fill_eps(z, 1e-4, seed_mode=0).astype(np.float32)
    valid = z != np.float32(-9999.0)
    assert np.array_equal(fill32[valid], want[valid])
    on_disk, _ = raster_io.read_raster(str(tmp_path / "u_WL.tif"))
    assert on_disk.dtype == np.float32
    fill64, _ = dep.fill_pits(dem, wbt, flat_increment=1e-4, downcast=False)
    assert fill64.dtype == np.float64
    with pytest.raises(NotImplementedError):
        wbt.fill_depressions_wang_and_liu(dem, str(tmp_path / "x.tif"), fix_flats=True, flat_increment=None)
