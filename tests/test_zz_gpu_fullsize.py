"""BASELINE.json's full size (C2: 10801 x 10801), checked two ways: bit for bit against the single-thread oracle (about a minute), and the chain is
checked through size-independent properties (tests/_properties.py, themselves checked against the oracle at small
size in the CPU suite) and through two independent code paths that must agree bit for bit -- the whole-raster chain
and the row-band chain (3 bands with seam exchange).  Runs last (file name) because it is the longest GPU test."""
import pytest

from tests import _properties as P


@pytest.mark.gpu
def test_c2_full_size_properties_and_band_agreement():
    import torch
    import pydrodem_b200 as hc
    from pydrodem_b200 import bands, synth
    ny = nx = 10801
    dev = torch.device("cuda", 0)
    z = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    for r0 in range(0, ny, 1024):                       # the bench's C2 recipe, generated on the device in row chunks
        nr = min(1024, ny - r0)
        z[r0:r0 + nr] = synth.make_dem(ny, nx, 1002, row0=r0, nrows=nr, device=dev)
    nodata = -9999.0
    f, d, s, a = hc.fill_d8_accum(z, nodata=nodata, dx=30.0, dy=30.0, seed_mode=hc.SEED_TAUDEM, table=hc.TABLE_TAUDEM,
                                  flats=True, want_slope=False)
    torch.cuda.synchronize()
    assert f.shape == z.shape and d.dtype == torch.uint8
    # fill: above the input, seeds fixed, no pit left; idempotent
    assert P.fill_violations(z, f, nodata) == 0
    assert int((f > z).sum()) > 100000                                  # the recipe has pits: the fill did something
    assert torch.equal(hc.fill(f, nodata=nodata, seed_mode=hc.SEED_TAUDEM), f)
    # directions: never uphill on the filled surface; the frame is "contaminated" (255) and nothing else is (the recipe
    # has no nodata); after flat resolution the only pointer-less cells are flats whose way out is a frame cell on
    # their own level -- a small minority
    assert P.uphill_violations(f, d, hc.TABLE_TAUDEM) == 0
    assert int((d[1:-1, 1:-1] == 255).sum()) == 0 and bool((d[0] == 255).all()) and bool((d[:, -1] == 255).all())
    assert int((d[1:-1, 1:-1] == 0).sum()) < 1e-3 * ny * nx
    # accumulation: the exact in-flow identity on every cell, and all mass arrives at the path ends
    assert P.accumulation_violations(d, a, hc.TABLE_TAUDEM) == 0
    down = P.downstream(d, hc.TABLE_TAUDEM).reshape(-1)
    ends = (down < 0) & (d.reshape(-1) != 255)
    assert int(a.reshape(-1)[ends].to(torch.int64).sum()) == int((d != 255).sum())
    del down, ends
    # an independent path: the same DEM as three row bands with seam-graph exchange must give the same rasters
    bf, bd, bs, ba = bands.fill_d8_accum_banded(z, 3, nodata=nodata, dx=30.0, dy=30.0, table=hc.TABLE_TAUDEM, flats=True,
                                                want_slope=False, seed_mode=hc.SEED_TAUDEM)
    assert torch.equal(bf, f) and torch.equal(bd, d) and torch.equal(ba, a)


def _device_dem(ny, nx, seed, dev, **kw):
    import torch
    from pydrodem_b200 import synth
    z = torch.empty((ny, nx), dtype=torch.float32, device=dev)
    for r0 in range(0, ny, 1024):
        nr = min(1024, ny - r0)
        z[r0:r0 + nr] = synth.make_dem(ny, nx, seed, row0=r0, nrows=nr, device=dev, **kw)
    return z


def _assert_chain_equals_oracle(z, oracle, seed_mode, table, nodata=-9999.0):
    """CUDA chain vs the CPU oracle, bit for bit, on the same raster (the oracle runs a few seconds per 13 Mcells)."""
    import numpy as np
    import torch
    import pydrodem_b200 as hc
    f, d, s, a = hc.fill_d8_accum(z, nodata=nodata, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table, flats=True,
                                  want_slope=False)
    torch.cuda.synchronize()
    zh = z.cpu().numpy()
    of, od, os_, oa = oracle.fill_d8_accum(zh, nodata=nodata, dx=30.0, dy=30.0, seed_mode=seed_mode, table=table)
    gf, gd, ga = f.cpu().numpy(), d.cpu().numpy(), a.cpu().numpy().view(np.uint32)
    assert np.array_equal(gf, of), f"fill: {(gf != of).sum()} cells differ"
    assert np.array_equal(gd, od), f"D8 + flats: {(gd != od).sum()} cells differ"
    assert np.array_equal(ga, oa), f"accumulation: {(ga != oa).sum()} cells differ"
    return int((gf > zh).sum())


@pytest.mark.gpu
@pytest.mark.parametrize("seed_mode,table", [(1, 1), (0, 0)], ids=["taudem", "wbt"])
def test_c1_size_bit_exact_vs_oracle(seed_mode, table, oracle, cuda):
    """BASELINE.json configs[0] size (3601^2, seed 1001, nodata frame on two sides): both seed rules / code tables."""
    z = _device_dem(3601, 3601, 1001, cuda, nodata_frame=True)
    assert _assert_chain_equals_oracle(z, oracle, seed_mode, table) > 10000


@pytest.mark.gpu
def test_c2_full_size_bit_exact_vs_oracle(oracle, cuda):
    """BASELINE.json configs[1] at full size with bench.py's exact recipe and seed (10801^2, seed 1002, TauDEM semantics):
    the whole chain against the single-thread oracle, bit for bit (about a minute of oracle time)."""
    z = _device_dem(10801, 10801, 1002, cuda)
    assert _assert_chain_equals_oracle(z, oracle, 1, 1) > 100000


@pytest.mark.gpu
def test_c4_recipe_3601_bit_exact_vs_oracle(oracle, cuda):
    """configs[3] recipe (flattened lakes, burned rivers: large exact flats) at 3601^2 against the oracle."""
    from pydrodem_b200 import synth
    z, water = synth.make_c4(3601, 3601, 1004, device=cuda)
    assert int((water > 0).sum()) > 1000
    _assert_chain_equals_oracle(z, oracle, 1, 1)
