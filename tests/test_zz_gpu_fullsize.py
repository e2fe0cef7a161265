"""BASELINE.json's full size (C2: 10801 x 10801), where the single-thread oracle cannot follow in test time: the chain is
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
