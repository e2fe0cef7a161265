"""Least-cost breach on the GPU (csrc/leastcost.cu, one CTA per pit) against its CPU specification
oracle/breach_leastcost.py, bit for bit: both orderings (cheapest / shortest), cost limit, channel gradient, nodata
holes, windows larger than the raster; then through the WhiteboxTools-shaped boundary as dep.breach_pits(method='LC')
calls it (tools/depressions.py:1370-1374, :1384-1387)."""
import os

import numpy as np
import pytest

from oracle import breach_leastcost as lc
from oracle import hydro_oracle as ho
from pydrodem_b200 import synth

ND = -9999.0


def scene(ny, nx, seed, kind):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        z = rng.integers(0, 60, (ny, nx)).astype(np.float32) * np.float32(0.25)
    elif kind == "terrain":
        z = (synth.make_dem_numpy(ny, nx, seed, n_bowls=4) * 0.05).astype(np.float32)
        z += rng.integers(0, 3, z.shape).astype(np.float32) * np.float32(0.5)
    else:   # steps: wide flats, many ties
        z = np.floor(synth.make_dem_numpy(ny, nx, seed, n_bowls=3) * 0.01).astype(np.float32)
    z[ny // 3:ny // 3 + 3, nx // 2:nx // 2 + 5] = ND
    z[:, -2:] = ND
    return z


CASES = [
    (40, 50, 1, "noise", dict(dist=6)),
    (40, 50, 2, "noise", dict(dist=6, min_dist=True)),
    (64, 70, 3, "terrain", dict(dist=12, max_cost=3.0)),
    (64, 70, 4, "terrain", dict(dist=12, flat_increment=0.01)),
    (50, 90, 5, "steps", dict(dist=9, min_dist=True, max_cost=10.0, flat_increment=0.001)),
    (33, 37, 6, "noise", dict(dist=40)),                       # the window is larger than the raster
    (70, 64, 7, "steps", dict(dist=5)),
]


@pytest.mark.gpu
@pytest.mark.parametrize("ny,nx,seed,kind,kw", CASES)
def test_device_least_cost_breach_equals_the_specification(ny, nx, seed, kind, kw):
    import pydrodem_b200 as hc
    z = scene(ny, nx, seed, kind)
    want, st_want = lc.breach_least_cost(z, ND, return_stats=True, **kw)
    got, st = hc.breach_depressions_least_cost(z, kw["dist"], nodata=ND, max_cost=kw.get("max_cost"),
                                               min_dist=kw.get("min_dist", False),
                                               flat_increment=kw.get("flat_increment"), fill=False, return_stats=True)
    assert st == st_want
    assert st["pits"] > 0 and st["breached"] > 0
    assert np.array_equal(got, want)
    # fill=True: the library's fill of the cut surface
    full = hc.breach_depressions_least_cost(z, kw["dist"], nodata=ND, max_cost=kw.get("max_cost"),
                                            min_dist=kw.get("min_dist", False), flat_increment=kw.get("flat_increment"))
    assert np.array_equal(full, ho.fill(want, ND, ho.SEED_WBT))


@pytest.mark.gpu
def test_dist_beyond_the_shared_memory_window_is_refused():
    import pydrodem_b200 as hc
    from pydrodem_b200._lib import HydrocoreError
    with pytest.raises(HydrocoreError):
        hc.breach_depressions_least_cost(scene(30, 30, 1, "noise"), 66, nodata=ND)


@pytest.mark.gpu
def test_breach_pits_method_lc_through_the_wbt_boundary(tmp_path):
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools import depressions as dep
    from pydrodem_b200.wbt import WhiteboxToolsGPU
    z = scene(60, 72, 11, "terrain")
    src = os.path.join(tmp_path, "dem.tif")
    raster_io.write_raster(src, z, None, nodata=ND)
    carve, diff = dep.breach_pits(src, 10, 4.0, WhiteboxToolsGPU(), min_dist=False, flat_increment=0.0,
                                  breach_increment=0.0001, nodataval=ND, method='LC')
    cut = lc.breach_least_cost(z, ND, dist=10, max_cost=4.0, min_dist=False, flat_increment=0.0001)
    want = ho.fill(cut, ND, ho.SEED_WBT)
    assert os.path.exists(src.replace(".tif", "_LC_10.tif")) and os.path.exists(src.replace(".tif", "_LC_10_filled.tif"))
    assert np.array_equal(np.asarray(carve, np.float32), want)
    assert np.array_equal(np.asarray(diff, np.float32), np.round(want - z, 3).astype(np.float32))
