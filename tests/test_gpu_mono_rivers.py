"""wtr.enforce_monotonic_rivers (tools/water.py:1407-1779) on the GPU path against the reference's own function
(tests/golden/make_mono_golden.py executes its source unmodified; the WhiteboxTools interpolation at the end is the
oracle's restatement, so the interpolated cells are oracle-pinned) -- and hc_fill_missing_data_dev against that oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "mono_rivers.npz"))


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_enforce_monotonic_rivers_equals_reference(gold, cuda, name, tmp_path):
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_mono_golden import scene
    from pydrodem_b200.tools import water
    dem, swo, wm, donut, inner = scene(int(gold[f"{name}_seed"]))
    d0, w0 = dem.copy(), wm.copy()
    d_out, w_out = water.enforce_monotonic_rivers(None, str(tmp_path), dem, swo, wm, donut, name, SWO_cutoff=20,
                                                  radius=int(gold[f"{name}_radius"]),
                                                  strict_endcap_geom=bool(gold[f"{name}_strict"]), nodata=-9999, wbt=None,
                                                  max_r_drop=50.0,
                                                  inner_boundary=inner if bool(gold[f"{name}_use_inner"]) else None)
    assert d_out is dem and w_out is wm                      # modified in place and returned, like the reference
    want_d, want_w = d0.copy(), w0.copy()
    rc = tuple(gold[f"{name}_changed_rc"])
    want_d[rc] = gold[f"{name}_dem_changed"]
    want_w[rc] = gold[f"{name}_wm_changed"]
    assert len(rc[0]) > 400 and set(np.unique(want_w)) >= {2, 6, 7}
    assert np.array_equal(w_out, want_w), f"{(w_out != want_w).sum()} mask cells differ"
    assert np.array_equal(d_out, want_d), f"{(d_out != want_d).sum()} cells differ, max {np.abs(d_out - want_d).max()}"


def test_enforce_monotonic_rivers_through_the_wbt_mirror(gold, cuda, tmp_path):
    """same result when the interpolation goes through the GeoTIFF protocol (wbt.fill_missing_data on files)"""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    from make_mono_golden import scene
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools import water
    from pydrodem_b200.wbt import WhiteboxToolsGPU
    dem, swo, wm, donut, inner = scene(int(gold["a_seed"]))
    ref_tif = str(tmp_path / "dem.tif")
    raster_io.write_raster(ref_tif, dem, None, nodata=-9999)
    a = water.enforce_monotonic_rivers(ref_tif, str(tmp_path), dem.copy(), swo, wm.copy(), donut, "t", radius=50,
                                       max_r_drop=50.0, inner_boundary=inner, wbt=WhiteboxToolsGPU())
    b = water.enforce_monotonic_rivers(None, str(tmp_path), dem.copy(), swo, wm.copy(), donut, "t", radius=50,
                                       max_r_drop=50.0, inner_boundary=inner, wbt=None)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert os.path.exists(str(tmp_path / "t_non_mono_interp.tif"))


@pytest.mark.parametrize("filt,weight", [(11, 2.0), (17, 2.0), (7, 1.5)])
def test_fill_missing_data_equals_oracle(cuda, filt, weight):
    from oracle.fill_missing_oracle import fill_missing_data as want_fn
    from pydrodem_b200.tools import water
    rng = np.random.default_rng(filt)
    z = (rng.normal(100, 5, (90, 120))).astype(np.float32)
    z[rng.random(z.shape) < 0.03] = -9999.0
    z[20:45, 30:70] = -9999.0            # a hole wider than the filter: its middle stays nodata
    z[0:3, :] = -9999.0
    z[60, 60] = np.nan
    want = want_fn(z, -9999.0, filt, weight)
    got = water.fill_missing_data(z, -9999.0, filt, weight)
    keep = ~np.isnan(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert (want[30:35, 45:55] == -9999.0).all()
    if weight == 2.0:
        assert np.array_equal(got[keep], want[keep])
    else:
        assert np.allclose(got[keep], want[keep], rtol=1e-6, atol=0)
