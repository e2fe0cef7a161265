"""Water stencils against the reference's own functions (tests/golden/make_water_golden.py executes the unmodified
source of tools/water.py:786-1045).  min / max filters and selects move values, so river_minimize and
raise_small_rivers are bit-exact by construction.  raise_small_streams and smooth_rivers also use a 5x5 mean, which
is accumulated in float64 and narrowed to float32 like scipy's but can differ from a given scipy build by one
float32 ulp on rounding ties (see test_conv_f64...); on these scenes no such cell decides an output, so all
comparisons below are exact."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "water_reference.npz")
CASES = ("a", "b", "c")   # 150x181; 23x29 (smaller than the 51-cell window: multiple reflections); 64x40


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _scene(gold, n):
    return (gold[f"{n}_dem"], gold[f"{n}_burn"], gold[f"{n}_wm"], gold[f"{n}_swo"])


def test_golden_scenes_exercise_every_function(gold):
    for n in CASES:
        burn = gold[f"{n}_burn"]
        for k in ("rm_default", "rm_call", "rss_default", "sr_wide"):
            assert np.count_nonzero(gold[f"{n}_{k}"] != burn) > 0, (n, k)
    assert np.count_nonzero(gold["a_rsr_call"] != gold["a_burn"]) > 50
    assert np.count_nonzero(gold["a_sr_default"] != gold["a_burn"]) > 50


@pytest.mark.gpu
@pytest.mark.parametrize("n", CASES)
def test_minmax_filters_match_scipy(gold, n):
    from pydrodem_b200.tools import water
    dem, burn, wm, swo = _scene(gold, n)
    assert np.array_equal(water.minimum_filter(burn, 51), gold[f"{n}_min51"])
    assert np.array_equal(water.maximum_filter(wm, 35), gold[f"{n}_max35_u8"])
    assert np.array_equal(water.minimum_filter(burn, 4), gold[f"{n}_min4"])      # even size: scipy's origin rule
    from scipy import ndimage
    for size in (1, 3, 5, 15, 21, 31):
        assert np.array_equal(water.maximum_filter(burn, size), ndimage.maximum_filter(burn, size=size))
        assert np.array_equal(water.minimum_filter(wm, size), ndimage.minimum_filter(wm, size=size))


@pytest.mark.gpu
@pytest.mark.parametrize("n", CASES)
def test_river_minimize(gold, n):
    from pydrodem_b200.tools import water
    dem, burn, wm, swo = _scene(gold, n)
    assert np.array_equal(water.river_minimize(burn.copy(), wm, swo), gold[f"{n}_rm_default"])
    got = water.river_minimize(burn.copy(), wm, swo, max_burn=2, river_filter_size=5, nodataval=-9999)
    assert got.dtype == np.float32 and np.array_equal(got, gold[f"{n}_rm_call"])
    assert np.array_equal(water.river_minimize(burn, np.zeros_like(wm), swo), burn)   # no river: identity


@pytest.mark.gpu
@pytest.mark.parametrize("n", CASES)
def test_raise_small_rivers(gold, n):
    from pydrodem_b200.tools import water
    dem, burn, wm, swo = _scene(gold, n)
    b = burn.copy()
    r = water.raise_small_rivers(b, b, wm, swo, -9999, SWO_cutoff_low=5, SWO_cutoff_high=50, filter_size=31,
                                 extra_raise=0.10, max_raise=2.0)
    assert r is b and np.array_equal(b, gold[f"{n}_rsr_call"])                   # in place, like the reference
    assert np.array_equal(water.raise_small_rivers(dem, burn.copy(), wm, swo, -9999), gold[f"{n}_rsr_default"])


@pytest.mark.gpu
@pytest.mark.parametrize("n", CASES)
def test_raise_small_streams(gold, n):
    from pydrodem_b200.tools import water
    dem, burn, wm, swo = _scene(gold, n)
    assert np.array_equal(water.raise_small_streams(dem, burn.copy(), wm, -9999, extra_raise=1.0, max_raise=2.0),
                          gold[f"{n}_rss_call"])
    assert np.array_equal(water.raise_small_streams(dem, burn.copy(), wm, -9999), gold[f"{n}_rss_default"])
    b = burn.copy()
    assert np.array_equal(water.raise_small_streams(b, b, wm, -9999, extra_raise=0.25, max_raise=1.0),
                          gold[f"{n}_rss_alias"])


@pytest.mark.gpu
@pytest.mark.parametrize("n", CASES)
def test_smooth_rivers(gold, n):
    from pydrodem_b200.tools import water
    dem, burn, wm, swo = _scene(gold, n)
    assert np.array_equal(water.smooth_rivers(dem, burn.copy(), wm, swo, -9999), gold[f"{n}_sr_default"])
    assert np.array_equal(water.smooth_rivers(dem, burn.copy(), wm, swo, -9999, SWO_cut=40, max_lower=3.0,
                                              filter_size=9), gold[f"{n}_sr_wide"])


@pytest.mark.gpu
def test_device_tensors_stay_on_device_and_update_in_place(gold):
    import torch
    from pydrodem_b200.tools import water
    dem, burn, wm, swo = _scene(gold, "a")
    d, b = torch.from_numpy(dem).cuda(), torch.from_numpy(burn.copy()).cuda()
    w, s = torch.from_numpy(wm).cuda(), torch.from_numpy(swo).cuda()
    r = water.raise_small_streams(d, b, w, -9999)
    assert r is b and r.is_cuda and np.array_equal(b.cpu().numpy(), gold["a_rss_default"])
    o = water.river_minimize(torch.from_numpy(burn).cuda(), w, s)
    assert o.is_cuda and np.array_equal(o.cpu().numpy(), gold["a_rm_default"])


@pytest.mark.gpu
def test_conv_f64_matches_reference_float64_path(gold):
    """apply_convolve_filter(float32 raster, float64 kernel): scipy convolves in float64 and narrows"""
    import ctypes
    import torch
    from scipy import signal
    from pydrodem_b200 import _lib
    burn = gold["a_burn"]
    k = (1. / 25.0) * np.ones((5, 5))
    want = signal.convolve2d(burn, k, boundary='symm', mode='same').astype(np.float32)
    a = torch.from_numpy(burn).cuda()
    out = torch.empty_like(a)
    _lib.check(_lib.load().hc_conv2d_symm_f64_dev(_lib.context(0), ctypes.c_void_p(a.data_ptr()),
                                                  ctypes.c_void_p(k.ctypes.data), 5, 5, ctypes.c_void_p(out.data_ptr()),
                                                  a.shape[0], a.shape[1], None), "hc_conv2d_symm_f64_dev")
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    # The float64 sums agree with scipy's to ~1e-16; narrowing to float32 hides that except where the exact mean of
    # 25 float32 values falls on a float32 rounding tie, which is not rare (0.27 % of this raster) -- there the
    # summation order of the scipy build decides, so those cells may differ by one float32 ulp.
    assert np.count_nonzero(got != want) <= 0.005 * got.size
    assert np.allclose(got, want, rtol=1.2e-7, atol=0)
    rng = np.random.default_rng(3)
    k2 = rng.random((3, 7))                              # asymmetric kernel: checks the flip of a true convolution
    want = signal.convolve2d(burn, k2, boundary='symm', mode='same').astype(np.float32)
    _lib.check(_lib.load().hc_conv2d_symm_f64_dev(_lib.context(0), ctypes.c_void_p(a.data_ptr()),
                                                  ctypes.c_void_p(k2.ctypes.data), 3, 7, ctypes.c_void_p(out.data_ptr()),
                                                  a.shape[0], a.shape[1], None), "hc_conv2d_symm_f64_dev")
    torch.cuda.synchronize()
    assert np.allclose(out.cpu().numpy(), want, rtol=2.4e-7, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("nbands", [2, 3])
def test_water_stencils_over_row_bands(gold, nbands):
    """bands + halo rows: each band runs the whole-raster function on its rows plus
    water.halo_rows(...) rows of its neighbours; the stitched result is bit-identical"""
    from pydrodem_b200 import bands
    from pydrodem_b200.tools import water
    dem, burn, wm, swo = _scene(gold, "a")
    got = bands.apply_banded(lambda b, w, s: water.river_minimize(b, w, s, 2, 5), [burn, wm, swo], nbands,
                             water.halo_rows("river_minimize", river_filter_size=5))
    assert np.array_equal(got, gold["a_rm_call"])
    got = bands.apply_banded(lambda d, b, w: water.raise_small_streams(d, b, w, -9999), [dem, burn, wm], nbands,
                             water.halo_rows("raise_small_streams"))
    assert np.array_equal(got, gold["a_rss_default"])
    got = bands.apply_banded(lambda d, b, w, s: water.raise_small_rivers(d, b, w, s, -9999, filter_size=31),
                             [dem, burn, wm, swo], nbands, water.halo_rows("raise_small_rivers", filter_size=31))
    assert np.array_equal(got, water.raise_small_rivers(dem, burn.copy(), wm, swo, -9999, filter_size=31))
    got = bands.apply_banded(lambda d, b, w, s: water.smooth_rivers(d, b, w, s, -9999), [dem, burn, wm, swo], nbands,
                             water.halo_rows("smooth_rivers"))
    assert np.array_equal(got, gold["a_sr_default"])
