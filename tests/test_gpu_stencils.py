"""GPU stencils vs golden vectors produced by the reference's own tools/derive.py
(tests/golden/make_stencil_golden.py).  Tolerances are stated per check."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stencils_reference.npz")


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(G))


def test_build_kernel_matches_reference(gold):
    from pydrodem_b200 import stencils
    for n in (9, 7, 5, 3):
        k = stencils.build_kernel(stencils.DW_WEIGHTS[n], dtype=np.float32, normalize=True)
        assert np.array_equal(k, gold[f"kernel{n}"])


@pytest.mark.parametrize("n", [9, 7, 5, 3])
def test_dw_convolution(gold, cuda, n):
    """float32 accumulation in a different order than scipy's: |err| <= 2e-6 * max|z| (about 2 ulp)."""
    from pydrodem_b200 import stencils
    got = stencils.apply_convolve_filter(gold["dem"], gold[f"kernel{n}"])
    want = gold[f"z{n}"]
    assert got.dtype == np.float32 and got.shape == want.shape
    tol = 2e-6 * float(np.abs(want).max())
    assert float(np.abs(got - want).max()) <= tol


def test_box_dilation_is_exact(gold, cuda):
    from pydrodem_b200 import stencils
    got = stencils.apply_convolve_filter(gold["dilate_in"], np.ones((3, 3)))
    assert np.array_equal(got > 0, gold["dilate_out"] > 0)
    assert np.array_equal(got, gold["dilate_out"])  # small integers: exact in float32


def test_gradient_magnitudes_bit_exact(gold, cuda):
    """np.gradient / sqrt / the float64 diagonal kernels are reproduced operation by operation."""
    from pydrodem_b200 import stencils
    z5 = gold["z5"]
    dx, dy = float(gold["dx"]), float(gold["dy"])
    assert np.array_equal(stencils.slope_D2(z5, dx, dy, unit="none"), gold["slope_d2_mag"])
    assert np.array_equal(stencils.slope_D4(z5, dx, dy, unit="none"), gold["slope_d4_mag"])


def test_slope_degrees(gold, cuda):
    """int8 degrees go through atanf: at most 1 degree off, and only at truncation boundaries (< 0.1 % of cells)."""
    from pydrodem_b200 import stencils
    z5 = gold["z5"]
    dx, dy = float(gold["dx"]), float(gold["dy"])
    for fn, key in ((stencils.slope_D2, "slope_d2_deg"), (stencils.slope_D4, "slope_d4_deg")):
        got = fn(z5, dx, dy)
        assert got.dtype == np.int8
        diff = np.abs(got.astype(int) - gold[key].astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3


def test_curvature_bit_exact(gold, cuda):
    from pydrodem_b200 import stencils
    dx, dy = float(gold["dx"]), float(gold["dy"])
    assert np.array_equal(stencils.curvature_D2(gold["z5"], dx, dy), gold["curv_geo"])
    assert np.array_equal(stencils.curvature_D2(gold["z5"], dx, dy, geometric=False), gold["curv_lap"])


def test_sigma(gold, cuda):
    """float64 two-pass std vs numpy's float32 pairwise std: relative 1e-5."""
    from pydrodem_b200 import stencils
    m, s = stencils.mean_std(gold["curv_geo"])
    assert abs(s - float(gold["sigma"])) <= 1e-5 * float(gold["sigma"])
    assert abs(m - float(np.mean(gold["curv_geo"].astype(np.float64)))) <= 1e-9 + 1e-6 * abs(m)


@pytest.mark.parametrize("variant", ["plain", "masked"])
def test_filter_selection_exact_given_reference_inputs(gold, cuda, variant):
    from pydrodem_b200 import stencils
    masks = {} if variant == "plain" else dict(swo=gold["swo"], edm=gold["edm"], tx=gold["tx"], artifact=gold["artifact"])
    zf, filt = stencils.dtm_select(gold["dem"], gold["z9"], gold["z7"], gold["z5"], gold["z3"], gold["slope_d4_deg"],
                                   gold["curv_geo"], gold["sigma"], **masks)
    assert np.array_equal(filt, gold[f"filter_{variant}"])
    assert np.array_equal(zf, gold[f"zf_{variant}"])


def test_smooth_core_end_to_end(gold, cuda):
    """Whole core on the GPU (its own convolutions, sigma, slope): filter mask may differ where a
    smoothed value sits within float32 rounding of a slope/curvature break; elevations within 2e-6 rel."""
    from pydrodem_b200 import stencils
    zf, filt, slope = stencils.smooth_core(gold["dem"], float(gold["dx"]), float(gold["dy"]), gdir=4, swo=gold["swo"],
                                           edm=gold["edm"], tx=gold["tx"], artifact=gold["artifact"])
    same = filt == gold["filter_masked"]
    assert same.mean() > 0.999
    tol = 2e-6 * float(np.abs(gold["zf_masked"]).max())
    assert float(np.abs(zf - gold["zf_masked"])[same].max()) <= tol
    assert (np.abs(slope.astype(int) - gold["slope_final_masked"].astype(int))[same] <= 1).all()


@pytest.mark.parametrize("nbands", [2, 3, 5])
def test_smoothing_core_row_bands_match_whole_raster(gold, cuda, nbands):
    """the smoothing core shards into row bands with an 8-row halo; sigma is combined from per-band
    float64 sums.  Zf / filter / slope must equal the whole-raster result cell for cell."""
    from pydrodem_b200 import stencils
    from pydrodem_b200 import synth
    dem = synth.make_dem_numpy(230, 190, 1005, quantise=False)
    dem = (dem + np.random.default_rng(3).normal(0, 1.5, dem.shape)).astype(np.float32)
    zf, filt, slope = stencils.smooth_core(dem, 30.0, 30.0)
    bzf, bfilt, bslope = stencils.smooth_core_banded(dem, nbands, 30.0, 30.0)
    assert np.array_equal(bfilt, filt)
    assert np.array_equal(bzf, zf)
    assert np.array_equal(bslope, slope)


@pytest.mark.parametrize("shape", [(230, 190), (64, 64), (65, 129), (1, 40), (37, 1), (5, 5), (300, 517)])
@pytest.mark.parametrize("gdir", [4, 2])
def test_fused_core_equals_unfused_chain(cuda, shape, gdir):
    """hc_dtm_pre_dev + hc_dtm_apply_dev (two tile passes over the DEM, no smoothed raster in HBM) against the chain of
    single-purpose kernels (dw4 -> slope -> curvature -> sigma -> select): Zf, filter code and final slope bit for bit,
    masks included; sigma within float64 summation noise."""
    from pydrodem_b200 import stencils, synth
    ny, nx = shape
    rng = np.random.default_rng(ny * 1000 + nx)
    dem = synth.make_dem_numpy(max(ny, 8), max(nx, 8), 1005, quantise=False)[:ny, :nx]
    dem = (dem + rng.normal(0, 1.5, dem.shape)).astype(np.float32)
    swo = (rng.random(dem.shape) < 0.05).astype(np.uint8)
    edm = rng.integers(0, 7, dem.shape).astype(np.uint8)
    tx = (rng.random(dem.shape) < 0.03).astype(np.uint8)
    art = (rng.random(dem.shape) < 0.03).astype(np.uint8)
    for masks in ({}, dict(swo=swo, edm=edm, tx=tx, artifact=art)):
        a = stencils.smooth_core_unfused(dem, 30.0, 25.0, gdir=gdir, **masks)
        b = stencils.smooth_core(dem, 30.0, 25.0, gdir=gdir, **masks)
        for x, y, name in zip(a, b, ("zf", "filter", "slope")):
            assert np.array_equal(x, y), f"{name}: {(x != y).sum()} cells differ"
