"""Vegetation-bias lookup against the replayed reference (tests/golden/make_veg_golden.py)."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "veg_reference.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_load_bias_cube_matches_reference_replay(gold):
    from pydrodem_b200 import veg
    cube = veg.load_bias_cube(json.loads(str(gold["table_json"])))
    assert cube.dtype == np.float64
    assert np.array_equal(cube, gold["bias_cube"])


@pytest.mark.gpu
def test_box3_index_bit_exact(gold):
    from pydrodem_b200 import veg
    ocean = (gold["edm"] == 4).astype(np.uint8)
    idx, sm = veg.box3_index(gold["TC"], ocean, zero_above=100.0, return_smooth=True)
    assert np.array_equal(idx, gold["TC_idx"])
    assert np.array_equal(sm, gold["TC_smooth"])
    th = veg.box3_index(gold["Theight"], in_max=int(gold["max_treeheight"]))
    assert np.array_equal(th, gold["Th_idx"])


@pytest.mark.gpu
def test_lut_gather_bit_exact_on_reference_indices(gold):
    """the gather itself, fed the reference's own index arrays: float64 values copied, so bit-exact"""
    import torch
    from pydrodem_b200 import veg
    tc = torch.from_numpy(gold["TC_idx"]).cuda()
    th = torch.from_numpy(gold["Th_idx"]).cuda()
    s = torch.from_numpy(gold["slope_uncapped"].copy()).cuda()
    edm = gold["edm"]
    wm = torch.from_numpy(((edm == 2) | (edm == 3)).astype(np.uint8)).cuda()
    z = torch.from_numpy(gold["Zarr"]).cuda()
    bias, dem = veg.lut_gather(gold["bias_cube"], tc, th, s, veg.SLOPE_CAP_3D, wm, z)
    assert np.array_equal(s.cpu().numpy(), gold["slope_idx"])
    assert np.array_equal(bias.cpu().numpy(), gold["bias"])
    assert np.array_equal(dem.cpu().numpy(), gold["dem_out"])
    s = torch.from_numpy(gold["slope_uncapped"].copy()).cuda()
    raw, none = veg.lut_gather(gold["bias_cube"], tc, th, s, veg.SLOPE_CAP_3D)
    assert none is None and np.array_equal(raw.cpu().numpy(), gold["bias_raw"])


@pytest.mark.gpu
def test_lut_gather_rejects_out_of_table_index(gold):
    import torch
    from pydrodem_b200 import veg
    tc = torch.full((4, 5), 101, dtype=torch.int32, device="cuda")
    th = torch.zeros((4, 5), dtype=torch.int32, device="cuda")
    s = torch.zeros((4, 5), dtype=torch.int8, device="cuda")
    with pytest.raises(RuntimeError, match="outside the lookup table"):
        veg.lut_gather(gold["bias_cube"], tc, th, s, veg.SLOPE_CAP_3D)


@pytest.mark.gpu
def test_compute_veg_bias_3D_matches_reference(gold):
    """Whole function.  Tree-cover / tree-height indices are exact; the slope bin comes from a float32 9x9
    convolution whose summation order differs from scipy's (tolerance class, 2e-6 relative on the smoothed DEM),
    so a cell may fall in the neighbouring slope bin only where the reference's own slope sits that close to a whole
    degree -- everywhere else bias and dem - bias are bit-exact."""
    from pydrodem_b200 import veg
    edm = gold["edm"]
    ocean = np.where(edm == 4, 1, 0)
    wm = ((edm == 2) | (edm == 3)).astype(np.uint8)
    dx, dy = float(gold["dx"]), float(gold["dy"])
    bias, slope, mth, dem = veg.compute_veg_bias_3D(gold["TC"], gold["Theight"], gold["Zarr"], gold["bias_cube"], dx, dy,
                                                    ocean_mask=ocean, watermaskTX=wm, return_dem=True)
    assert mth == int(gold["max_treeheight"])
    assert bias.dtype == np.float64 and dem.dtype == np.float64 and slope.dtype == np.int8
    deg = np.rad2deg(np.arctan(gold["slope_raw"].astype(np.float64)))
    near = np.abs(deg - np.round(deg)) < 1e-3          # reference slope within 0.001 degree of a bin edge
    same = slope == gold["slope_idx"]
    assert np.all(same | near)
    assert np.abs(slope.astype(int) - gold["slope_idx"].astype(int)).max() <= 1
    assert np.array_equal(bias[same], gold["bias"][same])
    assert np.array_equal(dem[same], gold["dem_out"][same])
    assert (~same).mean() < 1e-3
    b3 = veg.compute_veg_bias_3D(gold["TC"], gold["Theight"], gold["Zarr"], gold["bias_cube"], dx, dy, ocean_mask=ocean)
    assert len(b3) == 3 and np.array_equal(b3[0][same], gold["bias_raw"][same])


@pytest.mark.gpu
def test_compute_veg_bias_2Dtable_matches_reference(gold):
    from pydrodem_b200 import veg
    bias, slope = veg.compute_veg_bias_2Dtable(gold["TC"], gold["Zarr"], gold["bias_table"], float(gold["dx"]),
                                               float(gold["dy"]))
    same = slope == gold["slope2"]
    assert (~same).mean() < 1e-3
    assert np.array_equal(bias[same], gold["bias2"][same])


@pytest.mark.gpu
def test_compute_veg_bias_3D_refuses_unfilled_treecover(gold):
    from pydrodem_b200 import veg
    tc = gold["TC"].copy()
    tc[10:20, 10:20] = 255
    with pytest.raises(NotImplementedError):
        veg.compute_veg_bias_3D(tc, gold["Theight"], gold["Zarr"], gold["bias_cube"], 30.0, 30.0)


@pytest.mark.gpu
def test_veg_bias_over_row_bands(gold):
    from pydrodem_b200 import bands, veg
    edm = gold["edm"]
    ocean = np.where(edm == 4, 1, 0).astype(np.uint8)
    dx, dy = float(gold["dx"]), float(gold["dy"])
    whole = veg.compute_veg_bias_3D(gold["TC"], gold["Theight"], gold["Zarr"], gold["bias_cube"], dx, dy, ocean_mask=ocean)
    got = bands.apply_banded(lambda tc, th, z, oc: veg.compute_veg_bias_3D(tc, th, z, gold["bias_cube"], dx, dy,
                                                                            ocean_mask=oc)[:2],
                             [gold["TC"], gold["Theight"], gold["Zarr"], ocean], 3, veg.HALO_ROWS)
    assert np.array_equal(got[0], whole[0]) and np.array_equal(got[1], whole[1])
