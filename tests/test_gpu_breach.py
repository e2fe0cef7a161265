"""Breach through the product API.  Default: the order-independent flow-path formulation, all passes on the GPU
(`hc_breach_flowpath_dev`), checked against its CPU specification oracle/breach_flowpath.py; the host priority-order
formulation (csrc/breach.cu, method='flood') stays available and is checked against its own host entry point."""
import ctypes
import os

import numpy as np
import pytest

from pydrodem_b200 import _lib, synth

ND = -9999.0


def host_breach(z, max_length=0, max_depth=0.0, fill_pits=True, do_fill=True):
    z = np.ascontiguousarray(z, np.float32)
    out = np.empty_like(z)
    st = (ctypes.c_int64 * 5)()
    assert _lib.load().hc_breach_depressions_host(z.ctypes.data, out.ctypes.data, z.shape[0], z.shape[1], ND, max_length,
                                                  max_depth, int(fill_pits), int(do_fill), st) == 0
    return out, list(st)


def scene(ny, nx, seed):
    rng = np.random.default_rng(seed)
    z = (synth.make_dem_numpy(ny, nx, seed) * 0.02).astype(np.float32)
    z += rng.integers(0, 3, z.shape).astype(np.float32)
    z[:2, :] = ND
    z[:, -3:] = ND
    z[ny // 2:ny // 2 + 5, nx // 3:nx // 3 + 7] = ND      # interior hole: neither a seed nor crossed
    return z


@pytest.mark.gpu
@pytest.mark.parametrize("ny,nx,seed,ml", [(120, 150, 1, 10), (257, 190, 2, 0), (64, 300, 3, 3)])
def test_flood_method_gpu_fill_of_the_cut_surface_equals_host_fill(ny, nx, seed, ml):
    import torch
    import pydrodem_b200 as hc
    z = scene(ny, nx, seed)
    want, st = host_breach(z, max_length=ml, fill_pits=True, do_fill=True)
    got, stats = hc.breach_depressions(z, nodata=ND, max_length=ml or None, fill_pits=True, return_stats=True, method='flood')
    assert got.dtype == np.float32 and np.array_equal(got, want)
    assert [stats["pits"], stats["breached"], stats["left_to_fill"], stats["cells_lowered"]] == st[:4]
    dev = hc.breach_depressions(torch.from_numpy(z).cuda(), nodata=ND, max_length=ml or None, fill_pits=True, method='flood')
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), want)
    assert np.array_equal(hc.fill(got, nodata=ND, seed_mode=hc.SEED_WBT), got)       # depression free
    with pytest.raises(NotImplementedError):
        hc.breach_depressions(z, flat_increment=0.001, method='flood')
    # default method + flat_increment: the breached surface closed by the epsilon fill (float64, like WhiteboxTools' output)
    from oracle import breach_flowpath as bf
    from oracle import hydro_oracle as ho
    cut, _ = bf.breach_flowpath(z, ND, max_length=ml, fill_pits=True, do_fill=False)
    graded = hc.breach_depressions(z, nodata=ND, max_length=ml or None, fill_pits=True, flat_increment=1e-4)
    assert graded.dtype == np.float64 and np.array_equal(graded[z != ND], ho.fill_eps(cut, 1e-4, ND, ho.SEED_WBT)[z != ND])


@pytest.mark.gpu
@pytest.mark.parametrize("ny,nx,seed,ml", [(120, 150, 1, 10), (257, 190, 2, 0), (64, 300, 3, 3), (97, 120, 5, 6)])
def test_default_flowpath_breach_matches_its_specification(ny, nx, seed, ml):
    """the shipped default (every pass on the device), the host-walk composition and the CPU specification agree bit
    for bit, statistics included; numpy in -> numpy out, CUDA tensor in -> CUDA tensor out"""
    import torch
    import pydrodem_b200 as hc
    from oracle import breach_flowpath as bf
    z = scene(ny, nx, seed)
    want, sw = bf.breach_flowpath(z, ND, max_length=ml, fill_pits=True, do_fill=True)
    for method in (None, 'flowpath'):
        kw = {} if method is None else {"method": method}
        got, st = hc.breach_depressions(z, nodata=ND, max_length=ml or None, fill_pits=True, return_stats=True, **kw)
        assert got.dtype == np.float32 and np.array_equal(got, want), method
        assert [st["pits"], st["breached"], st["left_to_fill"], st["cells_lowered"]] == \
            [sw["pits"], sw["breached"], sw["left"], sw["cut"]], method
    dev = hc.breach_depressions(torch.from_numpy(z).cuda(), nodata=ND, max_length=ml or None, fill_pits=True)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), want)
    assert np.array_equal(hc.fill(want, nodata=ND, seed_mode=hc.SEED_WBT), want)     # depression free
    assert np.all(want[z != ND] <= hc.fill(z, nodata=ND, seed_mode=hc.SEED_WBT)[z != ND])   # never above the plain fill


@pytest.mark.gpu
def test_default_flowpath_breach_depth_constraint_and_no_pit_fill():
    """combined_solution's call: max_depth set, fill_pits=False (tools/depressions.py:1012-1014)"""
    import pydrodem_b200 as hc
    from oracle import breach_flowpath as bf
    z = scene(110, 140, 8)
    for md in (0.5, 2.0):
        want, sw = bf.breach_flowpath(z, ND, max_length=9, max_depth=md, fill_pits=False, do_fill=True)
        got, st = hc.breach_depressions(z, nodata=ND, max_length=9, max_depth=md, fill_pits=False, return_stats=True)
        assert np.array_equal(got, want), md
        assert [st["pits"], st["breached"], st["left_to_fill"], st["cells_lowered"]] == \
            [sw["pits"], sw["breached"], sw["left"], sw["cut"]], md


@pytest.mark.gpu
def test_wbt_and_breach_pits_file_protocol(tmp_path):
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools import depressions as dep
    from pydrodem_b200.wbt import WhiteboxToolsGPU
    z = scene(90, 110, 4)
    dem_tif = str(tmp_path / "unit.tif")
    raster_io.write_raster(dem_tif, z, None, nodata=ND)
    wbt = WhiteboxToolsGPU()
    from oracle import breach_flowpath as bf
    carve, diff = dep.breach_pits(dem_tif, 12, 200, wbt, nodataval=ND)
    want, _ = bf.breach_flowpath(z, ND, max_length=12, fill_pits=True, do_fill=True)
    assert np.array_equal(carve, want)
    assert np.array_equal(diff, np.round(want - z, 3))
    assert os.path.exists(str(tmp_path / "unit_SC_12.tif")) and os.path.exists(str(tmp_path / "unit_SC_12_DIFF.tif"))
    # cached: an existing carve raster is re-used, not recomputed (tools/depressions.py:1369)
    raster_io.write_raster(str(tmp_path / "unit_SC_12.tif"), z, None, nodata=ND)
    carve2, diff2 = dep.breach_pits(dem_tif, 12, 200, wbt, nodataval=ND)
    assert np.array_equal(carve2, z) and not diff2.any()
    # (method='LC': tests/test_gpu_breach_leastcost.py)
