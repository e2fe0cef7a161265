"""The order-independent breach formulation planned for the device (oracle/breach_flowpath.py):
its defining properties, checked on the CPU so that the kernel that will implement it has a pinned specification."""
import numpy as np
import pytest

from oracle import breach_flowpath as bf
from oracle import breach_oracle as bo
from oracle import hydro_oracle as ho
from pydrodem_b200 import synth

ND = -9999.0


def scene(seed):
    rng = np.random.default_rng(seed)
    ny, nx = int(rng.integers(10, 40)), int(rng.integers(10, 40))
    z = (synth.make_dem_numpy(ny, nx, seed) * 0.01).astype(np.float32)
    z += rng.integers(0, 4, z.shape).astype(np.float32)
    if seed % 3 == 0:
        z[rng.random(z.shape) < 0.04] = ND
    if seed % 4 == 1:
        z[:3, :5] = ND
    return z


@pytest.mark.parametrize("seed", range(8))
def test_result_does_not_depend_on_the_order_of_the_pits(seed):
    z = scene(seed)
    rng = np.random.default_rng(seed)
    for ml in (0, 4):
        base, st = bf.breach_flowpath(z, ND, max_length=ml, do_fill=False)
        for _ in range(3):
            perm = rng.permutation(st["pits"])
            other, st2 = bf.breach_flowpath(z, ND, max_length=ml, do_fill=False, order=perm)
            assert np.array_equal(other, base)
            assert (st2["breached"], st2["left"]) == (st["breached"], st["left"])


@pytest.mark.parametrize("seed", range(8))
def test_unconstrained_breach_drains_everything_without_raising(seed):
    z = scene(seed + 50)
    valid = z != ND
    w0 = bo.fill_single_cell_pits(z, np.float32(ND))
    cut, st = bf.breach_flowpath(z, ND, do_fill=False)
    assert st["left"] == 0
    assert np.all(cut[valid] <= w0[valid]) and np.array_equal(cut[~valid], z[~valid])
    assert np.array_equal(ho.fill(cut, ND, ho.SEED_WBT), cut)            # already depression free: the fill is a no-op
    # every cut cell now sits exactly at the level of some pit upstream of it
    changed = cut != w0
    assert set(np.unique(cut[changed])) <= set(np.unique(w0[valid]))


@pytest.mark.parametrize("seed", range(6))
def test_constrained_breach_then_fill(seed):
    z = scene(seed + 90)
    valid = z != ND
    out, st = bf.breach_flowpath(z, ND, max_length=3, max_depth=2.0)
    assert st["pits"] == st["breached"] + st["left"]
    assert np.array_equal(ho.fill(out, ND, ho.SEED_WBT), out)
    assert np.all(out[valid] <= ho.fill(z, ND, ho.SEED_WBT)[valid])
    # the flood formulation (the shipped one) tests "is a pit" on the surface as already cut, this one on the original
    # surface: it can only see more pits, never fewer
    _, sf = bo.breach_depressions(z, ND, max_length=3, max_depth=2.0)
    assert sf["pits"] <= st["pits"]


def test_hand_worked_dam_matches_the_flood_formulation():
    z = np.array([[9, 9, 9, 9, 9, 9, 9],
                  [9, 8, 8, 8, 8, 8, 9],
                  [9, 8, 2, 7, 6, 1, 0],
                  [9, 8, 8, 8, 8, 8, 9],
                  [9, 9, 9, 9, 9, 9, 9]], np.float32)
    a, sa = bf.breach_flowpath(z, ND, fill_pits=False)
    b, sb = bo.breach_depressions(z, ND, fill_pits=False)
    assert np.array_equal(a, b) and sa["cut"] == sb["cut"] == 2


@pytest.mark.parametrize("seed", range(10))
def test_host_walk_equals_the_specification(seed):
    """hc_breach_flowpath_host (the pointer-chasing half, product code) fed the oracle's filled surface and pointers"""
    import ctypes
    from pydrodem_b200 import _lib
    z = scene(seed + 130)
    nd = np.float32(ND)
    w = bo.fill_single_cell_pits(z, nd)
    seed, filled, d = bf.pointers(w, nd)
    got_seed = np.empty(w.shape, np.uint8)
    assert _lib.load().hc_wbt_seed_mask_host(w.ctypes.data, got_seed.ctypes.data, w.shape[0], w.shape[1], ND) == 0
    assert np.array_equal(got_seed.astype(bool), seed)
    for ml, md in ((0, 0.0), (3, 0.0), (5, 1.5)):
        out = np.empty_like(w)
        st = (ctypes.c_int64 * 4)()
        rc = _lib.load().hc_breach_flowpath_host(w.ctypes.data, filled.ctypes.data, d.ctypes.data, out.ctypes.data,
                                                 w.shape[0], w.shape[1], ND, ml, md, st)
        assert rc == 0, _lib.load().hc_last_error()
        want, sw = bf.breach_flowpath(z, ND, max_length=ml, max_depth=md or None, fill_pits=True, do_fill=False)
        assert np.array_equal(out, want)
        assert list(st) == [sw["pits"], sw["breached"], sw["left"], sw["cut"]]
