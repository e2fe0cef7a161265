"""Host breach (csrc/breach.cu, Lindsay 2016 as wbt.breach_depressions is called at tools/depressions.py:1376-1380)
against the independent restatement in oracle/breach_oracle.py, hand-worked cases and the properties the algorithm
guarantees.  PARITY UNPINNED against WhiteboxTools itself (binary and source unavailable)."""
import ctypes

import numpy as np
import pytest

from oracle import breach_oracle as bo
from oracle import hydro_oracle as ho
from pydrodem_b200 import _lib, synth

ND = -9999.0


def breach(z, max_length=0, max_depth=0.0, fill_pits=True, do_fill=True, nodata=ND):
    z = np.ascontiguousarray(z, np.float32)
    out = np.empty_like(z)
    st = (ctypes.c_int64 * 5)()
    rc = _lib.load().hc_breach_depressions_host(z.ctypes.data, out.ctypes.data, z.shape[0], z.shape[1], nodata,
                                                max_length, max_depth, int(fill_pits), int(do_fill), st)
    assert rc == 0, _lib.load().hc_last_error()
    return out, list(st)


def scene(seed):
    rng = np.random.default_rng(seed)
    ny, nx = int(rng.integers(8, 44)), int(rng.integers(8, 44))
    z = (synth.make_dem_numpy(ny, nx, seed) * 0.01).astype(np.float32)
    z += rng.integers(0, 4, z.shape).astype(np.float32)          # many ties and pits
    if seed % 3 == 0:
        z[rng.random(z.shape) < 0.05] = ND                        # interior holes
    if seed % 4 == 1:
        z[:3, :5] = ND                                            # exterior nodata corner
    return z


@pytest.mark.parametrize("seed", range(10))
def test_matches_independent_restatement(seed):
    z = scene(seed)
    for ml, md, fp, df in ((0, None, True, True), (3, None, True, True), (6, 1.5, False, True), (2, None, True, False),
                           (0, 0.5, True, True)):
        got, st = breach(z, ml, md or 0.0, fp, df)
        want, sw = bo.breach_depressions(z, ND, ml, md, fp, df)
        assert np.array_equal(got, want), (seed, ml, md, fp, df)
        assert st[:4] == [sw["pits"], sw["breached"], sw["left"], sw["cut"]]


def test_hand_worked_dam():
    """a pit behind a two-cell dam: the channel is cut level at the pit's elevation along the flood's back-links"""
    z = np.array([[9, 9, 9, 9, 9, 9, 9],
                  [9, 8, 8, 8, 8, 8, 9],
                  [9, 8, 2, 7, 6, 1, 0],
                  [9, 8, 8, 8, 8, 8, 9],
                  [9, 9, 9, 9, 9, 9, 9]], np.float32)
    out, st = breach(z, fill_pits=False)
    want = z.copy()
    want[2, 3] = 2
    want[2, 4] = 2
    assert np.array_equal(out, want)
    assert st == [1, 1, 0, 2, 0]
    # a one-cell channel budget cannot reach lower ground: the pit is filled to the dam instead
    out, st = breach(z, max_length=1, fill_pits=False)
    want = z.copy()
    want[2, 2] = 7
    assert np.array_equal(out, want) and st[:4] == [1, 0, 1, 0] and st[4] == 1
    # depth budget: cutting 5 m is refused at 4 m, allowed at 5 m
    assert breach(z, max_depth=4.0, fill_pits=False)[1][:3] == [1, 0, 1]
    assert breach(z, max_depth=5.0, fill_pits=False)[1][:3] == [1, 1, 0]


def test_single_cell_pits_are_filled_first():
    z = np.full((5, 5), 5, np.float32)
    z[2, 2] = 1
    z[1, 2] = 4                                  # lowest neighbour
    out, st = breach(z, fill_pits=True, do_fill=False)
    # the pit is raised to 4; the two-cell flat it now forms with (1, 2) drains through one frame cell cut to 4
    assert out[2, 2] == 4 and st[3] == 1 and np.count_nonzero(out[0] == 4) == 1 and out.min() == 4
    out, st = breach(z, fill_pits=False, do_fill=False)
    assert out[2, 2] == 1 and st[3] >= 1         # breached instead: a neighbour is cut down to 1


@pytest.mark.parametrize("seed", range(6))
def test_properties(seed):
    z = scene(seed + 20)
    valid = z != ND
    filled = ho.fill(z, ND, ho.SEED_WBT)
    scp = bo.fill_single_cell_pits(z, np.float32(ND))
    # unconstrained: everything is breached, nothing is raised above the pit-filled input
    out, st = breach(z)
    assert st[2] == 0 and st[4] == 0
    assert np.all(out[valid] <= scp[valid])
    assert np.array_equal(out[~valid], z[~valid])
    assert np.array_equal(ho.fill(out, ND, ho.SEED_WBT), out)         # depression free
    # constrained: still depression free, never above the plain fill, breach-only stage never raises
    out, st = breach(z, max_length=2)
    assert np.array_equal(ho.fill(out, ND, ho.SEED_WBT), out)
    assert np.all(out[valid] <= filled[valid])
    cut, _ = breach(z, max_length=2, do_fill=False)
    assert np.all(cut[valid] <= scp[valid])
    assert np.array_equal(ho.fill(cut, ND, ho.SEED_WBT), out)         # host fill == oracle fill of the cut surface
    assert st[0] == st[1] + st[2]


def test_arguments():
    lib = _lib.load()
    z = np.zeros((3, 3), np.float32)
    assert lib.hc_breach_depressions_host(z.ctypes.data, z.ctypes.data, 3, 3, ND, 0, 0.0, 1, 1, None) != 0
    assert lib.hc_breach_depressions_host(None, None, 0, 0, ND, 0, 0.0, 1, 1, None) == 0   # empty raster
    out, st = breach(np.full((4, 6), ND, np.float32))
    assert np.all(out == ND) and st == [0, 0, 0, 0, 0]
    out, st = breach(np.arange(12, dtype=np.float32).reshape(1, 12))
    assert np.array_equal(out, np.arange(12, dtype=np.float32).reshape(1, 12))
