"""CPU suite for the oracle itself: hand-worked micro-DEMs, an independent second algorithm,
domain properties, and the committed golden hashes.  No GPU, no /root/reference at run time."""
import hashlib
import json
import os

import numpy as np
import pytest

N = -9999.0
HERE = os.path.dirname(os.path.abspath(__file__))


def A(rows):
    return np.array(rows, dtype=np.float32)


# ---- hand-worked known answers -------------------------------------------------------------------

def test_single_pit_fills_to_outlet(oracle):
    z = A([[9, 9, 9, 9, 9],
           [9, 5, 5, 5, 9],
           [9, 5, 1, 5, 9],
           [9, 5, 5, 5, 8],
           [9, 9, 9, 9, 9]])
    want = z.copy()
    want[1:4, 1:4] = 8  # the only way out is over the 8 on the east edge
    for mode in (oracle.SEED_WBT, oracle.SEED_TAUDEM):
        assert np.array_equal(oracle.fill(z, N, mode), want)


def test_nested_pits(oracle):
    # inner pit (1) spills at 3 into an outer bowl that spills at 6
    z = A([[9, 9, 9, 9, 9, 9, 9],
           [9, 4, 4, 4, 4, 4, 9],
           [9, 4, 3, 3, 3, 4, 9],
           [9, 4, 3, 1, 3, 4, 6],
           [9, 4, 3, 3, 3, 4, 9],
           [9, 4, 4, 4, 4, 4, 9],
           [9, 9, 9, 9, 9, 9, 9]])
    want = z.copy()
    want[1:6, 1:6] = 6
    assert np.array_equal(oracle.fill(z, N, oracle.SEED_WBT), want)


def test_interior_nodata_hole_seed_conventions(oracle):
    z = A([[9, 9, 9, 9, 9, 9, 9],
           [9, 3, 3, 3, 3, 3, 9],
           [9, 3, 2, 2, 2, 3, 9],
           [9, 3, 2, N, 2, 3, 9],
           [9, 3, 2, 2, 2, 3, 9],
           [9, 3, 3, 3, 3, 3, 7],
           [9, 9, 9, 9, 9, 9, 9]])
    # TauDEM: cells next to ANY nodata keep their elevation -> nothing is raised
    assert np.array_equal(oracle.fill(z, N, oracle.SEED_TAUDEM), z)
    # WBT: the hole is not connected to the outside, so it is a wall: everything pools to 7
    want = z.copy()
    want[1:6, 1:6] = 7
    want[3, 3] = N
    assert np.array_equal(oracle.fill(z, N, oracle.SEED_WBT), want)
    seeds = oracle.seed_mask(z, N, oracle.SEED_WBT)
    assert seeds[1:6, 1:6].sum() == 0 and seeds[0].all() and seeds[:, 0].all()
    assert oracle.seed_mask(z, N, oracle.SEED_TAUDEM)[2:5, 2:5].sum() == 8


def test_exterior_nodata_is_outside(oracle):
    # nodata bay connected to the frame: its neighbours are seeds in both conventions
    z = A([[9, 9, N, 9, 9],
           [9, 2, N, 2, 9],
           [9, 2, 1, 2, 9],
           [9, 2, 2, 2, 9],
           [9, 9, 9, 9, 9]])
    for mode in (oracle.SEED_WBT, oracle.SEED_TAUDEM):
        assert np.array_equal(oracle.fill(z, N, mode), z)  # the pit at 1 touches the bay


def test_pit_on_raster_edge_is_never_raised(oracle):
    z = A([[5, 5, 5], [1, 5, 5], [5, 5, 5]])
    assert np.array_equal(oracle.fill(z, N, oracle.SEED_WBT), z)


def test_walled_in_cells_keep_their_elevation(oracle):
    z = A([[9, 9, 9, 9, 9, 9, 9],
           [9, N, N, N, N, N, 9],
           [9, N, 4, 4, 4, N, 9],
           [9, N, 4, 1, 4, N, 9],
           [9, N, 4, 4, 4, N, 9],
           [9, N, N, N, N, N, 9],
           [9, 9, 9, 9, 9, 9, 9]])
    assert np.array_equal(oracle.fill(z, N, oracle.SEED_WBT), z)  # never reached by the flood


def test_d8_tie_order_and_codes(oracle):
    z = A([[3, 3, 3], [3, 2, 1], [3, 1, 3]])  # equal drops to E and S
    d = oracle.d8(z, N, 1.0, 1.0, oracle.TABLE_WBT, want_slope=False)
    assert d[1, 1] == 2  # WBT searches NE,E,SE,S,...: E first
    zz = np.pad(z, 1, constant_values=9).astype(np.float32)  # TauDEM needs a clean ring
    d = oracle.d8(zz, N, 1.0, 1.0, oracle.TABLE_TAUDEM, want_slope=False)
    assert d[2, 2] == 1  # TauDEM k=1 is E
    z2 = A([[9, 9, 9, 9, 9], [9, 3, 1, 3, 9], [9, 1, 2, 3, 9], [9, 3, 3, 3, 9], [9, 9, 9, 9, 9]])
    d = oracle.d8(z2, N, 1.0, 1.0, oracle.TABLE_TAUDEM, want_slope=False)
    assert d[2, 2] == 3  # N (k=3) comes before W (k=5)
    assert d[0, 0] == oracle.DIR_NODATA and d[4, 2] == oracle.DIR_NODATA  # contaminated border
    d = oracle.d8(z2, N, 1.0, 1.0, oracle.TABLE_WBT, want_slope=False)
    assert d[2, 2] == 32  # WBT: ...S,SW,W,NW,N -> W (32) before N (128)


def test_d8_slope_and_diagonal_length(oracle):
    z = A([[9, 9, 9, 9], [9, 5, 9, 9], [9, 9, 3, 9], [9, 9, 9, 9]])
    d, s = oracle.d8(z, N, 30.0, 30.0, oracle.TABLE_WBT)
    assert d[1, 1] == 4  # SE
    assert s[1, 1] == np.float32(2.0 / np.sqrt(1800.0))
    assert d[2, 2] == 0 and s[2, 2] == 0


def test_accumulation_chain(oracle):
    z = A([[5, 4, 3, 2, 1]])
    d = oracle.d8(z, N, 1.0, 1.0, oracle.TABLE_WBT, want_slope=False)
    assert list(d[0]) == [2, 2, 2, 2, 0]
    assert list(oracle.accum(d, oracle.TABLE_WBT)[0]) == [1, 2, 3, 4, 5]
    z = A([[9] * 7, [9, 5, 4, 3, 2, 1, 0], [9] * 7])
    d = oracle.d8(z, N, 1.0, 1.0, oracle.TABLE_TAUDEM, want_slope=False)
    a = oracle.accum(d, oracle.TABLE_TAUDEM)
    assert list(d[1, 1:6]) == [1, 1, 1, 1, 1] and list(a[1, 1:6]) == [1, 2, 3, 4, 5]
    assert a[0].sum() == 0 and a[1, 6] == 0  # 255 cells hold nothing


def test_flat_with_one_outlet_drains_to_it(oracle):
    z = A([[9, 9, 9, 9, 9, 9],
           [9, 4, 4, 4, 4, 9],
           [9, 4, 4, 4, 4, 3],
           [9, 4, 4, 4, 4, 9],
           [9, 9, 9, 9, 9, 9]])
    d = oracle.d8(z, N, 1.0, 1.0, oracle.TABLE_WBT, want_slope=False)
    assert (d[1:4, 1:4] == 0).all() and d[2, 4] == 2  # only the cell next to the 3 has a direction
    r = oracle.resolve_flats(z, d, N, oracle.TABLE_WBT)
    assert (r[1:4, 1:5] != 0).all()
    a = oracle.accum(r, oracle.TABLE_WBT)
    assert a[2, 5] == z.size  # every cell of the raster ends at the 3 on the east edge
    assert d[1, 4] == 4 and d[3, 4] == 1  # the plateau's east column already sees the 3 diagonally
    assert r[2, 3] == 2 and r[2, 2] == 2 and r[2, 1] == 2  # middle row walks east to the outlet


def test_flat_without_outlet_stays_noflow(oracle):
    z = A([[9, 9, 9, 9], [9, 4, 4, 9], [9, 4, 4, 9], [9, 9, 9, 9]])
    d = oracle.d8(z, N, 1.0, 1.0, oracle.TABLE_WBT, want_slope=False)
    assert np.array_equal(oracle.resolve_flats(z, d, N, oracle.TABLE_WBT), d)


def test_single_cell_pit_tools(oracle):
    z = A([[9, 9, 9, 9, 9], [9, 5, 6, 7, 9], [9, 6, 1, 6, 9], [9, 7, 6, 5, 9], [9, 9, 9, 9, 9]])
    f = oracle.fill_single_cell_pits(z, N)
    want = z.copy()
    want[2, 2] = 5
    assert np.array_equal(f, want)
    z2 = np.full((7, 7), 9, np.float32)
    z2[3, 3] = 4
    z2[3, 5] = 2  # two cells east, lower than the pit: bridge (3,4) is cut to the mean
    b = oracle.breach_single_cell_pits(z2, N)
    assert b[3, 4] == 3 and (b != z2).sum() == 1


def test_label_order_is_raster_first_pixel(oracle):
    m = np.array([[1, 0, 1], [1, 0, 1], [1, 1, 1], [0, 0, 0], [0, 1, 0]], np.uint8)
    lab, n = oracle.label(m, 8)
    assert n == 2 and lab[0, 0] == 1 and lab[0, 2] == 1 and lab[4, 1] == 2
    m = np.array([[1, 0], [0, 1]], np.uint8)
    assert oracle.label(m, 8)[1] == 1 and oracle.label(m, 4)[1] == 2


# ---- independent second algorithm ----------------------------------------------------------------

@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("mode", [0, 1])
def test_priority_flood_equals_planchon_darboux(oracle, seed, mode):
    from pydrodem_b200 import synth
    kind = ["noise", "cone", "bowl", "noise", "smooth", "noise"][seed]
    z = synth.random_dem(40 + 7 * seed, 55 - 3 * seed, 100 + seed, kind, nodata_holes=seed % 3,
                         levels=[None, 12, 30, 4, None, 2][seed])
    assert np.array_equal(oracle.fill(z, N, mode), oracle.fill_planchon_darboux(z, N, mode))


# ---- domain properties ----------------------------------------------------------------------------

@pytest.mark.parametrize("seed", range(4))
def test_chain_properties(oracle, seed):
    from pydrodem_b200 import synth
    z = synth.random_dem(120, 90, 200 + seed, ["noise", "smooth", "bowl", "cone"][seed], nodata_holes=2, levels=25)
    valid = z != np.float32(N)
    for mode, table in ((0, 0), (1, 1)):
        f = oracle.fill(z, N, mode)
        seeds = oracle.seed_mask(z, N, mode).astype(bool)
        assert (f[valid] >= z[valid]).all() and np.array_equal(f[seeds], z[seeds]) and np.array_equal(f[~valid], z[~valid])
        assert np.array_equal(oracle.fill(f, N, mode), f)  # idempotent
        assert set(np.unique(f[valid])) <= set(np.unique(z[valid]))  # no new elevations when eps = 0
        d, s = oracle.d8(f, N, 30.0, 30.0, table)
        r = oracle.resolve_flats(f, d, N, table)
        a = oracle.accum(r, table)
        # mass balance: every counted cell ends in exactly one terminal cell
        counted = r != 255
        k_of = {c: k for k, c in enumerate([1, 2, 4, 8, 16, 32, 64, 128] if table == 0 else range(1, 9))}
        dr = [-1, 0, 1, 1, 1, 0, -1, -1] if table == 0 else [0, -1, -1, -1, 0, 1, 1, 1]
        dc = [1, 1, 1, 0, -1, -1, -1, 0] if table == 0 else [1, 1, 0, -1, -1, -1, 0, 1]
        term = 0
        ny, nx = z.shape
        for i in range(ny):
            for j in range(nx):
                if not counted[i, j]:
                    assert a[i, j] == 0
                    continue
                c = int(r[i, j])
                if c == 0:
                    term += int(a[i, j])
                    continue
                ii, jj = i + dr[k_of[c]], j + dc[k_of[c]]
                if not (0 <= ii < ny and 0 <= jj < nx) or r[ii, jj] == 255:
                    term += int(a[i, j])
                else:
                    assert f[ii, jj] <= f[i, j]  # never uphill
        assert term == int(counted.sum())


# ---- golden hashes (regression pin of the oracle on the seeded generators) ------------------------

def _digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_golden_hashes(oracle):
    with open(os.path.join(HERE, "golden", "oracle_hashes.json")) as fh:
        golden = json.load(fh)
    from tests.golden.make_golden import cases
    for name, z, mode, table in cases():
        f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=mode, table=table)
        assert _digest(z) == golden[name]["input"], f"{name}: generator drifted"
        assert _digest(f) == golden[name]["fill"], name
        assert _digest(d) == golden[name]["dir"], name
        assert _digest(a) == golden[name]["acc"], name


def test_fill_missing_oracle_hand_case():
    """oracle/fill_missing_oracle.py on a hand-worked gap: one nodata cell between two edge cells takes their
    inverse-distance-weighted mean; a cell out of reach stays nodata; valid cells never change"""
    from oracle.fill_missing_oracle import fill_missing_data
    nd = -9999.0
    z = np.full((1, 7), nd, np.float32)
    z[0, 0], z[0, 2] = 10.0, 20.0
    out = fill_missing_data(z, nd, filter=3, weight=2.0)      # reach 1.5 cells
    assert out[0, 0] == 10.0 and out[0, 2] == 20.0
    assert out[0, 1] == np.float32(15.0)                       # two samples at distance 1
    assert out[0, 3] == np.float32(20.0)                       # one sample at distance 1
    assert (out[0, 4:] == nd).all()
    z2 = np.array([[1.0, nd, nd, 4.0]], np.float32)
    out2 = fill_missing_data(z2, nd, filter=7, weight=2.0)    # reach 3.5: both samples, weights 1 and 1/4
    assert out2[0, 1] == np.float32((1.0 * 1.0 + 0.25 * 4.0) / 1.25)
    assert out2[0, 2] == np.float32((0.25 * 1.0 + 1.0 * 4.0) / 1.25)
