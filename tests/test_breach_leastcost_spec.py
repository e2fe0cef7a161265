"""The least-cost breach specification (oracle/breach_leastcost.py) on hand-worked surfaces and through its properties:
it is what csrc/leastcost.cu is compared with bit for bit on the GPU (tests/test_gpu_breach_leastcost.py)."""
import numpy as np

from oracle import breach_leastcost as lc
from oracle import hydro_oracle as ho

ND = -9999.0


def ramp(ny, nx):
    """a plane that drains to the right: no pits"""
    return (100.0 - np.arange(nx, dtype=np.float32))[None, :].repeat(ny, 0).copy()


def test_no_pit_no_change():
    z = ramp(9, 12)
    out, st = lc.breach_least_cost(z, ND, dist=5, return_stats=True)
    assert st["pits"] == 0 and np.array_equal(out, z)


def test_single_pit_cuts_the_cheapest_wall():
    z = ramp(9, 12)
    z[4, 5] = 80.0                       # a hole in the slope: neighbours 96..94, lower ground (< 80) from column 21 on: none
    z[:, 9:] -= 30.0                     # ... so drop the right part: columns 9.. are 61, 60, 59
    # walls between the pit (column 5) and column 9: columns 6, 7, 8 on row 4
    z[4, 6:9] = (81.0, 92.0, 91.0)       # row 4 is the cheap way (1 + 12 + 11 above 80), any other cell costs a metre more
    out, st = lc.breach_least_cost(z, ND, dist=6, return_stats=True)
    assert st == dict(pits=1, breached=1, left=0, cells_lowered=3)
    assert np.array_equal(out[4, 6:9], np.float32([80, 80, 80]))
    changed = out != z
    changed[4, 6:9] = False
    assert not changed.any()


def test_gradient_and_max_cost_and_window():
    z = ramp(9, 12)
    z[4, 5] = 80.0
    z[:, 9:] -= 30.0
    z[4, 6:9] -= 1.0                     # the straight way is the cheapest: 13 + 12 + 11 = 36 m
    out = lc.breach_least_cost(z, ND, dist=6, flat_increment=0.5)
    assert np.array_equal(out[4, 6:9], np.float32([79.5, 79.0, 78.5]))     # 80 - i * 0.5, target (61) untouched
    assert out[4, 9] == z[4, 9]
    assert lc.breach_least_cost(z, ND, dist=6, max_cost=35.9, return_stats=True)[1]["left"] == 1
    assert lc.breach_least_cost(z, ND, dist=6, max_cost=36.0, return_stats=True)[1]["breached"] == 1
    # the lower ground starts 4 cells away: a window of radius 3 does not see it
    assert lc.breach_least_cost(z, ND, dist=3, return_stats=True)[1]["left"] == 1


def test_min_dist_prefers_the_short_way():
    z = np.full((11, 11), 50.0, np.float32)
    z[5, 5] = 10.0                       # pit in a plateau
    z[5, 8] = 5.0                        # lower cell 3 steps east behind two 40 m wall cells
    z[4, 6:8] = z[6, 6:8] = 51.0         # (the straight way east is strictly the cheapest)
    z[5, 3] = 5.0                        # lower cell 2 steps west behind one very high wall cell
    z[4:7, 4] = (600.0, 500.0, 600.0)
    cheap = lc.breach_least_cost(z, ND, dist=4, min_dist=False)
    short = lc.breach_least_cost(z, ND, dist=4, min_dist=True)
    assert cheap[5, 6] == 10.0 and cheap[5, 7] == 10.0 and cheap[5, 4] == 500.0
    assert short[5, 4] == 10.0 and short[5, 6] == 50.0


def test_flat_has_one_representative_and_edges_drain():
    z = np.full((8, 8), 20.0, np.float32)
    z[3:5, 3:5] = 10.0                   # a 2 x 2 flat pit: one search, toward the frame (edge cells are targets)
    out, st = lc.breach_least_cost(z, ND, dist=4, return_stats=True)
    assert st["pits"] == 1 and st["breached"] == 1
    # the edge target is no lower than the pit, so it is cut too: the channel reaches the frame
    assert (out < z).sum() >= 2
    filled = ho.fill(out, ND, ho.SEED_TAUDEM)
    assert np.array_equal(filled, out)   # nothing left to fill: the flat drains off the raster


def test_properties_on_noise():
    rng = np.random.default_rng(5)
    z = rng.integers(0, 40, (40, 50)).astype(np.float32)
    z[10:13, 20:24] = ND
    out, st = lc.breach_least_cost(z, ND, dist=8, return_stats=True)
    assert (out <= z).all() and st["pits"] == st["breached"] + st["left"]
    # breaching can only reduce what a fill has to add afterwards
    v0 = (ho.fill(z, ND, ho.SEED_TAUDEM) - z)[z != ND].sum()
    v1 = (ho.fill(out, ND, ho.SEED_TAUDEM) - out)[z != ND].sum()
    assert v1 <= v0
    # with an unlimited budget and a window that sees the whole raster every pit drains: a fill adds nothing
    out_all = lc.breach_least_cost(z, ND, dist=60)
    assert np.array_equal(ho.fill(out_all, ND, ho.SEED_TAUDEM), out_all)
