"""The property checkers of tests/_properties.py against the oracle at small size: they accept the oracle's outputs and
reject corrupted ones -- before the GPU suite relies on them at 10801^2."""
import numpy as np
import pytest
import torch

from oracle import hydro_oracle as ho
from pydrodem_b200 import synth
from tests import _properties as P

ND = -9999.0


@pytest.mark.parametrize("table,seed_mode", [(ho.TABLE_TAUDEM, ho.SEED_TAUDEM), (ho.TABLE_WBT, ho.SEED_WBT)])
def test_checkers_accept_the_oracle_and_reject_corruption(table, seed_mode):
    z = synth.make_dem_numpy(260, 331, 77)
    z[40:60, 100:130] = ND
    z[:, 0] = ND
    f = ho.fill(z, ND, seed_mode)
    d = ho.resolve_flats(f, ho.d8(f, ND, 30.0, 30.0, table, want_slope=False), ND, table)
    acc = ho.accum(d, table)
    zt, ft, dt, at = (torch.from_numpy(np.ascontiguousarray(x)) for x in (z, f, d, acc.astype(np.int64)))
    assert P.accumulation_violations(dt, at, table) == 0
    assert P.uphill_violations(ft, dt, table) == 0
    assert P.fill_violations(zt, ft, ND, seed_any_nodata=(seed_mode == ho.SEED_TAUDEM)) == 0
    # the downstream map is the oracle's own: mass arrives at the path ends
    down = P.downstream(dt, table).reshape(-1)
    ends = (down < 0) & (dt.reshape(-1) != 255)
    assert int(at.reshape(-1)[ends].sum()) == int((dt != 255).sum())
    # corruption is seen
    a2 = at.clone()
    a2[130, 200] += 1
    assert P.accumulation_violations(dt, a2, table) >= 1
    f2 = ft.clone()
    r, c = np.argwhere((f > z) & (z != ND))[0]
    f2[r, c] = zt[r, c] - 1.0
    assert P.fill_violations(zt, f2, ND, seed_any_nodata=(seed_mode == ho.SEED_TAUDEM)) >= 1
    d2 = dt.clone()
    rr, cc = np.argwhere((d != 0) & (d != 255) & (f > f.min() + 1))[50]
    codes = list(range(1, 9)) if table == ho.TABLE_TAUDEM else [1 << k for k in range(8)]
    worst = 0
    for code in codes:                                   # point the cell at each neighbour in turn: some are uphill
        d2[rr, cc] = code
        worst = max(worst, P.uphill_violations(ft, d2, table))
    assert worst >= 1
