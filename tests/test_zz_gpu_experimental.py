"""Opt-in checks of EXPERIMENTAL compositions that have not had a GPU run yet (HC_TEST_EXPERIMENTAL=1 enables them).
They are kept out of the default `-m gpu` selection so that an unverified path can never mask the verified suite."""
import os

import numpy as np
import pytest

run = os.environ.get("HC_TEST_EXPERIMENTAL") == "1"


@pytest.mark.gpu
@pytest.mark.skipif(not run, reason="experimental: set HC_TEST_EXPERIMENTAL=1")
def test_flowpath_breach_composition_matches_its_specification():
    import pydrodem_b200 as hc
    from oracle import breach_flowpath as bf
    from pydrodem_b200 import synth
    rng = np.random.default_rng(5)
    for seed in range(4):
        z = (synth.make_dem_numpy(90 + 7 * seed, 120, seed) * 0.02).astype(np.float32)
        z += rng.integers(0, 3, z.shape).astype(np.float32)
        z[:2, :] = -9999.0
        for ml in (0, 6):
            want, sw = bf.breach_flowpath(z, -9999.0, max_length=ml, fill_pits=True, do_fill=True)
            for method in ('flowpath', 'flowpath_dev'):
                got, st = hc.breach_depressions(z, nodata=-9999.0, max_length=ml or None, fill_pits=True,
                                                return_stats=True, method=method)
                assert np.array_equal(got, want), method
                assert [st["pits"], st["breached"], st["left_to_fill"], st["cells_lowered"]] == \
                    [sw["pits"], sw["breached"], sw["left"], sw["cut"]], method
