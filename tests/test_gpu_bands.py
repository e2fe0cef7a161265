"""Row-band sharding: N bands emulated on ONE GPU (world_size 1, N local bands — the
seam exchanges are the same all_gathers, short-circuited) must reproduce the whole-raster chain and
the CPU oracle bit for bit: filled elevations, D8 codes (flats resolved across seams), slope, counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # (ny, nx, kind, seed, holes, levels, nbands)
    (4, 9, "noise", 21, 0, None, 2),
    (6, 40, "noise", 22, 0, 4, 3),
    (64, 64, "noise", 23, 0, None, 2),
    (65, 130, "noise", 24, 2, 8, 4),
    (130, 97, "noise", 25, 3, 16, 2),
    (200, 300, "cone", 26, 2, 40, 3),
    (200, 300, "bowl", 27, 2, 40, 5),
    (257, 511, "smooth", 28, 3, None, 4),
    (300, 200, "noise", 29, 0, 3, 7),
    (640, 640, "smooth", 30, 4, None, 8),
    (1000, 777, "bowl", 31, 0, 200, 6),
    (90, 333, "bowl", 32, 0, 5, 45),      # two-row bands
]


def _dem(case):
    from pydrodem_b200 import synth
    ny, nx, kind, seed, holes, levels, _ = case
    return synth.random_dem(ny, nx, seed, kind, nodata_holes=holes, levels=levels)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}x{c[1]}-{c[2]}-b{c[6]}")
@pytest.mark.parametrize("table", [1, 0], ids=["taudem", "wbt"])
def test_banded_chain_bit_exact(case, table, oracle, cuda):
    from pydrodem_b200 import bands
    z = _dem(case)
    f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=20.0, seed_mode=1, table=table)
    gf, gd, gs, ga = bands.fill_d8_accum_banded(z, case[6], dx=30.0, dy=20.0, table=table)
    assert np.array_equal(gf, f), f"fill: {(gf != f).sum()} cells differ"
    assert np.array_equal(gd, d), f"dir: {(gd != d).sum()} cells differ"
    assert np.array_equal(gs, s), "slope differs"
    assert np.array_equal(ga, a), f"accum: {(ga != a).sum()} cells differ"


def test_banded_no_flats_and_single_band(oracle, cuda):
    from pydrodem_b200 import bands, synth
    z = synth.random_dem(150, 210, 77, "bowl", nodata_holes=1, levels=30)
    f, d, s, a = oracle.fill_d8_accum(z, dx=10.0, dy=10.0, seed_mode=1, table=1, flats=False)
    for nb in (1, 3):
        gf, gd, gs, ga = bands.fill_d8_accum_banded(z, nb, dx=10.0, dy=10.0, table=1, flats=False)
        assert np.array_equal(gf, f) and np.array_equal(gd, d) and np.array_equal(gs, s) and np.array_equal(ga, a)


def test_banded_c2_recipe(oracle, cuda):
    """A 1500 x 1201 cut of the bench recipe (pits + bowls + 1 cm ties), 4 bands, nodata frame."""
    from pydrodem_b200 import bands, synth
    z = synth.make_dem_numpy(1500, 1201, 1003, nodata_frame=True)
    f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=1, table=1)
    gf, gd, gs, ga = bands.fill_d8_accum_banded(z, 4, dx=30.0, dy=30.0, table=1)
    assert np.array_equal(gf, f)
    assert np.array_equal(gd, d)
    assert np.array_equal(ga, a)


@pytest.mark.parametrize("case", [CASES[3], CASES[5], CASES[7], CASES[9]], ids=lambda c: f"{c[0]}x{c[1]}-{c[2]}-b{c[6]}")
def test_banded_wbt_seed_rule(case, oracle, cuda):
    """WhiteboxTools' seed rule on bands: interior nodata holes are walls, only nodata connected to the raster frame
    seeds; the exterior-nodata raster comes from the banded labelling of the nodata mask."""
    from pydrodem_b200 import bands
    z = _dem(case)
    z[0, :5] = -9999.0          # exterior nodata touching the frame, next to an interior hole in some cases
    f, d, s, a = oracle.fill_d8_accum(z, dx=30.0, dy=20.0, seed_mode=0, table=0)
    gf, gd, gs, ga = bands.fill_d8_accum_banded(z, case[6], dx=30.0, dy=20.0, table=0, seed_mode=0)
    assert np.array_equal(gf, f), f"fill: {(gf != f).sum()} cells differ"
    assert np.array_equal(gd, d) and np.array_equal(gs, s) and np.array_equal(ga, a)


def test_bands_over_nccl_two_ranks(cuda):
    """One band per GPU with the real NCCL exchange (needs >= 2 GPUs; the single-GPU box skips)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run: gpurun --gpus 2 -- python -m pytest tests/test_gpu_bands.py -k nccl)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "run_bands_nccl.py"), "--ny", "2500",
           "--nx", "1800", "--full"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "BANDS_NCCL_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.parametrize("conn", [4, 8])
@pytest.mark.parametrize("shape,p,seed,nbands", [((40, 30), 0.5, 1, 2), ((64, 64), 0.6, 3, 4), ((130, 257), 0.45, 4, 5),
                                                 ((500, 333), 0.55, 5, 7), ((300, 300), 1.0, 7, 3), ((300, 300), 0.0, 8, 3),
                                                 ((90, 120), 0.62, 9, 45)])
def test_banded_labels_match_scipy(shape, p, seed, nbands, conn, oracle, cuda):
    """Row-band labelling: ids identical to scipy.ndimage.label over the WHOLE mask (components snake across many seams at
    p ~ 0.6: the min-root propagation needs several rounds)."""
    from pydrodem_b200 import bands
    rng = np.random.default_rng(seed)
    m = (rng.random(shape) < p).astype(np.uint8)
    want, n_want = oracle.label(m, conn)
    got, n_got = bands.label_banded(m, nbands, conn)
    assert n_got == n_want
    assert np.array_equal(got, want), f"{(got != want).sum()} cells differ (rounds {bands.label_banded.last_rounds})"
