"""Host-side logic that needs no GPU: the synthetic-DEM generator and bench.py's reference arm."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generator_is_band_consistent():
    from pydrodem_b200 import synth
    whole = synth.make_dem_numpy(300, 257, 1003, nodata_frame=True, seam_rows=(150,), plateau_rows=(100,))
    parts = [synth.make_dem(300, 257, 1003, nodata_frame=True, seam_rows=(150,), plateau_rows=(100,), row0=r0,
                            nrows=n).numpy() for r0, n in ((0, 77), (77, 100), (177, 123))]
    assert np.array_equal(np.concatenate(parts), whole)
    assert (whole[0] == synth.NODATA).all() and (whole[:, 0] == synth.NODATA).all()
    assert len(np.unique(whole[80:120, 257 // 4:257 // 4 + 100])) == 1  # the plateau is exactly flat


def test_generator_is_seeded_and_quantised():
    from pydrodem_b200 import synth
    a = synth.make_dem_numpy(128, 128, 5)
    b = synth.make_dem_numpy(128, 128, 5)
    c = synth.make_dem_numpy(128, 128, 6)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert np.allclose(a * 100, np.round(a * 100), atol=2e-2)


def test_bench_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-size", "200"], capture_output=True, text=True, check=True, cwd=ROOT)
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mcells/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_percentile_from_order_stats_is_numpys():
    """endo.find_sink's threshold: the GPU supplies two exact order statistics, the interpolation reuses numpy's own
    index / gamma / lerp helpers — so it must equal np.percentile for float32 input, ties and tiny arrays included."""
    from pydrodem_b200.tools import endo
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 10, 101, 4097, 250_000):
        x = rng.normal(0, 50, n).astype(np.float32)
        x[rng.integers(0, n, max(1, n // 7))] = np.float32(1.25)
        s = np.sort(x)
        for q in (0, 1, 2.5, 50, 99, 100):
            got = endo._percentile_from_order_stats(n, q, lambda i: s[i])
            want = np.percentile(x, q)
            assert got == want and np.asarray(got).dtype == np.asarray(want).dtype, (n, q)


def test_band_helpers_without_a_gpu():
    from pydrodem_b200 import bands, stencils
    assert bands.split_rows(40000, 8) == [5000] * 8 and bands.split_rows(7, 3) == [3, 2, 2]
    # sigma from all-reduced partial sums == the two-pass population std
    rng = np.random.default_rng(2)
    x = rng.normal(3.0, 2.0, 100_000)
    parts = np.array_split(x, 5)
    n = sum(p.size for p in parts)
    mean = sum(p.sum() for p in parts) / n
    d1 = sum((p - mean).sum() for p in parts)
    d2 = sum(((p - mean) ** 2).sum() for p in parts)
    assert abs(stencils._sigma_from_sums(n, d1, d2) - np.std(x)) < 1e-12


def test_catalog_api_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("only meaningful on a box without a GPU")
    from pydrodem_b200 import HydrocoreError
    from pydrodem_b200.tools import depressions as dep
    z = np.zeros(16, np.float32)
    ids = np.zeros(16, np.int32)
    with pytest.raises(HydrocoreError):
        dep.build_pit_catalog(z, -9999.0, z, z, z, ids, ids, ids, verbose_print=False)
    with pytest.raises(NameError):        # the reference's own behaviour without a water mask
        dep.select_solution(None, {}, {}, {}, z, z, ids, ids, 1.0, 1.0, 1.0)
    with pytest.raises(ValueError):
        dep.select_solution(None, {}, {}, {}, z, z, ids, ids, 1.0, 1.0, 1.0, water_mask_array=z, numpy_scalars="float16")


def test_golden_files_are_present_and_described():
    g = os.path.join(ROOT, "tests", "golden")
    for f, gen in (("stencils_reference.npz", "make_stencil_golden.py"), ("catalog_reference.npz", "make_catalog_golden.py"),
                   ("oracle_hashes.json", "make_golden.py"), ("veg_reference.npz", "make_veg_golden.py"),
                   ("water_reference.npz", "make_water_golden.py")):
        assert os.path.exists(os.path.join(g, f)) and os.path.exists(os.path.join(g, gen)), f
    d = np.load(os.path.join(g, "catalog_reference.npz"))
    codes = set()
    for k in d.files:
        if k.endswith("_solution"):
            codes |= set(int(v) for v in np.unique(d[k]))
    assert {1, 7, 1002, 1006, 1200, 2000, 2004, 2006, 2009, 2100, 4000} <= codes


def test_fix_shifted_data_repairs_a_whole_raster_shift():
    """tools/depressions.py:1406-1468: a WhiteboxTools output shifted by (+1, +2) cells is moved back and the
    rounded difference map rebuilt; an unshifted raster passes through"""
    from pydrodem_b200.tools import depressions as dep
    rng = np.random.default_rng(0)
    nd = -9999.0
    z = np.full((12, 15), nd, np.float32)
    z[2:9, 3:11] = rng.random((7, 8)).astype(np.float32) * 10
    filled = z.copy()
    filled[4:6, 5:8] += 1.25
    shifted = np.full_like(z, nd)
    shifted[3:10, 5:13] = filled[2:9, 3:11]
    out, diff = dep.fix_shifted_data(z, shifted.copy(), nd)
    assert np.array_equal(out, filled)
    assert np.array_equal(diff, np.round(filled - z, 3)) and diff.max() == 1.25
    out2, diff2 = dep.fix_shifted_data(z, filled.copy(), nd)
    assert np.array_equal(out2, filled) and np.array_equal(diff2, diff)


def test_water_veg_and_breach_apis_fail_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("only meaningful on a box without a GPU")
    from pydrodem_b200 import veg
    from pydrodem_b200.tools import water
    z = np.zeros((6, 7), np.float32)
    m = np.zeros((6, 7), np.uint8)
    for call in (lambda: water.river_minimize(z, m, m), lambda: water.raise_small_streams(z, z.copy(), m, -9999),
                 lambda: water.minimum_filter(z, 5), lambda: veg.box3_index(m),
                 lambda: veg.compute_veg_bias_3D(m, m, z, np.zeros((101, 5, 20)), 30.0, 30.0)):
        with pytest.raises((RuntimeError, AssertionError)):      # torch refuses .cuda(); nothing computes on the CPU
            call()
    import pydrodem_b200 as hc
    with pytest.raises(Exception):                               # the closing fill needs the GPU: no host substitute
        hc.breach_depressions(z)


def test_window_meta_shifts_the_tiepoint_only():
    from pydrodem_b200 import raster_io
    from pydrodem_b200.tools.depressions import _window_meta
    meta = raster_io.RasterMeta({"nodata": -9999.0, "geo": {33550: (12, [30.0, 30.0, 0.0]),
                                                            33922: (12, [0.0, 0.0, 0.0, 500000.0, 4100000.0, 0.0])}})
    w = _window_meta(meta, 10, 4)
    assert list(w["geo"][33922][1]) == [0.0, 0.0, 0.0, 500000.0 + 4 * 30.0, 4100000.0 - 10 * 30.0, 0.0]
    assert list(meta["geo"][33922][1])[3:5] == [500000.0, 4100000.0] and w["nodata"] == -9999.0
    assert _window_meta(None, 3, 3).get("geo") == {}


def test_sink_pixel_geometry_round_trip():
    from pydrodem_b200.tools import endo
    gt = (12.5, 0.000277, 0.0, 45.25, 0.0, -0.000277)
    x, y = endo.get_xy(17, 41, gt)
    assert endo.coords_to_pixel(x, y, gt[1], gt[5], gt[0], gt[3]) == (41, 17)
    a = np.zeros((30, 50), np.float32)
    out = endo.set_nodata_points((x, y), a, gt[1], gt[5], gt[0], gt[3], -9999)
    assert out is a and a[17, 41] == -9999 and np.count_nonzero(a) == 1


def test_reference_module_names_resolve(tmp_path):
    """the import swaps INTEGRATION.md shows: tools.derive / tools.build_operators / tools.convert / tools.water ..."""
    import importlib
    from pydrodem_b200 import raster_io
    for m, names in (("derive", ("apply_convolve_filter", "slope_D2", "slope_D4", "curvature_D2", "DEMdifference")),
                     ("build_operators", ("build_kernel",)), ("convert", ("npy2tif", "tifdowncast", "tifcompress")),
                     ("water", ("river_minimize", "smooth_rivers", "raise_small_streams", "raise_small_rivers")),
                     ("depressions", ("fill_pits", "breach_pits", "initial_fill_carve", "create_clump_array",
                                      "fill_shallow_pits", "build_pit_catalog", "select_solution", "apply_solution",
                                      "combined_solution", "fix_shifted_data")),
                     ("endo", ("find_sink_pixel", "get_xy", "coords_to_pixel", "set_nodata_points"))):
        mod = importlib.import_module(f"pydrodem_b200.tools.{m}")
        for n in names:
            assert callable(getattr(mod, n)), (m, n)
    pdt = importlib.import_module("pydrodem_b200.tools.derive")
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    raster_io.write_raster(str(tmp_path / "a.tif"), a, None, nodata=-9999)
    raster_io.write_raster(str(tmp_path / "b.tif"), a * 2 + 1, None, nodata=-9999)
    d = pdt.DEMdifference(str(tmp_path / "a.tif"), str(tmp_path / "b.tif"), "d.tif", -9999)
    assert np.array_equal(d, a + 1) and np.array_equal(raster_io.read_raster(str(tmp_path / "d.tif"))[0], a + 1)
    k = importlib.import_module("pydrodem_b200.tools.build_operators").build_kernel([1., 0.5], dtype=np.float32)
    assert k.shape == (3, 3) and k[1, 1] == 1.0 and k[0, 0] == 0.5


def test_clipbuffer_arr_matches_the_reference_semantics():
    from pydrodem_b200 import stencils
    a = np.arange(56, dtype=np.float32).reshape(7, 8)
    assert np.array_equal(stencils.clipbuffer_arr(a, 2), a[2:-2, 2:-2])
    mask = np.zeros_like(a, dtype=np.int8)
    mask[2:-2, 2:-2] = 1
    assert np.array_equal(stencils.clipbuffer_arr(a, 2, set_nodata=-9999), np.where(mask == 1, a, -9999))
    u = np.arange(56, dtype=np.uint8).reshape(7, 8)
    assert np.array_equal(stencils.clipbuffer_arr(u, 1, set_nodata=99)[0], np.full(8, 99, np.uint8)) and u[0, 0] == 0


def test_unit_discovery_helpers(tmp_path):
    """tools/log.py: which `<unit>_<id>` directories exist / are unfinished (reference tools/log.py:66-106, 189-195)"""
    import os
    from pydrodem_b200.tools import log as logt
    for name in ("basin_6010", "basin_6020", "tile_N10E020", "notes"):
        os.makedirs(tmp_path / name)
    open(tmp_path / "basin_6010" / "DEMcompleted.txt", "w").close()
    open(tmp_path / "basin_stray.txt", "w").close()
    assert logt.find_all_ids(str(tmp_path), unit="basin", dir_search="basin_") == [6010, 6020]
    assert logt.find_unfinished_ids(str(tmp_path), unit="basin", dir_search="basin_", tifname="DEMcompleted.txt") == [6020]
    assert logt.find_all_ids(str(tmp_path), unit="tile", dir_search="tile_") == ["N10E020"]
    assert logt.find_all_basins_IDs(str(tmp_path)) == [6010, 6020]
    import time
    assert logt.log_time("X", time.time() - 3725).startswith("X COMPLETE -- elapsed 01:02:0")


def test_u64_order_flip_turns_signed_reductions_into_unsigned_ones():
    """bands.BandedCatalog reduces 64-bit (value, cell index) keys across ranks with signed int64 MAX / MIN all_reduces"""
    import torch
    from pydrodem_b200.bands import u64_order_flip
    keys = [0, 1, 0x7FFFFFFFFFFFFFFF, 0x8000000000000000, 0x8000000000000001, 0xFFFFFFFFFFFFFFFE, 0xFFFFFFFFFFFFFFFF,
            (0xC2F00000 << 32) | 12345, (0x3D100000 << 32) | 0xFFFFFFFF]
    as_i64 = torch.tensor([k - (1 << 64) if k >= (1 << 63) else k for k in keys], dtype=torch.int64)
    flipped = u64_order_flip(as_i64)
    order_unsigned = sorted(range(len(keys)), key=lambda i: keys[i])
    order_flipped = sorted(range(len(keys)), key=lambda i: int(flipped[i]))
    assert order_unsigned == order_flipped
    assert torch.equal(u64_order_flip(flipped), as_i64)
    a, b = as_i64[:5], as_i64[4:9]
    n = min(len(a), len(b))
    got = u64_order_flip(torch.maximum(u64_order_flip(a[:n]), u64_order_flip(b[:n])))
    want = [max(keys[i], keys[4 + i]) for i in range(n)]
    assert [int(v) % (1 << 64) for v in got] == want
