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
