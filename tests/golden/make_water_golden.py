"""Golden vectors for the water stencils, produced by the REFERENCE'S OWN FUNCTIONS.

`tools/water.py` cannot be imported here (geopandas / GDAL / shapely are not installed), but the four
array-level functions on the path -- river_minimize, smooth_rivers, raise_small_streams,
raise_small_rivers -- only need numpy, scipy.ndimage and `tools.derive.apply_convolve_filter`.  Their
source text is cut out of /root/reference/tools/water.py with `ast` and executed UNMODIFIED in a namespace
holding exactly those three names, so the stored outputs are the reference's.
Run from the repo root:  python tests/golden/make_water_golden.py
"""
import ast
import os
import sys

import numpy as np
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FUNCS = ("river_minimize", "smooth_rivers", "raise_small_streams", "raise_small_rivers")


def reference_functions():
    sys.path.insert(0, REF)
    from tools.derive import apply_convolve_filter
    src = open(os.path.join(REF, "tools", "water.py")).read()
    tree = ast.parse(src)
    ns = {"np": np, "ndimage": ndimage, "apply_convolve_filter": apply_convolve_filter}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in FUNCS:
            exec(compile(ast.Module([node], []), "tools/water.py", "exec"), ns)
    return {f: ns[f] for f in FUNCS}


def scene(ny, nx, seed):
    """DEM + water mask (0 land, 1-2 river banks, 3-7 river classes, 8-15 stream classes) + SWO percent"""
    from pydrodem_b200 import synth
    rng = np.random.default_rng(seed)
    dem = synth.make_dem_numpy(ny, nx, seed, quantise=False)
    dem = (dem * 0.004 + rng.normal(0, 0.4, dem.shape)).astype(np.float32)   # gentle relief: a few metres across the scene
    yy, xx = np.mgrid[0:ny, 0:nx]
    wm = np.zeros((ny, nx), np.uint8)
    swo = np.zeros((ny, nx), np.uint8)
    # a meandering river, half-width varying 2..6 cells, classes 3..7 inside, banks 1..2 around
    centre = nx * 0.45 + nx * 0.18 * np.sin(yy / max(ny, 1) * 5.0) + 3 * np.sin(yy / 3.0)
    half = 2 + 4 * (0.5 + 0.5 * np.sin(yy / 11.0))
    d = np.abs(xx - centre)
    river = d <= half
    bank = (d > half) & (d <= half + 2)
    wm[river] = (3 + (yy // 7) % 5)[river]
    wm[bank] = (1 + (xx + yy) % 2)[bank]
    swo[river] = np.clip(60 + 40 * np.cos(yy / 9.0) + rng.normal(0, 15, (ny, nx)), 0, 100).astype(np.uint8)[river]
    swo[bank] = rng.integers(0, 12, (ny, nx)).astype(np.uint8)[bank]
    # a second, narrow river with little observed water
    r2 = (np.abs(yy - (ny * 0.7 + 4 * np.sin(xx / 6.0))) <= 1) & (wm == 0)
    wm[r2] = 4
    swo[r2] = rng.integers(0, 9, (ny, nx)).astype(np.uint8)[r2]
    # tributary streams (classes 8..15, including the excluded 11 and 15), one cell wide
    for k, row in enumerate(range(6, ny - 4, 9)):
        cols = np.arange(0, nx)
        rr = np.clip(row + (np.sin(cols / 5.0 + k) * 2).astype(int), 0, ny - 1)
        sel = wm[rr, cols] == 0
        wm[rr[sel], cols[sel]] = 8 + (k % 8)
    # flattened water surface: rivers a little below the terrain, with noise so that the moving minimum matters
    burn = dem.copy()
    riv_all = (wm > 0) & (wm < 8)
    burn[riv_all] -= (1.5 + rng.random((ny, nx)) * 1.5).astype(np.float32)[riv_all]
    dry = riv_all & (swo < 5)                              # little observed water: sits lower than the wet channel
    burn[dry] -= (0.5 + rng.random((ny, nx)) * 1.5).astype(np.float32)[dry]
    strm = wm > 7
    burn[strm] -= (rng.random((ny, nx)) * 6.0 - 1.0).astype(np.float32)[strm]
    # nodata holes, some on water
    # (few and in one corner: the reference's 5x5 mean averages -9999 in, which drags every moving minimum within
    # reach of a nodata cell down to where no rule fires)
    nod = np.zeros((ny, nx), bool)
    nod[:4, :9] = True
    nod[ny // 8, int(centre[ny // 8, 0])] = True          # one on the river
    nod[min(6, ny - 1), nx // 5] = True                    # one on the first stream row
    dem[nod] = -9999.0
    burn[nod] = -9999.0
    return dem, burn.astype(np.float32), wm, swo


def main():
    ref = reference_functions()
    out = {}
    cases = {"a": (150, 181, 11), "b": (23, 29, 12), "c": (64, 40, 13)}
    for name, (ny, nx, seed) in cases.items():
        dem, burn, wm, swo = scene(ny, nx, seed)
        out[f"{name}_dem"], out[f"{name}_burn"], out[f"{name}_wm"], out[f"{name}_swo"] = dem, burn, wm, swo
        out[f"{name}_rm_default"] = ref["river_minimize"](burn.copy(), wm, swo)
        out[f"{name}_rm_call"] = ref["river_minimize"](burn.copy(), wm, swo, max_burn=2, river_filter_size=5,
                                                       nodataval=-9999)                     # depression_handling.py:647
        b = burn.copy()
        out[f"{name}_rsr_call"] = ref["raise_small_rivers"](b, b, wm, swo, -9999, SWO_cutoff_low=5, SWO_cutoff_high=50,
                                                            filter_size=31, extra_raise=0.10, max_raise=2.0)  # :889
        out[f"{name}_rsr_default"] = ref["raise_small_rivers"](dem, burn.copy(), wm, swo, -9999)
        out[f"{name}_rss_call"] = ref["raise_small_streams"](dem, burn.copy(), wm, -9999, extra_raise=1.0,
                                                             max_raise=2.0)                 # :898
        out[f"{name}_rss_default"] = ref["raise_small_streams"](dem, burn.copy(), wm, -9999)
        b = burn.copy()
        out[f"{name}_rss_alias"] = ref["raise_small_streams"](b, b, wm, -9999, extra_raise=0.25, max_raise=1.0)
        out[f"{name}_sr_default"] = ref["smooth_rivers"](dem, burn.copy(), wm, swo, -9999)
        out[f"{name}_sr_wide"] = ref["smooth_rivers"](dem, burn.copy(), wm, swo, -9999, SWO_cut=40, max_lower=3.0,
                                                      filter_size=9)
        # the primitives
        out[f"{name}_min51"] = ndimage.minimum_filter(burn, size=51)
        out[f"{name}_max35_u8"] = ndimage.maximum_filter(wm, size=35)
        out[f"{name}_min4"] = ndimage.minimum_filter(burn, size=4)
        print(name, {k[2:]: int(np.count_nonzero(out[k] != burn)) for k in out
                     if k.startswith(name + "_r") or k.startswith(name + "_s") and k[2:] != "swo"})
    # an all-land raster: every function is the identity
    dem, burn, wm, swo = scene(40, 50, 14)
    wm0 = np.zeros_like(wm)
    assert np.array_equal(ref["river_minimize"](burn.copy(), wm0, swo), burn)
    assert np.array_equal(ref["raise_small_streams"](dem, burn.copy(), wm0, -9999), burn)
    path = os.path.join(HERE, "water_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", {k: str(v.dtype) for k, v in out.items() if k.startswith("b_")})


if __name__ == "__main__":
    main()
