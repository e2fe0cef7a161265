"""Golden vectors for wtr.enforce_monotonic_rivers, produced by the REFERENCE'S OWN FUNCTION.

The function's source text is cut out of /root/reference/tools/water.py with `ast` and executed UNMODIFIED; only the
names it reaches outside numpy / scipy are stubbed: `pyct.npy2tif` / `gdal_array.LoadFile` keep rasters in a dictionary
instead of GeoTIFFs, and `wbt.fill_missing_data` (a WhiteboxTools subprocess upstream) is the restatement in
oracle/fill_missing_oracle.py -- so everything up to the final interpolation is pinned by the reference, the interpolated
cells by the oracle.  Run from the repo root:  python tests/golden/make_mono_golden.py
"""
import ast
import os
import sys
import time

import numpy as np
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def reference_function():
    sys.path.insert(0, REF)
    from tools.derive import apply_convolve_filter
    from oracle.fill_missing_oracle import fill_missing_data
    store = {}

    class _Pyct:
        @staticmethod
        def npy2tif(array, ref_tif, out_tif, nodata=None, dtype=None):
            store[out_tif] = (np.array(array, copy=True), nodata)

    class _GdalArray:
        @staticmethod
        def LoadFile(path):
            return store[path][0]

    class _Wbt:
        @staticmethod
        def fill_missing_data(i, output, filter=11, weight=2.0):
            arr, nodata = store[i]
            store[output] = (fill_missing_data(arr, nodata, filter, weight), nodata)

    src = open(os.path.join(REF, "tools", "water.py")).read()
    ns = {"np": np, "ndimage": ndimage, "apply_convolve_filter": apply_convolve_filter, "os": os, "time": time,
          "pyct": _Pyct, "gdal_array": _GdalArray, "print_time": lambda *a, **k: None}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "enforce_monotonic_rivers":
            exec(compile(ast.Module([node], []), "tools/water.py", "exec"), ns)
    return ns["enforce_monotonic_rivers"], _Wbt


def scene(seed, ny=400, nx=420):
    """A flattened river running diagonally down a tilted valley, stepping down 2 m every 60 rows, with raised reaches:
    one 3 m above its surroundings (part of it 5.5 m: it takes a second pass), one too small, one at the frame."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:ny, 0:nx]
    centre = 60 + 0.6 * yy
    d = np.abs(xx - centre)
    wm = np.zeros((ny, nx), np.uint8)
    river = d <= 6
    bank = (d > 6) & (d <= 8)
    wm[river] = 5
    wm[bank] = (1 + (xx + yy) % 2)[bank]
    level = (50.0 - 2.0 * (yy // 60)).astype(np.float32)
    dem = (level + 3.0 + 0.02 * d + rng.normal(0, 0.3, (ny, nx))).astype(np.float32)
    dem[river] = level[river]
    bump = river & (yy >= 200) & (yy <= 230)
    dem[bump] += 3.0
    dem[bump & (yy >= 210) & (yy <= 220)] += 2.5
    small = river & (yy >= 262) & (yy <= 265)
    dem[small] += 1.5
    edge_bump = river & (yy >= 380)
    dem[edge_bump] += 4.0
    swo = np.zeros((ny, nx), np.uint8)
    swo[river] = np.clip(70 + rng.normal(0, 15, (ny, nx)), 0, 100).astype(np.uint8)[river]
    donut = np.zeros((ny, nx), np.uint8)
    donut[:10] = 1
    donut[-10:] = 1
    donut[:, :10] = 1
    donut[:, -10:] = 1
    inner = np.zeros((ny, nx), np.uint8)
    inner[20:-20, 20:-20] = 1
    return dem, swo, wm, donut, inner


def main():
    fn, wbt = reference_function()
    out = {}
    for name, seed, use_inner, radius, strict in (("a", 11, True, 50, True), ("b", 12, False, 50, True), ("c", 13, True, 30, False)):
        dem, swo, wm, donut, inner = scene(seed)
        d_in, w_in = dem.copy(), wm.copy()
        d_out, w_out = fn("dem.tif", "/tmp", dem, swo, wm, donut, name, SWO_cutoff=20, radius=radius,
                          strict_endcap_geom=strict, nodata=-9999, wbt=wbt, max_r_drop=50.0,
                          inner_boundary=inner if use_inner else None, verbose_print=True)
        changed = int((d_out != d_in).sum())
        print(name, "cells changed", changed, "mask changed", int((w_out != w_in).sum()), "classes", np.unique(w_out))
        ch = np.nonzero((d_out != d_in) | (w_out != w_in))
        out.update({f"{name}_seed": np.array(seed), f"{name}_use_inner": np.array(use_inner), f"{name}_radius": np.array(radius),
                    f"{name}_strict": np.array(strict), f"{name}_changed_rc": np.stack(ch).astype(np.int32),
                    f"{name}_dem_changed": d_out[ch], f"{name}_wm_changed": w_out[ch],
                    f"{name}_dem_sum": np.array(d_out.astype(np.float64).sum()), f"{name}_wm_sum": np.array(int(w_out.sum()))})
    np.savez_compressed(os.path.join(HERE, "mono_rivers.npz"), **out)


if __name__ == "__main__":
    main()
