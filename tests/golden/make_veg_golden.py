"""Golden vectors for the vegetation-bias lookup (models/dem_noise_removal.py:633-807, 556-631).

`CreateDTM.compute_veg_bias_3D` is a method wrapped in GDAL / JSON file I/O and uses `np.int`, which the
numpy in this image no longer has, so it cannot be called as it stands.  Its array statements are replayed
here one by one (`np.int` -> `int`), calling the REFERENCE's own `pdt.apply_convolve_filter`, `pdt.slope_D2`
and `bop.build_kernel` (imported unmodified from /root/reference) for the smoothing and slope, and the same
per-pixel list comprehension for the lookup.  Run from the repo root:  python tests/golden/make_veg_golden.py
"""
import os
import sys

import numpy as np
import scipy.ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make_table(rng, n_tc=101, n_th=41, n_s=20):
    """A synthetic veg_bias_3d.json: {str(slope): [[bias over tree height] over tree cover]}"""
    tc = np.arange(n_tc)[:, None] / 100.0
    th = np.arange(n_th)[None, :]
    data = {}
    for s in range(n_s):
        plane = 0.55 * tc * th * (1.0 - 0.02 * s) + rng.normal(0, 0.3, (n_tc, n_th))
        data[str(s)] = np.round(plane, 4).tolist()
    return data


def main():
    sys.path.insert(0, REF)
    import tools.build_operators as bop
    import tools.derive as pdt
    from pydrodem_b200 import synth

    rng = np.random.default_rng(2024)
    ny, nx = 137, 151
    dx, dy = 24.0, 30.0
    Zarr = synth.make_dem_numpy(ny, nx, 77, quantise=False)
    Zarr = (Zarr * 3.0 + rng.normal(0, 1.0, Zarr.shape)).astype(np.float32)

    # tree cover (Byte, percent; a handful of nodata 255 cells, fewer than the 10 the reference tolerates)
    yy, xx = np.mgrid[0:ny, 0:nx]
    TC0 = (50 + 50 * np.sin(yy / 17.0) * np.cos(xx / 23.0) + rng.normal(0, 8, (ny, nx))).clip(0, 100).astype(np.uint8)
    TC0[40:60, 30:70] = 100
    TC0[90:120, 100:140] = 0
    for r, c in ((5, 5), (70, 80), (136, 150), (0, 149), (99, 3), (50, 50), (51, 50)):
        TC0[r, c] = 255
    Th0 = (18 + 16 * np.cos(yy / 11.0) + rng.normal(0, 5, (ny, nx))).clip(0, 60).astype(np.uint8)
    edm = np.zeros((ny, nx), np.uint8)
    edm[:, :12] = 4            # ocean strip
    edm[60:75, 90:110] = 2     # TanDEM-X water body
    edm[20:24, 60:66] = 3
    ocean_mask = np.where(edm == 4, 1, 0)
    watermaskTX = np.where((edm == 2) | (edm == 3))

    data = make_table(rng)

    # ---- :660-683 ----
    sample_arr = np.array(data.get(list(data.keys())[0]))
    slope_axis = np.array(list(data.keys())).astype(int)
    max_treeheight = sample_arr.shape[1] - 1
    bias_cube = np.empty((sample_arr.shape[0], sample_arr.shape[1], len(list(data.keys()))))
    for s in slope_axis:
        bias_cube[:, :, s] = np.array(data.get(str(int(s))))
    bias_cube[0, :, :] = 0.0
    bias_cube[:, 0, :] = 0.0
    weights = np.array([.25, .5, 1.0, .5, .25])
    weights = (1 / np.sum(weights)) * weights
    bias_cube = scipy.ndimage.convolve1d(bias_cube, weights, axis=2)

    # ---- :685-691 ----
    TC_array = TC0.copy()
    TC_array[ocean_mask == 1] = 0
    TC_nodata = np.nonzero(TC_array > 100)
    assert len(TC_nodata[0]) <= 10

    # ---- :764-781 ----
    k_s = 3
    TC_array = pdt.apply_convolve_filter(TC_array, np.ones((k_s, k_s)), normalize_kernel=True)
    TC_smooth = TC_array
    TC_array[TC_array > 100] = 0
    TC_smooth = TC_smooth.copy()
    TC_array = TC_array.astype(int)
    Theight_array = Th0.copy()
    Theight_array[Theight_array > max_treeheight] = max_treeheight
    Theight_array = pdt.apply_convolve_filter(Theight_array, np.ones((k_s, k_s)), normalize_kernel=True)
    Theight_array = Theight_array.astype(int)

    # ---- :783-794 ----
    weights = [1., 0.8, 0.6, 0.4, 0.2]
    kernel = bop.build_kernel(weights, dtype=np.float32, normalize=True)
    Zarr_smooth9 = pdt.apply_convolve_filter(Zarr, kernel)
    slope_raw = pdt.slope_D2(Zarr_smooth9, dx, dy, unit='none')
    slope_veg_array = pdt.slope_D2(Zarr_smooth9, dx, dy, unit='degree')
    slope_uncapped = slope_veg_array.copy()
    slope_veg_array[slope_veg_array > 19] = 19
    slope_veg_array = slope_veg_array.astype(int)

    # ---- :796-805 ----
    npts = TC_array.size
    TC_vec = TC_array.reshape(npts,)
    Theight_vec = Theight_array.reshape(npts,)
    slope_veg_vec = slope_veg_array.reshape(npts,)
    bias_vec = [bias_cube[TC, TH, S] for TC, TH, S in zip(TC_vec, Theight_vec, slope_veg_vec)]
    bias_vec = np.asarray(bias_vec)
    bias_array = bias_vec.reshape(TC_array.shape)
    bias_raw = bias_array.copy()

    # ---- run_tile :361-364 ----
    bias_array[watermaskTX] = 0.0
    dem_array = Zarr - bias_array

    # ---- '2Dtable' branch of compute_veg_bias_2D (:580-590, :610-629) on a synthetic slope x treecover table ----
    bias_table = np.round(rng.normal(2.0, 1.0, (51, 101)), 4)
    TC2 = TC0.copy()
    TC2[TC2 > 100] = 0
    TC2 = TC2.astype(int)
    slope2 = slope_uncapped.copy()
    TC2_vec = TC2.reshape(npts,)
    slope2_vec = slope2.reshape(npts,)
    slope2_vec[slope2_vec > 50] = 50
    bias2 = np.asarray([bias_table[S, TC] for S, TC in zip(slope2_vec, TC2_vec)]).reshape(ny, nx)

    import json
    out = dict(Zarr=Zarr, TC=TC0, Theight=Th0, edm=edm, dx=dx, dy=dy, table_json=np.array(json.dumps(data)),
               bias_cube=bias_cube, TC_idx=TC_array.astype(np.int32), TC_smooth=TC_smooth,
               Th_idx=Theight_array.astype(np.int32), slope_raw=slope_raw, slope_uncapped=slope_uncapped.astype(np.int8),
               slope_idx=slope_veg_array.astype(np.int8), bias_raw=bias_raw, bias=bias_array, dem_out=dem_array,
               max_treeheight=max_treeheight, bias_table=bias_table, bias2=bias2, slope2=slope2.astype(np.int8))
    path = os.path.join(HERE, "veg_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;",
          "TC idx range", TC_array.min(), TC_array.max(), "Th", Theight_array.min(), Theight_array.max(),
          "slope", slope_uncapped.min(), slope_uncapped.max(), "bias", bias_array.min(), bias_array.max(),
          "dem dtype", dem_array.dtype)


if __name__ == "__main__":
    main()
