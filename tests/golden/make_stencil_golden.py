"""Golden vectors for the smoothing stencils, produced by the REFERENCE ITSELF.

Unlike fill/D8/accumulation, this part of the path is pinned: `tools/derive.py` and
`tools/build_operators.py` import and run unmodified in the build container, so this script calls
them on a seeded DEM and stores their outputs.  /root/reference does not exist on the GPU box, hence
the committed .npz.  Run from the repo root:  python tests/golden/make_stencil_golden.py

The filter-selection block of CreateDTM.run_tile is a method body full of GDAL I/O, so its array
logic (models/dem_noise_removal.py:412-471) is replayed here statement by statement with numpy on
the reference-computed inputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    sys.path.insert(0, REF)
    import tools.build_operators as bop
    import tools.derive as pdt
    from pydrodem_b200 import synth

    dx, dy = 30.0, 30.0
    dem = synth.make_dem_numpy(150, 131, 1005, quantise=False)
    rng = np.random.default_rng(1005)
    dem = (dem + rng.normal(0, 1.5, dem.shape)).astype(np.float32)  # radar-like noise so that all filters fire
    fweights = [[1., 0.8, 0.6, 0.4, 0.2], [1., 0.75, 0.5, 0.25], [1., 0.7, 0.35], [1., 0.5]]  # setfweights 'DW' :1265-1270
    kernels = [bop.build_kernel(w, dtype=np.float32, normalize=True) for w in fweights]
    zs = [pdt.apply_convolve_filter(dem, k) for k in kernels]
    out = {"dem": dem, "dx": dx, "dy": dy}
    for k, z, n in zip(kernels, zs, (9, 7, 5, 3)):
        out[f"kernel{n}"] = k
        out[f"z{n}"] = z
    out["slope_d2_deg"] = pdt.slope_D2(zs[2], dx, dy)
    out["slope_d4_deg"] = pdt.slope_D4(zs[2], dx, dy)
    out["slope_d2_mag"] = pdt.slope_D2(zs[2], dx, dy, unit="none")
    out["slope_d4_mag"] = pdt.slope_D4(zs[2], dx, dy, unit="none")
    out["curv_geo"] = pdt.curvature_D2(zs[2], dx, dy)
    out["curv_lap"] = pdt.curvature_D2(zs[2], dx, dy, geometric=False)
    ones = np.ones((3, 3))
    m = (rng.random(dem.shape) < 0.05).astype(np.int8)
    out["dilate_in"] = m
    out["dilate_out"] = pdt.apply_convolve_filter(m, ones)

    # --- replay of models/dem_noise_removal.py:412-471 on the reference-computed arrays ------------
    slopebreaks, curvebreaks = [0, 2, 6, 15, 30], [1.0, 2.5]
    slope_0, curvature = out["slope_d4_deg"], out["curv_geo"]
    std_geoC = np.std(curvature)
    cbreak = [c * std_geoC for c in curvebreaks]
    peaks = np.where((curvature < -cbreak[1]))
    peakfill = np.where((curvature >= -cbreak[1]) & (curvature < -cbreak[0]))
    valleys = np.where((curvature > cbreak[1]))
    valleyfill = np.where((curvature <= cbreak[1]) & (curvature > cbreak[0]))
    row_steep, col_steep = np.where(slope_0 >= slopebreaks[-1])
    bound_rows = list(valleys[0]) + list(peaks[0]) + list(row_steep)
    bound_cols = list(valleys[1]) + list(peaks[1]) + list(col_steep)
    fill_rows = list(valleyfill[0]) + list(peakfill[0])
    fill_cols = list(valleyfill[1]) + list(peakfill[1])
    swo = rng.random(dem.shape) < 0.03
    edm = np.where(rng.random(dem.shape) < 0.04, 5, 0).astype(np.uint8)
    tx = rng.random(dem.shape) < 0.02
    artifact = (rng.random(dem.shape) < 0.01).astype(np.uint8)
    for variant, (m_swo, m_edm, m_tx, m_art) in {"plain": (None, None, None, None),
                                                 "masked": (swo, edm, tx, artifact)}.items():
        Zf = zs[0].copy()
        filter_array = 9 * np.ones((Zf.shape[0], Zf.shape[1])).astype(np.int8)
        for i, Zs_uniform in enumerate(zs[1::]):
            filtersize = int(2 * len(fweights[i + 1]) - 1)
            min_slope, max_slope = slopebreaks[i + 1], slopebreaks[i + 2]
            rows, cols = np.where((slope_0 >= min_slope) & (slope_0 <= max_slope))
            if filtersize == 3:
                rows = list(rows) + fill_rows
                cols = list(cols) + fill_cols
            filter_array[rows, cols] = filtersize
            Zf[rows, cols] = Zs_uniform[rows, cols]
        filter_array[row_steep, col_steep] = 1
        filter_array[peaks] = 2
        filter_array[valleys] = 0
        Zf[bound_rows, bound_cols] = dem[bound_rows, bound_cols]
        if m_swo is not None:
            w = np.where(m_swo)
            filter_array[w] = 3
            Zf[w] = zs[-1][w]
            DEMfill = np.where((m_edm == 5) & (filter_array < 3))
            filter_array[DEMfill] = 3
            Zf[DEMfill] = zs[-1][DEMfill]
            w = np.where(m_tx)
            filter_array[w] = 1
            Zf[w] = dem[w]
            Zf[m_art == 1] = dem[m_art == 1]
        out[f"zf_{variant}"] = Zf
        out[f"filter_{variant}"] = filter_array
        out[f"slope_final_{variant}"] = pdt.slope_D4(Zf, dx, dy)
    out["sigma"] = np.float32(std_geoC)
    out["swo"], out["edm"], out["tx"], out["artifact"] = swo, edm, tx, artifact
    np.savez_compressed(os.path.join(HERE, "stencils_reference.npz"), **out)
    print("wrote stencils_reference.npz with", len(out), "arrays; filters used:",
          {int(k): int(v) for k, v in zip(*np.unique(out["filter_plain"], return_counts=True))})


if __name__ == "__main__":
    main()
