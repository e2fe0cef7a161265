"""Golden vectors for the pit catalog / solution selection path, produced by the REFERENCE ITSELF.

tools/depressions.py of /root/reference imports here once osgeo / geopandas are stubbed (it only needs
them for file I/O) and `np.int` is restored (the reference predates numpy 1.24).  This script builds two
seeded scenarios — DEM, filled surface, a synthetic "carved + back-filled" surface with trenches cut
through pit rims, water / wall / sink masks — runs the reference's build_pit_catalog, select_solution,
apply_solution and fill_shallow_pits on them, and stores inputs and outputs in
tests/golden/catalog_reference.npz.  /root/reference does not exist on the GPU box, hence the committed
file.  Run from the repo root:  python tests/golden/make_catalog_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
NODATA = -9999.0


def import_reference():
    np.int = int  # removed in numpy 1.24; the reference uses it (tools/depressions.py:352)

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith('__'):
                raise AttributeError(k)
            m = _Any(self.__name__ + '.' + k)
            setattr(self, k, m)
            return m

        def __call__(self, *a, **k):
            return None
    for name in ['osgeo', 'osgeo.gdal', 'osgeo.gdal_array', 'osgeo.ogr', 'osgeo.osr', 'osgeo.gdalconst', 'geopandas',
                 'shapely', 'shapely.geometry', 'shapely.ops', 'fiona', 'rasterio', 'pyproj', 'boto3', 'whitebox_tools']:
        sys.modules[name] = _Any(name)
    sys.modules['osgeo'].gdal = sys.modules['osgeo.gdal']
    sys.modules['osgeo'].gdal_array = sys.modules['osgeo.gdal_array']
    sys.path.insert(0, REF)
    import tools.depressions as dep
    return dep


def scenario(seed, ny, nx):
    from scipy import ndimage
    from oracle import hydro_oracle as O
    from pydrodem_b200 import synth
    rng = np.random.default_rng(seed)
    # gentle rolling terrain (smoothed noise + slight tilt) with single-cell pits, so that depressions of all sizes exist
    base = ndimage.gaussian_filter(rng.standard_normal((ny, nx)), 7.0) * 60.0
    yy, xx = np.mgrid[0:ny, 0:nx]
    z = base + 0.02 * yy + 0.01 * xx + rng.standard_normal((ny, nx)) * 0.05
    pit = rng.random((ny, nx)) < 0.01
    z[pit] -= rng.uniform(0.1, 2.0, pit.sum())
    z = (np.round(z * 100.0) / 100.0).astype(np.float32)
    z[0, :] = NODATA
    z[:, 0] = NODATA
    z[ny // 3: ny // 3 + 6, nx // 2: nx // 2 + 9] = NODATA      # an interior hole
    del synth
    valid = z != NODATA
    fill = O.fill(z, seed_mode=O.SEED_WBT)
    diff_fill = np.round(fill - z, 3)
    s8 = np.ones((3, 3), int)
    # pit ids as initial_fill_carve builds them (:212): 3x3 dilation of diff > 0, then 4-connected labels, so
    # that a pit's footprint overlaps the first cells of a carve leaving it
    s4 = [[0, 1, 0], [1, 1, 1], [0, 1, 0]]
    fillID, npits = ndimage.label(ndimage.binary_dilation(diff_fill > 0, structure=s8), structure=s4)
    # synthetic breach: for ~70 % of the pits, cut a strictly descending trench from the pit's lowest cell along
    # the filled surface's drainage path (D8 with flats resolved) until ground lower than the pit bottom is
    # reached — what a breach does; a few random trenches on top give carves that miss or only graze pits
    carved = z.copy()
    d8, _ = O.d8(fill, dx=1.0, dy=1.0, table=O.TABLE_TAUDEM)
    d8 = O.resolve_flats(fill, d8, table=O.TABLE_TAUDEM)
    DR = [0, -1, -1, -1, 0, 1, 1, 1]
    DC = [1, 1, 0, -1, -1, -1, 0, 1]
    trenches = []
    for p in range(1, npits + 1):
        if rng.random() > 0.7:
            continue
        cells = np.argwhere(fillID == p)
        r, c = cells[np.argmin(z[cells[:, 0], cells[:, 1]])]
        bottom = float(z[r, c])
        lvl, path = bottom, []
        max_len = int(rng.integers(5, 120))
        for _ in range(max_len):
            k = int(d8[r, c])
            if k < 1 or k > 8:
                break
            r, c = r + DR[k - 1], c + DC[k - 1]
            if not (0 <= r < ny and 0 <= c < nx) or not valid[r, c]:
                break
            lvl -= 0.002
            if z[r, c] < lvl:
                break
            path.append((r, c, lvl))
        for (rr, cc, lv) in path:
            carved[rr, cc] = min(carved[rr, cc], np.float32(lv))
        if path:
            trenches.append(path)
    pit_cells = np.argwhere(diff_fill > 0.02)
    for _ in range(25):
        r, c = pit_cells[rng.integers(len(pit_cells))]
        ang = rng.uniform(0, 2 * np.pi)
        lvl = float(z[r, c])
        for t in range(1, int(rng.integers(6, 30))):
            rr, cc = int(round(r + t * np.sin(ang))), int(round(c + t * np.cos(ang)))
            if not (0 <= rr < ny and 0 <= cc < nx) or not valid[rr, cc]:
                break
            lvl -= 0.01
            carved[rr, cc] = min(carved[rr, cc], np.float32(lvl))
    carve = O.fill(carved, seed_mode=O.SEED_WBT)       # breach_depressions(fill_pits=True): what is left gets filled
    diff_carve = np.round(carve - z, 3)
    diff_carve[~valid] = -1                             # tools/depressions.py:241
    carveID, _ = ndimage.label(diff_carve < 0, structure=s8)
    carvefillID, _ = ndimage.label(diff_carve > 0, structure=s8)
    water = np.zeros((ny, nx), np.uint8)
    for _ in range(10):
        r, c = rng.integers(0, ny), rng.integers(0, nx)
        h, w = rng.integers(3, 25), rng.integers(3, 25)
        water[r:r + h, c:c + w] = rng.choice([1, 3, 9])
    for path in trenches[::5]:                          # some carves lie completely inside the water mask (-> 1002)
        for (rr, cc, _) in path:
            water[max(rr - 1, 0):rr + 2, max(cc - 1, 0):cc + 2] = 1
    wall = np.zeros((ny, nx), np.uint8)
    sink = np.zeros((ny, nx), np.uint8)
    sizes = np.bincount(fillID.ravel())
    small = [p for p in range(1, npits + 1) if 3 <= sizes[p] <= 30]
    if small:
        wp = small[len(small) // 2]
        wall[fillID == wp] = 1
        sp = small[len(small) // 3]
        rr, cc = np.argwhere(fillID == sp)[0]
        sink[rr, cc] = 1
    return dict(z=z, fill=fill, carve=carve, diff_fill=diff_fill.astype(np.float32), diff_carve=diff_carve.astype(np.float32),
                fillID=fillID.astype(np.int32), carveID=carveID.astype(np.int32), carvefillID=carvefillID.astype(np.int32),
                water=water, wall=wall, sink=sink)


def main():
    dep = import_reference()
    out = {}
    for tag, seed, shape in (("a", 2101, (150, 190)), ("b", 2102, (210, 160))):
        sc = scenario(seed, *shape)
        ny, nx = shape
        v = {k: np.ravel(a) for k, a in sc.items()}
        DF, fillcells, carveonly, carvecells, carvefill = dep.build_pit_catalog(
            v["z"], NODATA, v["carve"], v["diff_carve"], v["diff_fill"], v["fillID"], v["carveID"], v["carvefillID"],
            verbose_print=False, carveout=False)
        fvt = float(np.quantile(DF.volfill, 0.7))
        cvt = float(np.quantile(DF.volcarve[DF.volcarve > 0], 0.6)) if (DF.volcarve > 0).any() else 1.0
        mcd = float(np.quantile(DF.maxcarve[DF.maxcarve > 0], 0.8)) if (DF.maxcarve > 0).any() else 1.0
        for k, a in sc.items():
            out[f"{tag}_{k}"] = a
        out[f"{tag}_thresholds"] = np.array([fvt, cvt, mcd], np.float64)
        for col in ("maxfill", "volfill", "maxcarve", "volcarve", "volfillbehind", "maxfillbehind", "completecarve"):
            out[f"{tag}_cat_{col}"] = DF[col].to_numpy()
        # a few dictionary entries (the lazy views must reproduce them)
        pits = list(DF.index)
        probe = [p for p in pits if len(carvecells[p]) > 0][:4] + pits[:2]
        out[f"{tag}_probe"] = np.array(probe, np.int64)
        for p in probe:
            out[f"{tag}_fillcells_{p}"] = np.asarray(fillcells[p], np.uint32)
            out[f"{tag}_carvecells_{p}"] = np.asarray(carvecells[p], np.uint32)
            out[f"{tag}_carvefill_{p}"] = np.asarray(carvefill[p], np.uint32)
        big = [float(DF.volfill.max()) * 2, float(DF.volcarve.max()) * 2 + 1, float(DF.maxcarve.max()) * 2 + 1]
        tiny = [float(np.quantile(DF.volfill, 0.1)), float(np.quantile(DF.volcarve, 0.3)) + 1e-3, mcd * 2]
        out[f"{tag}_thresholds_sel2"] = np.array(big, np.float64)
        out[f"{tag}_thresholds_sel3"] = np.array(tiny, np.float64)
        for variant, asf, use_masks, thr in (("sel", False, True, (fvt, cvt, mcd)), ("selshallow", True, False, (fvt, cvt, mcd)),
                                             ("sel2", False, False, big), ("sel3", False, False, tiny)):
            D2 = dep.select_solution(DF.copy(), fillcells, carvecells, carveonly, v["fill"], v["carve"], v["fillID"],
                                     v["carveID"], thr[0], thr[1], thr[2], maxcarve_factor=1.25, apply_shallow_fill=asf,
                                     wall_array=sc["wall"] if use_masks else None, water_mask_array=sc["water"],
                                     sink_array=sc["sink"] if use_masks else None)
            out[f"{tag}_{variant}_solution"] = D2["solution"].to_numpy().astype(np.int64)
            D3 = D2[D2.solution < 3000]
            mod, mask = dep.apply_solution(D3, fillcells, carvecells, carvefill, v["z"], v["fill"], v["carve"])
            out[f"{tag}_{variant}_mod"] = mod
            out[f"{tag}_{variant}_mask"] = mask
            if variant == "sel":
                ex = [int(p) for p in D3.index[:3]]
                mod2, mask2 = dep.apply_solution(D3, fillcells, carvecells, carvefill, v["z"], v["fill"], v["carve"],
                                                 solution_mask_vec=mask.copy(), excludedpits=ex)
                out[f"{tag}_excluded"] = np.array(ex, np.int64)
                out[f"{tag}_sel_mod_excl"], out[f"{tag}_sel_mask_excl"] = mod2, mask2
        for sfill, use_water in ((0.3, True), (1.5, False)):
            m = dep.fill_shallow_pits(v["z"].copy(), sfill, v["fillID"], sc["diff_fill"], sc["fill"],
                                      water_vec=v["water"] if use_water else None, verbose_print=False)
            out[f"{tag}_shallow_{int(sfill*10)}_{int(use_water)}"] = m
        for variant in ("sel", "selshallow", "sel2", "sel3"):
            print(tag, variant, "pits", len(DF), {int(k): int(c) for k, c in zip(*np.unique(out[f"{tag}_{variant}_solution"], return_counts=True))})
    np.savez_compressed(os.path.join(HERE, "catalog_reference.npz"), **out)
    print("wrote", os.path.join(HERE, "catalog_reference.npz"))


if __name__ == "__main__":
    main()
