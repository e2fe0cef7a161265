"""Golden vectors for the pit catalog / solution selection at configuration C4's scale and recipe (BASELINE.json
configs[3]: flattened lakes + burned rivers + the water mask that steers the selection), produced by the REFERENCE's own
tools/depressions.py (imported from /root/reference with the GDAL stubs of make_catalog_golden.py).

Scene: synth.make_c4(1201, 1201, 1004) -> DEM + water mask (codes 1 river, 3 lake); fill = oracle; carve = the DEM with
breach-like trenches cut along the filled surface's drainage paths, then filled (what breach_depressions(fill_pits=True)
returns); pit / carve ids by scipy exactly as initial_fill_carve builds them.  Stored: z, water, carve (the rest is
recomputed by the test with the oracle and scipy) and the reference's catalog columns, solution codes and applied DEM.
Run from the repo root (takes a few minutes: the reference loops over pits in Python):
    python tests/golden/make_catalog_c4_golden.py"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
NODATA = -9999.0
N = 1201


def derive(z, carve):
    """everything initial_fill_carve derives from (DEM, carved surface): shared by this script and the test"""
    from scipy import ndimage
    from oracle import hydro_oracle as O
    valid = z != NODATA
    fill = O.fill(z, seed_mode=O.SEED_WBT)
    diff_fill = np.round(fill - z, 3)
    s8 = np.ones((3, 3), int)
    s4 = [[0, 1, 0], [1, 1, 1], [0, 1, 0]]
    fillID, npits = ndimage.label(ndimage.binary_dilation(diff_fill > 0, structure=s8), structure=s4)
    diff_carve = np.round(carve - z, 3)
    diff_carve[~valid] = -1
    carveID, _ = ndimage.label(diff_carve < 0, structure=s8)
    carvefillID, _ = ndimage.label(diff_carve > 0, structure=s8)
    return dict(fill=fill, diff_fill=diff_fill.astype(np.float32), diff_carve=diff_carve.astype(np.float32),
                fillID=fillID.astype(np.int32), carveID=carveID.astype(np.int32), carvefillID=carvefillID.astype(np.int32))


def scene():
    from oracle import hydro_oracle as O
    from pydrodem_b200 import synth
    z, water = synth.make_c4(N, N, 1004, device="cpu")
    z, water = z.numpy().copy(), water.numpy().copy()
    z[0, :] = NODATA
    z[:, 0] = NODATA
    valid = z != NODATA
    rng = np.random.default_rng(40004)
    fill = O.fill(z, seed_mode=O.SEED_WBT)
    d8, _ = O.d8(fill, dx=1.0, dy=1.0, table=O.TABLE_TAUDEM)
    d8 = O.resolve_flats(fill, d8, table=O.TABLE_TAUDEM)
    DR = [0, -1, -1, -1, 0, 1, 1, 1]
    DC = [1, 1, 0, -1, -1, -1, 0, 1]
    carved = z.copy()
    raised = np.argwhere((fill - z) > 0.3)
    # breach-like trenches from ~2500 depression cells down the filled surface's flow path
    for i in rng.choice(len(raised), size=min(2500, len(raised)), replace=False):
        r, c = raised[i]
        lvl = float(z[r, c])
        for _ in range(int(rng.integers(5, 90))):
            k = int(d8[r, c])
            if k < 1 or k > 8:
                break
            r, c = r + DR[k - 1], c + DC[k - 1]
            if not (0 <= r < N and 0 <= c < N) or not valid[r, c]:
                break
            lvl -= 0.002
            if z[r, c] < lvl:
                break
            carved[r, c] = min(carved[r, c], np.float32(lvl))
    carve = O.fill(carved, seed_mode=O.SEED_WBT)
    return z, water, carve


def main():
    from make_catalog_golden import import_reference
    dep = import_reference()
    t0 = time.time()
    z, water, carve = scene()
    d = derive(z, carve)
    v = {k: np.ravel(a) for k, a in d.items()}
    print("scene", time.time() - t0, "s; pits", int(d["fillID"].max()), "carves", int(d["carveID"].max()), "water cells", int((water > 0).sum()))
    t0 = time.time()
    DF, fillcells, carveonly, carvecells, carvefill = dep.build_pit_catalog(
        z.ravel(), NODATA, carve.ravel(), v["diff_carve"], v["diff_fill"], v["fillID"], v["carveID"], v["carvefillID"],
        verbose_print=False, carveout=False)
    print("reference build_pit_catalog", time.time() - t0, "s")
    sols, mods = {}, {}
    # the ini's thresholds scaled to this scene's small volumes, and a tight set that pushes pits into the carve / combined codes
    for tag, (fvt, cvt, mcd) in (("a", (12000.0 / 30.0, 4000.0 / 30.0, 4.0)), ("b", (15.0, 4.0, 0.8))):
        t0 = time.time()
        D2 = dep.select_solution(DF.copy(), fillcells, carvecells, carveonly, v["fill"], carve.ravel(), v["fillID"], v["carveID"],
                                 fvt, cvt, mcd, maxcarve_factor=1.25, apply_shallow_fill=False, wall_array=None,
                                 water_mask_array=water, sink_array=None)
        print("reference select_solution", tag, time.time() - t0, "s", {int(k): int(c) for k, c in zip(*np.unique(D2.solution, return_counts=True))})
        D3 = D2[D2.solution < 3000]
        mod, mask = dep.apply_solution(D3, fillcells, carvecells, carvefill, z.ravel(), v["fill"], carve.ravel())
        sols[tag] = (np.array([fvt, cvt, mcd]), D2["solution"].to_numpy().astype(np.int32), mod.astype(np.float32), mask.astype(np.int16))
    shallow = dep.fill_shallow_pits(z.ravel().copy(), 1.0, v["fillID"], d["diff_fill"], d["fill"], water_vec=water.ravel(),
                                    verbose_print=False)
    out = dict(z=z, water=water, carve=carve, shallow=shallow.astype(np.float32))
    for col in ("maxfill", "volfill", "maxcarve", "volcarve", "volfillbehind", "maxfillbehind", "completecarve"):
        out[f"cat_{col}"] = DF[col].to_numpy()
    # the applied / shallow-filled rasters differ from z in few cells: store them sparsely
    idx = np.flatnonzero(out["shallow"] != z.ravel())
    out["shallow_idx"], out["shallow_val"] = idx.astype(np.int32), out["shallow"][idx]
    del out["shallow"]
    for tag, (thr, sol, mod, mask) in sols.items():
        out[f"{tag}_thresholds"], out[f"{tag}_solution"] = thr, sol
        idx = np.flatnonzero(mod != z.ravel())
        out[f"{tag}_mod_idx"], out[f"{tag}_mod_val"] = idx.astype(np.int32), mod[idx]
        idx = np.flatnonzero(mask != 0)
        out[f"{tag}_mask_idx"], out[f"{tag}_mask_val"] = idx.astype(np.int32), mask[idx]
    cidx = np.flatnonzero(carve.ravel() != z.ravel())
    out["carve_idx"], out["carve_val"] = cidx.astype(np.int32), carve.ravel()[cidx]
    del out["carve"]
    p = os.path.join(HERE, "catalog_c4_reference.npz")
    np.savez_compressed(p, **out)
    print("wrote", p, os.path.getsize(p) / 1e6, "MB")


if __name__ == "__main__":
    main()
