"""Generates tests/golden/oracle_hashes.json and the micro-DEM fixtures (.npz).

The reference repository holds no golden vectors, so these pin OUR oracle on the
seeded generators: a change in either the generator or the oracle shows up as a hash mismatch.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def cases():
    from pydrodem_b200 import synth
    yield "c1_like_601_wbt", synth.make_dem_numpy(601, 601, 1001, nodata_frame=True), 0, 0
    yield "c1_like_601_taudem", synth.make_dem_numpy(601, 601, 1001, nodata_frame=True), 1, 1
    yield "c2_like_777x500_taudem", synth.make_dem_numpy(777, 500, 1002), 1, 1
    yield "noise_levels_200x300_wbt", synth.random_dem(200, 300, 7, "noise", nodata_holes=3, levels=12), 0, 0
    yield "bowl_257x129_taudem", synth.random_dem(257, 129, 8, "bowl", nodata_holes=1, levels=50), 1, 1


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    from oracle import hydro_oracle as O
    out = {}
    for name, z, mode, table in cases():
        f, d, s, a = O.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=mode, table=table)
        out[name] = {"shape": list(z.shape), "seed_mode": mode, "table": table, "input": digest(z), "fill": digest(f),
                     "dir": digest(d), "slope": digest(s), "acc": digest(a)}
    with open(os.path.join(HERE, "oracle_hashes.json"), "w") as fh:
        json.dump(out, fh, indent=1, sort_keys=True)
    # one small case committed in full so the GPU box can check values, not only hashes
    z = next(c for c in cases() if c[0] == "noise_levels_200x300_wbt")[1]
    f, d, s, a = O.fill_d8_accum(z, dx=30.0, dy=30.0, seed_mode=0, table=0)
    np.savez_compressed(os.path.join(HERE, "noise_levels_200x300_wbt.npz"), z=z, fill=f, dir=d, slope=s, acc=a)
    print("wrote", len(out), "golden cases")


if __name__ == "__main__":
    main()
