"""TEST INFRASTRUCTURE -- CPU restatement of the ORDER-INDEPENDENT breach formulation planned for the device
(DESIGN.md 8, item 5): every pit is breached along the flow path of the FILLED surface (fill -> seeds lowered by one
ulp -> D8 -> flat resolution, all from hydro_oracle), lengths and depths measured on the original surface, cuts combined with min().
Not used by the product yet (the shipped breach is csrc/breach.cu, pinned by oracle/breach_oracle.py); it exists so
that the next step -- a breach kernel of one thread per pit -- has a checked specification to be compared against.
Only tests may import it."""
import numpy as np

from . import breach_oracle as bo
from . import hydro_oracle as ho

DR = (-1, 0, 1, 1, 1, 0, -1, -1)          # WhiteboxTools pointer table: NE, E, SE, S, SW, W, NW, N = 1, 2, 4, ... 128
DC = (1, 1, 1, 0, -1, -1, -1, 0)
K_OF_CODE = {1 << k: k for k in range(8)}


def pits_of(w, seed, nodata):
    """non-seed valid cells without a lower valid neighbour, in raster order"""
    ny, nx = w.shape
    out = []
    for r in range(ny):
        for c in range(nx):
            v = w[r, c]
            if seed[r, c] or v == nodata or v != v:
                continue
            low = False
            for k in range(8):
                rr, cc = r + DR[k], c + DC[k]
                if 0 <= rr < ny and 0 <= cc < nx:
                    zn = w[rr, cc]
                    if zn != nodata and zn == zn and zn < v:
                        low = True
                        break
            if not low:
                out.append((r, c))
    return out


def pointers(w, nodata):
    """-> (seed mask, filled surface, pointer raster).  WhiteboxTools' pointer raster leaves 0 on a flat whose only way
    out is a seed on its own level (an edge cell has nowhere to point).  For the breach those flats must drain THROUGH
    the seed, so the pointers are taken from the filled surface with every seed lowered by one float32 ulp: the
    ordinary flat resolution then sees the seed as the flat's outlet, and no other pass is needed."""
    seed = ho.seed_mask(w, float(nodata), ho.SEED_WBT).astype(bool)
    filled = ho.fill(w, float(nodata), ho.SEED_WBT)
    lowered = filled.copy()
    lowered[seed] = np.nextafter(filled[seed], np.float32(-np.inf))
    d = ho.resolve_flats(lowered, ho.d8(lowered, float(nodata), table=ho.TABLE_WBT, want_slope=False), float(nodata),
                         ho.TABLE_WBT)
    return seed, filled, d


def breach_flowpath(z, nodata=-9999.0, max_length=0, max_depth=None, fill_pits=True, do_fill=True, order=None):
    """-> (surface, stats).  `order`: a permutation of the pit list (the result must not depend on it)."""
    z = np.asarray(z, np.float32)
    nodata = np.float32(nodata)
    w = bo.fill_single_cell_pits(z, nodata) if fill_pits else z.copy()
    seed, filled, d = pointers(w, nodata)
    orig = w.copy()
    pits = pits_of(orig, seed, nodata)
    if order is not None:
        pits = [pits[i] for i in order]
    stats = dict(pits=len(pits), breached=0, left=0, cut=0)
    ny, nx = w.shape
    for (r, c) in pits:
        zt = orig[r, c]
        q, steps, depth, ok, path = (r, c), 0, 0.0, True, []
        while True:
            if seed[q]:                       # a seed: the path leaves the raster here
                break
            k = K_OF_CODE.get(int(d[q]))
            assert k is not None, "every non-seed cell of the filled surface must have a pointer"
            q = (q[0] + DR[k], q[1] + DC[k])
            steps += 1
            assert 0 <= q[0] < ny and 0 <= q[1] < nx and steps <= ny * nx, "pointer forest must lead to a seed"
            if max_length and steps > max_length:
                ok = False
                break
            if orig[q] < zt:
                break
            depth = max(depth, float(orig[q]) - float(zt))
            if max_depth is not None and max_depth > 0 and depth > max_depth:
                ok = False
                break
            if orig[q] > zt:
                path.append(q)
        if ok:
            stats["breached"] += 1
            for q in path:
                if w[q] > zt:
                    w[q] = zt
        else:
            stats["left"] += 1
    stats["cut"] = int(np.count_nonzero(w < orig))      # cells lowered (counted once: independent of the pit order)
    if do_fill:
        w = ho.fill(w, float(nodata), ho.SEED_WBT)
    return w, stats
