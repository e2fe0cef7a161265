"""TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this; the product may not).

CPU restatement of WhiteboxTools' FillMissingData as pydrodem calls it (`wbt.fill_missing_data(i, output, filter=11 | 17,
weight=2)`, /root/reference/tools/water.py:1396,1774; models/burn_rivers.py:1283,1582; models/dem_noise_removal.py:739).
WhiteboxTools is a third-party binary that is not in /root/reference and not installed here; its version is not pinned by
the reference (README.md names the tool only).  The algorithm is restated from the tool's documentation -- "the sample
points are the valid cells on the edge of NoData areas; a NoData cell receives the inverse-distance-weighted value of the
sample points within the filter's reach" -- with these choices written down because a binary could make them differently:
an edge cell is a valid cell with a nodata 8-neighbour; reach = filter / 2 cells, Euclidean; weights 1 / d^weight in
float64; a cell with no sample in reach stays nodata.  PARITY UNPINNED (no upstream vector, no upstream binary)."""
import numpy as np


def fill_missing_data(z, nodata, filter=11, weight=2.0):
    z = np.asarray(z, np.float32)
    ny, nx = z.shape
    nd = (z == np.float32(nodata)) | np.isnan(z)
    p = np.pad(nd, 1, mode="constant")
    near = np.zeros_like(nd)
    for dr in range(3):
        for dc in range(3):
            near |= p[dr:dr + ny, dc:dc + nx]
    edge = near & ~nd
    out = z.copy()
    half = int(filter) // 2
    radius = float(filter) / 2.0
    for r, c in zip(*np.nonzero(nd)):
        r0, r1, c0, c1 = max(0, r - half), min(ny, r + half + 1), max(0, c - half), min(nx, c + half + 1)
        er, ec = np.nonzero(edge[r0:r1, c0:c1])
        if len(er) == 0:
            continue
        d = np.sqrt(((er + r0 - r) ** 2 + (ec + c0 - c) ** 2).astype(np.float64))
        keep = d <= radius
        if not keep.any():
            continue
        # row-major accumulation order, like the kernel's window walk
        w = 1.0 / (d[keep] * d[keep] if float(weight) == 2.0 else np.power(d[keep], float(weight)))
        vals = z[er[keep] + r0, ec[keep] + c0].astype(np.float64)
        sw = 0.0
        sz = 0.0
        for wi, vi in zip(w, vals):
            sw += wi
            sz += wi * vi
        out[r, c] = np.float32(sz / sw)
    return out
