"""TEST INFRASTRUCTURE -- an independent, slow restatement of the breach specification in
pydrodem_b200/csrc/breach.cu (Lindsay 2016 as used by wbt.breach_depressions, tools/depressions.py:1376-1380),
written with heapq and plain loops for small rasters.  PARITY UNPINNED: WhiteboxTools is not available, so this
pins the product to the written specification (seed rule, heap order, neighbour order, path rule), not to the binary.
Only tests may import it."""
import heapq

import numpy as np

DR = (-1, 0, 1, 1, 1, 0, -1, -1)
DC = (1, 1, 1, 0, -1, -1, -1, 0)


def _nod(v, nodata):
    return v == nodata or v != v


def seed_mask(z, nodata):
    ny, nx = z.shape
    seed = np.zeros((ny, nx), bool)
    ext = np.zeros((ny, nx), bool)
    stack = []
    for r in range(ny):
        for c in range(nx):
            if r in (0, ny - 1) or c in (0, nx - 1):
                if _nod(z[r, c], nodata):
                    if not ext[r, c]:
                        ext[r, c] = True
                        stack.append((r, c))
                else:
                    seed[r, c] = True
    while stack:
        r, c = stack.pop()
        for k in range(8):
            rr, cc = r + DR[k], c + DC[k]
            if 0 <= rr < ny and 0 <= cc < nx:
                if _nod(z[rr, cc], nodata):
                    if not ext[rr, cc]:
                        ext[rr, cc] = True
                        stack.append((rr, cc))
                else:
                    seed[rr, cc] = True
    return seed


def fill_single_cell_pits(z, nodata):
    ny, nx = z.shape
    out = z.copy()
    for r in range(1, ny - 1):
        for c in range(1, nx - 1):
            v = z[r, c]
            if _nod(v, nodata):
                continue
            nb = [z[r + DR[k], c + DC[k]] for k in range(8)]
            if all((not _nod(x, nodata)) and x > v for x in nb):
                out[r, c] = min(nb)
    return out


def priority_flood_fill(w, seed, nodata):
    ny, nx = w.shape
    w = w.copy()
    vis = seed.copy()
    heap = []
    seq = 0
    for r in range(ny):
        for c in range(nx):
            if seed[r, c]:
                heap.append((w[r, c], seq, r, c))
                seq += 1
    heapq.heapify(heap)
    while heap:
        zc, _, r, c = heapq.heappop(heap)
        for k in range(8):
            rr, cc = r + DR[k], c + DC[k]
            if not (0 <= rr < ny and 0 <= cc < nx) or vis[rr, cc] or _nod(w[rr, cc], nodata):
                continue
            vis[rr, cc] = True
            if w[rr, cc] < zc:
                w[rr, cc] = zc
            heapq.heappush(heap, (w[rr, cc], seq, rr, cc))
            seq += 1
    return w


def breach_depressions(z, nodata=-9999.0, max_length=0, max_depth=None, fill_pits=True, do_fill=True):
    z = np.asarray(z, np.float32)
    nodata = np.float32(nodata)
    ny, nx = z.shape
    w = fill_single_cell_pits(z, nodata) if fill_pits else z.copy()
    seed = seed_mask(w, nodata)
    vis = seed.copy()
    back = {}
    heap = []
    seq = 0
    for r in range(ny):
        for c in range(nx):
            if seed[r, c]:
                heap.append((w[r, c], seq, r, c))
                seq += 1
    heapq.heapify(heap)
    stats = dict(pits=0, breached=0, left=0, cut=0)
    while heap:
        _, _, r, c = heapq.heappop(heap)
        if not seed[r, c]:
            zt = w[r, c]
            pit = True
            for k in range(8):
                rr, cc = r + DR[k], c + DC[k]
                if 0 <= rr < ny and 0 <= cc < nx and not _nod(w[rr, cc], nodata) and w[rr, cc] < zt:
                    pit = False
            if pit:
                stats["pits"] += 1
                q, length, depth, ok, path = (r, c), 0, 0.0, True, []
                while q in back:
                    q = back[q]
                    length += 1
                    if max_length and max_length > 0 and length > max_length:
                        ok = False
                        break
                    if w[q] < zt:
                        break
                    depth = max(depth, float(w[q]) - float(zt))
                    if max_depth is not None and max_depth > 0 and np.isfinite(max_depth) and depth > max_depth:
                        ok = False
                        break
                    if w[q] > zt:
                        path.append(q)
                if ok:
                    stats["breached"] += 1
                    stats["cut"] += len(path)
                    for q in path:
                        w[q] = zt
                else:
                    stats["left"] += 1
        for k in range(8):
            rr, cc = r + DR[k], c + DC[k]
            if not (0 <= rr < ny and 0 <= cc < nx) or vis[rr, cc] or _nod(w[rr, cc], nodata):
                continue
            vis[rr, cc] = True
            back[(rr, cc)] = (r, c)
            heapq.heappush(heap, (w[rr, cc], seq, rr, cc))
            seq += 1
    if do_fill:
        w = priority_flood_fill(w, seed, nodata)
    return w, stats
