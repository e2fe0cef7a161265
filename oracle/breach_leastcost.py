"""TEST INFRASTRUCTURE -- CPU specification of the least-cost breach behind
    wbt.breach_depressions_least_cost(dem, output, dist, max_cost=, min_dist=, flat_increment=, fill=False)
as dep.breach_pits(method='LC') and dep.combined_solution(carve_method='LC') call it
(/root/reference/tools/depressions.py:1370-1374, :996-1000).  Only tests may import it.

WhiteboxTools is not in /root/reference (un-vendored, version unpinned): this restates the published method (Lindsay &
Dhun 2015, "Modelling surface drainage patterns in altered landscapes using LiDAR", IJGIS 29; the tool's documentation:
a least-cost path search from every pit inside a (2 dist + 1)^2 window, cost = the elevation that has to be removed,
breaches dearer than max_cost are not made, min_dist prefers the shortest breach, flat_increment is the gradient of the
channel) -- PARITY UNPINNED.  The tool breaches its pits one after the other; here every pit is searched on the INPUT
surface and the cuts are combined with min(), which makes the result independent of any processing order (the property
the device kernel needs: one CTA per pit, all at once).  The exact rules, shared with csrc/leastcost.cu:

  valid(c)   w[c] != nodata and not NaN
  edge(c)    valid, and on the raster frame or 8-adjacent to a cell that is not valid: water that gets there is gone
  pit(p)     valid, not edge, no 8-neighbour lower, no 8-neighbour of the same elevation with a smaller raster index
             (one representative per flat: once it drains, the rest of its flat has an outlet)
  cut(c)     min(floor(max(0, float32(w[c] - w[p])) * 1024), 2^30): what cell c must be lowered by, in 1/1024 units
             (integers: path sums are exact, so the least cost does not depend on the order of the relaxation)
  target(c)  c != p and (w[c] < w[p] or edge(c))
  label(c)   the least (cost, steps) -- (steps, cost) with min_dist -- over the 8-connected paths p = c0, c1 .. ck = c of
             valid window cells whose inner cells are not targets, cost = cut(c1) + .. + cut(ck) <= max_cost * 1024
  breach     toward the target with the least label (ties: smallest window index); none reachable: the pit is left
  path       followed back from the target: the predecessor is the first neighbour (3 x 3 row-major) n with
             label(n) + cut(cur) == label(cur)
  channel    cell c_i takes min(current, float32(w[p] - i * delta)) for i = 1 .. k - 1, in float64; delta = min(eps,
             (w[p] - w[target]) / k) when the target is lower than the pit, else delta = eps and the target (an edge
             cell no lower than the pit) is cut as well
"""
import heapq

import numpy as np

N8 = ((-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1))
COST_CAP = 1 << 40
CUT_CAP = 1 << 30


def masks(w, nodata):
    """-> (valid, edge) boolean rasters"""
    w = np.asarray(w, np.float32)
    valid = (w != np.float32(nodata)) & ~np.isnan(w)
    pad = np.zeros((w.shape[0] + 2, w.shape[1] + 2), bool)
    pad[1:-1, 1:-1] = valid
    allv = np.ones_like(valid)
    for dr, dc in N8:
        allv &= pad[1 + dr:1 + dr + w.shape[0], 1 + dc:1 + dc + w.shape[1]]
    return valid, valid & ~allv


def pits_of(w, valid, edge):
    ny, nx = w.shape
    out = []
    for r in range(1, ny - 1):
        for c in range(1, nx - 1):
            if not valid[r, c] or edge[r, c]:
                continue
            v = w[r, c]
            ok = True
            for k, (dr, dc) in enumerate(N8):
                zn = w[r + dr, c + dc]
                if zn < v or (zn == v and k < 4):   # neighbours 0..3 of the row-major 3 x 3 scan have the smaller index
                    ok = False
                    break
            if ok:
                out.append((r, c))
    return out


def _key(cost, steps, min_dist):
    return (steps, cost) if min_dist else (cost, steps)


def search(w, valid, edge, pr, pc, dist, maxcost_u, min_dist):
    """labels of the window cells that matter -> (target (r, c) or None, labels dict {(r, c): (cost, steps)})"""
    ny, nx = w.shape
    zp = w[pr, pc]

    def cut(r, c):
        d = np.float32(w[r, c]) - np.float32(zp)
        if not d > 0:
            return 0
        u = np.floor(np.float32(d) * np.float32(1024.0))
        return int(min(u, CUT_CAP))

    def is_target(r, c):
        return (r, c) != (pr, pc) and (w[r, c] < zp or edge[r, c])

    lab = {(pr, pc): (0, 0)}
    heap = [(_key(0, 0, min_dist), pr, pc)]
    best = None
    done = set()
    while heap:
        k, r, c = heapq.heappop(heap)
        if (r, c) in done:
            continue
        done.add((r, c))
        if best is not None and k > best[0]:
            break
        if is_target(r, c):
            wi = (r - pr + dist) * (2 * dist + 1) + (c - pc + dist)
            if best is None or (k, wi) < best[:2]:
                best = (k, wi, r, c)
            continue                      # a target is an end point, never an inner cell
        cost, steps = lab[(r, c)]
        for dr, dc in N8:
            rr, cc = r + dr, c + dc
            if abs(rr - pr) > dist or abs(cc - pc) > dist or not (0 <= rr < ny and 0 <= cc < nx) or not valid[rr, cc]:
                continue
            if (rr, cc) == (pr, pc):
                continue
            nc, ns = cost + cut(rr, cc), steps + 1
            if nc > maxcost_u:
                continue
            old = lab.get((rr, cc))
            if old is None or _key(nc, ns, min_dist) < _key(old[0], old[1], min_dist):
                lab[(rr, cc)] = (nc, ns)
                heapq.heappush(heap, (_key(nc, ns, min_dist), rr, cc))
    if best is None:
        return None, lab, cut, is_target
    return (best[2], best[3]), lab, cut, is_target


def breach_least_cost(w, nodata=-9999.0, dist=50, max_cost=None, min_dist=False, flat_increment=None, return_stats=False):
    """the cut surface (fill it afterwards when the tool's fill=True is wanted)"""
    w = np.ascontiguousarray(w, np.float32)
    ny, nx = w.shape
    dist = int(dist)
    eps = float(flat_increment) if (flat_increment is not None and flat_increment > 0) else 0.0
    maxcost_u = COST_CAP if max_cost is None else int(min(np.floor(np.float64(np.float32(max_cost)) * 1024.0), COST_CAP))
    valid, edge = masks(w, nodata)
    out = w.copy()
    pits = pits_of(w, valid, edge)
    breached = 0
    for pr, pc in pits:
        t, lab, cut, is_target = search(w, valid, edge, pr, pc, dist, maxcost_u, min_dist)
        if t is None:
            continue
        breached += 1
        zp = float(w[pr, pc])
        k = lab[t][1]
        lower = w[t] < w[pr, pc]
        delta = min(eps, (zp - float(w[t])) / k) if lower else eps
        cur = t
        while cur != (pr, pc):
            cost, steps = lab[cur]
            if cur != t or not lower:
                ch = np.float32(zp - steps * delta)
                if ch < out[cur]:
                    out[cur] = ch
            pred = None
            for dr, dc in N8:
                n = (cur[0] + dr, cur[1] + dc)
                if n not in lab or abs(n[0] - pr) > dist or abs(n[1] - pc) > dist:
                    continue
                if n != (pr, pc) and is_target(*n):
                    continue
                ncost, nsteps = lab[n]
                if ncost + cut(*cur) == cost and nsteps + 1 == steps:
                    pred = n
                    break
            assert pred is not None, "broken label chain"
            cur = pred
    if return_stats:
        return out, dict(pits=len(pits), breached=breached, left=len(pits) - breached, cells_lowered=int((out < w).sum()))
    return out
