#!/usr/bin/env python
"""CPU model of the row-band exchange protocol, run with world_size 2 over gloo (no GPU, no CUDA library):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port P \
        tests/run_bands_gloo.py

Each rank holds one band of a small DEM and runs the SAME protocol as pydrodem_b200/bands.py + the
hc_band_* kernels — seam terminals, local minimax to the nearest open node, terminal graph reduced to a
minimum spanning forest, one all_gather of fixed-size edge records through bands.SeamComm, redundant seam
solve; and for accumulation: parked outflow + exit_of per seam cell, one all_gather, seam forest solve,
inflow added along each seam cell's path — and for labelling: local components,
min-root propagation across the seam, owner offsets, (root, id) pairs — in plain Python/numpy.  Rank 0 compares with the whole-raster
CPU oracle.  This pins the algorithm of DESIGN.md §6 and the host plumbing (SeamComm.gather/any,
exchange_halos) independently of the CUDA code.  Prints BANDS_GLOO_OK.
"""
import heapq
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NODATA = -9999.0
NB8 = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
DR = [0, -1, -1, -1, 0, 1, 1, 1]   # TauDEM codes 1..8
DC = [1, 1, 0, -1, -1, -1, 0, 1]


def fill_band_local(zh, top, bot):
    """zh: (nb+2, nx) with halo rows.  Returns (filled_local, label) for own rows and the edge dict of the
    open-node graph {(a, b): w}; nodes: 0 = outside, 1 + side*nx + c = seam cell."""
    nb, nx = zh.shape[0] - 2, zh.shape[1]
    valid = zh != NODATA
    if not top:
        valid[0] = False
    if not bot:
        valid[-1] = False
    lab = np.full((nb, nx), -1, np.int64)
    lev = np.full((nb, nx), np.inf, np.float32)
    heap, edges = [], {}

    def add_edge(a, b, w):
        if a == b:
            return
        k = (min(a, b), max(a, b))
        if k not in edges or w < edges[k]:
            edges[k] = w
    for r in range(nb):
        for c in range(nx):
            if not valid[r + 1, c]:
                continue
            seed = (r == 0 and not top) or (r == nb - 1 and not bot) or c == 0 or c == nx - 1
            for dr, dc in NB8:
                rr, cc = r + 1 + dr, c + dc
                if 0 <= cc < nx and ((rr >= 1 and rr <= nb) or (rr == 0 and top) or (rr == nb + 1 and bot)):
                    if zh[rr, cc] == NODATA:
                        seed = True
            seam = (top and r == 0) or (bot and r == nb - 1)
            if seam:
                t = 1 + (0 if (top and r == 0) else nx) + c
                lab[r, c], lev[r, c] = t, zh[r + 1, c]
                heapq.heappush(heap, (float(zh[r + 1, c]), r, c))
                if seed:
                    add_edge(0, t, float(zh[r + 1, c]))
            elif seed:
                lab[r, c], lev[r, c] = 0, zh[r + 1, c]
                heapq.heappush(heap, (float(zh[r + 1, c]), r, c))
    done = np.zeros((nb, nx), bool)
    while heap:
        l, r, c = heapq.heappop(heap)
        if done[r, c]:
            continue
        done[r, c] = True
        for dr, dc in NB8:
            rr, cc = r + dr, c + dc
            if 0 <= rr < nb and 0 <= cc < nx and valid[rr + 1, cc] and lab[rr, cc] < 0:
                lab[rr, cc] = lab[r, c]
                lev[rr, cc] = max(np.float32(l), zh[rr + 1, cc])
                heapq.heappush(heap, (float(lev[rr, cc]), rr, cc))
    for r in range(nb):
        for c in range(nx):
            if lab[r, c] < 0:
                continue
            for dr, dc in NB8:
                rr, cc = r + dr, c + dc
                if 0 <= rr < nb and 0 <= cc < nx and lab[rr, cc] >= 0 and lab[rr, cc] != lab[r, c]:
                    add_edge(int(lab[r, c]), int(lab[rr, cc]), float(max(lev[r, c], lev[rr, cc])))
    return lev, lab, edges


def kruskal(edges):
    parent = {}

    def find(x):
        while parent.setdefault(x, x) != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    out = []
    for (a, b), w in sorted(edges.items(), key=lambda kv: kv[1]):
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
            out.append((a, b, w))
    return out


def minimax_from_zero(edge_list, nnodes):
    adj = {}
    for a, b, w in edge_list:
        adj.setdefault(a, []).append((b, w))
        adj.setdefault(b, []).append((a, w))
    lvl = np.full(nnodes, -np.inf, np.float32)   # unreachable: never raised (walled in)
    seen = np.zeros(nnodes, bool)
    heap = [(-np.inf, 0)]
    while heap:
        l, x = heapq.heappop(heap)
        if seen[x]:
            continue
        seen[x] = True
        lvl[x] = l
        for y, w in adj.get(x, ()):
            if not seen[y]:
                heapq.heappush(heap, (max(l, w), y))
    return lvl


def accum_band_local(dh, top, bot):
    """dh: (nb+2, nx) uint8 with halo rows.  Returns local counts, bout[2, nx], exit_of[2, nx], down pointers."""
    nb, nx = dh.shape[0] - 2, dh.shape[1]
    down = {}
    bexit = {}
    for r in range(nb):
        for c in range(nx):
            d = dh[r + 1, c]
            if d == 255 or d == 0:
                continue
            rr, cc = r + DR[d - 1], c + DC[d - 1]
            if not (0 <= cc < nx):
                continue
            if 0 <= rr < nb:
                if dh[rr + 1, cc] != 255:
                    down[(r, c)] = (rr, cc)
            elif (rr == -1 and top) or (rr == nb and bot):
                if dh[rr + 1, cc] != 255:
                    bexit[(r, c)] = (0 if rr == -1 else 1, cc)
    acc = np.where(dh[1:-1] != 255, 1, 0).astype(np.int64)
    indeg = np.zeros((nb, nx), np.int64)
    for t in down.values():
        indeg[t] += 1
    stack = [(r, c) for r in range(nb) for c in range(nx) if indeg[r, c] == 0 and dh[r + 1, c] != 255]
    bout = np.zeros((2, nx), np.int64)
    while stack:
        p = stack.pop()
        if p in down:
            t = down[p]
            acc[t] += acc[p]
            indeg[t] -= 1
            if indeg[t] == 0:
                stack.append(t)
        elif p in bexit:
            s, c = bexit[p]
            bout[s, c] += acc[p]
    exit_of = np.full((2, nx), -1, np.int64)
    for side, r in ((0, 0), (1, nb - 1)):
        if (side == 0 and not top) or (side == 1 and not bot):
            continue
        for c in range(nx):
            if dh[r + 1, c] == 255:
                continue
            p = (r, c)
            while p in down:
                p = down[p]
            if p in bexit:
                s, cc = bexit[p]
                exit_of[side, c] = (s << 30) | cc
    return acc, bout, exit_of, down


def main():
    import torch
    import torch.distributed as dist
    from pydrodem_b200 import bands
    from oracle import hydro_oracle as O
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = bands.SeamComm()
    assert comm.world == world
    ok = True
    for seed, (ny, nx), levels in ((1, (24, 17), None), (2, (30, 22), 6), (3, (16, 40), 3)):
        rng = np.random.default_rng(seed)
        z = rng.random((ny, nx)).astype(np.float32)
        if levels:
            z = (np.floor(z * levels) / levels).astype(np.float32)
        z[5:8, 3:6] = NODATA
        rows = bands.split_rows(ny, world)
        r0 = sum(rows[:rank])
        nb = rows[rank]
        top, bot = rank > 0, rank < world - 1
        zh = torch.full((nb + 2, nx), NODATA, dtype=torch.float32)
        zh[1:-1] = torch.from_numpy(z[r0:r0 + nb])
        bands.exchange_halos([zh], rank, world, comm)
        if top:
            assert np.array_equal(zh[0].numpy(), z[r0 - 1])
        if bot:
            assert np.array_equal(zh[-1].numpy(), z[r0 + nb])
        # ---- fill ----
        lev, lab, edges = fill_band_local(zh.numpy().copy(), top, bot)
        msf = kruskal(edges)
        gbase = rank * 2 * nx
        rec = []
        for a, b, w in msf:
            rec.append((gbase + a if a else 0, gbase + b if b else 0, w))
        if bot:
            for c in range(nx):
                if zh[nb, c] == NODATA:
                    continue
                for d in (-1, 0, 1):
                    cc = c + d
                    if 0 <= cc < nx and zh[nb + 1, cc] != NODATA:
                        rec.append((gbase + 1 + nx + c, (rank + 1) * 2 * nx + 1 + cc, float(max(zh[nb, c], zh[nb + 1, cc]))))
        cap = 5 * nx + 16
        assert len(rec) <= cap
        buf = torch.zeros((1, cap, 3), dtype=torch.int32)
        for i, (a, b, w) in enumerate(rec):
            buf[0, i, 0], buf[0, i, 1] = a, b
            buf[0, i, 2] = int(np.float32(w).view(np.int32))
        alle = comm.gather(buf).numpy()
        elist = [(int(a), int(b), float(np.int32(w).view(np.float32))) for band in alle for a, b, w in band if a != b]
        tl = minimax_from_zero(elist, 1 + world * 2 * nx)
        out = z[r0:r0 + nb].copy()
        for r in range(nb):
            for c in range(nx):
                if lab[r, c] >= 0:
                    t = lab[r, c]
                    out[r, c] = max(lev[r, c], tl[gbase + t] if t else -np.inf)
        want_f = O.fill(z, seed_mode=O.SEED_TAUDEM)
        if not np.array_equal(out, want_f[r0:r0 + nb]):
            ok = False
            print(f"rank {rank} seed {seed}: banded fill model differs at {(out != want_f[r0:r0+nb]).sum()} cells", flush=True)
        # ---- accumulation on the oracle's directions ----
        d_all, _ = O.d8(want_f, dx=1.0, dy=1.0, table=O.TABLE_TAUDEM)
        d_all = O.resolve_flats(want_f, d_all, table=O.TABLE_TAUDEM)
        dh = torch.full((nb + 2, nx), 255, dtype=torch.uint8)
        dh[1:-1] = torch.from_numpy(d_all[r0:r0 + nb])
        bands.exchange_halos([dh], rank, world, comm)
        acc, bout, exit_of, down = accum_band_local(dh.numpy(), top, bot)
        rec = torch.zeros((1, 2, 2, nx), dtype=torch.int64)
        rec[0, :, 0] = torch.from_numpy(exit_of)
        rec[0, :, 1] = torch.from_numpy(bout)
        alls = comm.gather(rec).numpy()
        M = world * 2 * nx
        X = np.zeros(M, np.int64)
        target = np.full(M, -1, np.int64)
        for g in range(M):
            band, side, c = g // (2 * nx), (g // nx) & 1, g % nx
            nbb = band + 1 if side else band - 1
            if 0 <= nbb < world:
                X[g] = alls[nbb, 0 if side else 1, 1, c]
            ex = alls[band, side, 0, c]
            if ex >= 0:
                via_bottom, col = (ex >> 30) & 1, ex & 0x3FFFFFFF
                target[g] = ((band + 1) * 2) * nx + col if via_bottom else ((band - 1) * 2 + 1) * nx + col
        indeg = np.zeros(M, np.int64)
        for g in range(M):
            if target[g] >= 0:
                indeg[target[g]] += 1
        stack = [g for g in range(M) if indeg[g] == 0]
        while stack:
            g = stack.pop()
            t = target[g]
            if t >= 0:
                X[t] += X[g]
                indeg[t] -= 1
                if indeg[t] == 0:
                    stack.append(t)
        for side, r in ((0, 0), (1, nb - 1)):
            if (side == 0 and not top) or (side == 1 and not bot):
                continue
            for c in range(nx):
                x = X[rank * 2 * nx + side * nx + c]
                if not x or dh[r + 1, c] == 255:
                    continue
                p = (r, c)
                acc[p] += x
                while p in down:
                    p = down[p]
                    acc[p] += x
        want_a = O.accum(d_all, table=O.TABLE_TAUDEM).astype(np.int64)
        if not np.array_equal(acc, want_a[r0:r0 + nb]):
            ok = False
            print(f"rank {rank} seed {seed}: banded accumulation model differs at {(acc != want_a[r0:r0+nb]).sum()} cells", flush=True)
    # ---- labelling: local components, min-root propagation across the seam, owner-offset ranking ----
    from scipy import ndimage
    for seed, (ny, nx), p, conn in ((5, (26, 19), 0.55, 8), (6, (31, 24), 0.6, 4), (7, (12, 33), 0.62, 8)):
        rng = np.random.default_rng(seed)
        m = (rng.random((ny, nx)) < p)
        rows = bands.split_rows(ny, world)
        r0 = sum(rows[:rank])
        nb = rows[rank]
        st = np.ones((3, 3), int) if conn == 8 else np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]])
        loc, nloc = ndimage.label(m[r0:r0 + nb], structure=st)
        # global root of a local component = global index of its first cell (raster order)
        first = {}
        flat = loc.ravel()
        for i in np.flatnonzero(flat):
            first.setdefault(int(flat[i]), r0 * nx + int(i))
        G = dict(first)
        NONE = 2 ** 31 - 1

        def seams():
            t = torch.full((1, 2, nx), NONE, dtype=torch.int64)
            for side, r in ((0, 0), (1, nb - 1)):
                for c in range(nx):
                    if loc[r, c]:
                        t[0, side, c] = G[int(loc[r, c])]
            return t
        rounds = 0
        while True:
            alls = comm.gather(seams()).numpy()
            changed = False
            for side, r, ob in ((0, 0, rank - 1), (1, nb - 1, rank + 1)):
                if ob < 0 or ob >= world:
                    continue
                row = alls[ob, 0 if side else 1]
                for c in range(nx):
                    if not loc[r, c]:
                        continue
                    cand = [row[cc] for cc in ((c - 1, c, c + 1) if conn == 8 else (c,)) if 0 <= cc < nx]
                    mn = min(cand)
                    if mn < G[int(loc[r, c])]:
                        G[int(loc[r, c])] = int(mn)
                        changed = True
            rounds += 1
            if not comm.any(changed, torch.device("cpu")):
                break
        owned = sorted(first[k] for k in first if G[k] == first[k])
        counts = comm.gather(torch.tensor([[len(owned)]], dtype=torch.int64)).numpy().ravel()
        offset = int(counts[:rank].sum())
        myid = {g: offset + i + 1 for i, g in enumerate(owned)}
        pairs = torch.zeros((1, 2 * nx, 2), dtype=torch.int64)
        k = 0
        for side, r in ((0, 0), (1, nb - 1)):
            for c in range(nx):
                if loc[r, c] and G[int(loc[r, c])] in myid:
                    pairs[0, k, 0], pairs[0, k, 1] = G[int(loc[r, c])], myid[G[int(loc[r, c])]]
                    k += 1
        allp = comm.gather(pairs).numpy().reshape(-1, 2)
        lut = {int(a): int(b) for a, b in allp if b > 0}
        lut.update(myid)
        out = np.zeros((nb, nx), np.int64)
        for k_, g in G.items():
            out[loc == k_] = lut[g]
        want, nwant = ndimage.label(m, structure=st)
        if not np.array_equal(out, want[r0:r0 + nb]) or int(counts.sum()) != nwant:
            ok = False
            print(f"rank {rank} seed {seed}: banded label model differs (rounds {rounds})", flush=True)
    # ---- local-window functions over bands: apply_rank's halo exchange (CPU stand-in functions) ----
    rng = np.random.default_rng(99)
    ny, nx = 61, 23
    full = rng.random((ny, nx)).astype(np.float32)
    fullm = (rng.random((ny, nx)) < 0.08).astype(np.uint8)
    rows = bands.split_rows(ny, world)
    r0 = sum(rows[:rank])

    def fn(a, b):
        return (torch.from_numpy(ndimage.minimum_filter(a.numpy(), size=15)),
                torch.from_numpy(ndimage.maximum_filter(b.numpy(), size=7)))
    got = bands.apply_rank(fn, [torch.from_numpy(full[r0:r0 + rows[rank]].copy()),
                                torch.from_numpy(fullm[r0:r0 + rows[rank]].copy())], 7, comm)
    want = (ndimage.minimum_filter(full, size=15), ndimage.maximum_filter(fullm, size=7))
    for g, w in zip(got, want):
        if not np.array_equal(g.numpy(), w[r0:r0 + rows[rank]]):
            ok = False
            print(f"rank {rank}: apply_rank differs from the whole-raster filter", flush=True)
    anybad = comm.any(not ok, torch.device("cpu"))
    assert comm.any(rank == 1, torch.device("cpu")) is True and comm.any(False, torch.device("cpu")) is False
    if rank == 0:
        print("BANDS_GLOO_FAIL" if anybad else "BANDS_GLOO_OK", flush=True)
    dist.destroy_process_group()
    return 1 if anybad else 0


if __name__ == "__main__":
    sys.exit(main())
