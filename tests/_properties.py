"""Size-independent properties of the chain's outputs, as torch functions (CPU or CUDA tensors): what the full-size GPU
test checks where the oracle cannot follow (BASELINE.json's 10801^2), and what the CPU suite first checks against the
oracle at small size.  Each returns the number of violating cells (0 = the property holds)."""
import torch

TAU_DR = (0, -1, -1, -1, 0, 1, 1, 1)        # TauDEM codes 1..8 = E, NE, N, NW, W, SW, S, SE
TAU_DC = (1, 1, 0, -1, -1, -1, 0, 1)
WBT_DR = (-1, 0, 1, 1, 1, 0, -1, -1)        # WhiteboxTools codes 1, 2, 4 ... 128 = NE, E, SE, S, SW, W, NW, N
WBT_DC = (1, 1, 1, 0, -1, -1, -1, 0)


def _tables(table):
    if table == 1:
        return TAU_DR, TAU_DC, tuple(range(1, 9))
    return WBT_DR, WBT_DC, tuple(1 << k for k in range(8))


def downstream(d, table):
    """flat index of the cell each cell points to, -1 where the path ends (no code, off the raster, onto a 255 cell)"""
    ny, nx = d.shape
    dr, dc, codes = _tables(table)
    rows = torch.arange(ny, device=d.device).view(-1, 1).expand(ny, nx)
    cols = torch.arange(nx, device=d.device).view(1, -1).expand(ny, nx)
    down = torch.full((ny, nx), -1, dtype=torch.int64, device=d.device)
    for k in range(8):
        m = d == codes[k]
        rr, cc = rows + dr[k], cols + dc[k]
        inb = m & (rr >= 0) & (rr < ny) & (cc >= 0) & (cc < nx)
        tgt = (rr.clamp(0, ny - 1) * nx + cc.clamp(0, nx - 1))
        ok = inb & (d.reshape(-1)[tgt.reshape(-1)].view(ny, nx) != 255)
        down = torch.where(ok, tgt, down)
    return down


def accumulation_violations(d, acc, table):
    """acc[i] == (dir[i] == 255 ? 0 : 1 + sum of acc over the cells pointing at i): an exact identity, so zero violations
    means the accumulation raster is THE accumulation of that pointer raster"""
    down = downstream(d, table).reshape(-1)
    a = acc.reshape(-1).to(torch.int64)
    has = down >= 0
    inflow = torch.zeros_like(a).index_add_(0, down[has], a[has])
    want = torch.where(d.reshape(-1) == 255, torch.zeros_like(a), 1 + inflow)
    return int((want != a).sum())


def uphill_violations(filled, d, table):
    """a pointer never leads to a higher cell of the filled surface"""
    down = downstream(d, table).reshape(-1)
    f = filled.reshape(-1)
    has = down >= 0
    return int((f[down[has]] > f[has]).sum())


def fill_violations(z, filled, nodata, seed_any_nodata=True):
    """filled >= z, nodata untouched, seeds (raster frame / next to nodata) unchanged, and no valid non-seed cell is
    lower than all of its valid neighbours (a filled surface has no pits)"""
    ny, nx = z.shape
    nd = (z == nodata) | (z != z)
    bad = int(((filled < z) & ~nd).sum()) + int((nd & (filled != z) & ~(z != z)).sum())
    big = torch.finfo(torch.float32).max
    fp = torch.nn.functional.pad(torch.where(nd, torch.full_like(filled, big), filled)[None, None], (1, 1, 1, 1),
                                 value=big)[0, 0]
    ndp = torch.nn.functional.pad(nd[None, None].to(torch.uint8), (1, 1, 1, 1), value=1)[0, 0].bool()
    lowest = torch.full_like(filled, big)
    near_nd = torch.zeros_like(nd)
    for dr in (0, 1, 2):
        for dc in (0, 1, 2):
            if dr == 1 and dc == 1:
                continue
            lowest = torch.minimum(lowest, fp[dr:dr + ny, dc:dc + nx])
            near_nd |= ndp[dr:dr + ny, dc:dc + nx]
    seed = ~nd & near_nd                       # frame cells have padded (nodata) neighbours
    if seed_any_nodata:
        bad += int((seed & (filled != z)).sum())
    bad += int((~nd & ~seed & (lowest > filled)).sum())
    return bad
