"""Seeded synthetic DEMs for the configurations BASELINE.json names.

The reference ships no sample data, so the workloads are generated: multi-octave value-noise terrain
(fBm-like, H ~ 0.8) on a tilted plane, random single-cell pits, Gaussian bowls, optional nodata frame,
quantised to 1 cm so that float32 ties (flats) occur.  Lattice values come from an integer hash of
the lattice coordinates, so any row band can be generated on its own (40000^2 is built band-wise and
no host ever holds more than a band).  torch is used as an array library only (CPU or CUDA).
"""
import math

import numpy as np
import torch

NODATA = -9999.0


def _hash01(ix, iy, seed):
    """Deterministic lattice hash -> float64 in [0, 1).  int64 tensors in, all ops wrap mod 2^64."""
    h = ix * 73856093 ^ iy * 19349663 ^ (seed * 83492791 + 0x9E3779B9)
    h = (h ^ (h >> 13)) * 1274126177
    h = (h ^ (h >> 16)) * 2246822519
    h = h ^ (h >> 15)
    return (h & 0xFFFFFF).to(torch.float64) / float(1 << 24)


def _value_noise(rows, cols, spacing, seed):
    """Bilinear (smoothstep) value noise on a lattice of `spacing` cells; rows/cols are int64 1-D."""
    y = rows.to(torch.float64) / spacing
    x = cols.to(torch.float64) / spacing
    iy = torch.floor(y).to(torch.int64)
    ix = torch.floor(x).to(torch.int64)
    fy = (y - iy.to(torch.float64))
    fx = (x - ix.to(torch.float64))
    fy = fy * fy * (3 - 2 * fy)
    fx = fx * fx * (3 - 2 * fx)
    IY, IX = iy[:, None], ix[None, :]
    v00 = _hash01(IX, IY, seed)
    v01 = _hash01(IX + 1, IY, seed)
    v10 = _hash01(IX, IY + 1, seed)
    v11 = _hash01(IX + 1, IY + 1, seed)
    FY, FX = fy[:, None], fx[None, :]
    top = v00 + (v01 - v00) * FX
    bot = v10 + (v11 - v10) * FX
    return top + (bot - top) * FY


def terrain_band(ny, nx, row0, nrows, seed, relief=1500.0, hurst=0.8, rough=0.12, device="cpu"):
    """Rows [row0, row0+nrows) of the ny x nx terrain as float64 (before pits/bowls/quantisation)."""
    rows = torch.arange(row0, row0 + nrows, dtype=torch.int64, device=device)
    cols = torch.arange(0, nx, dtype=torch.int64, device=device)
    base = float(max(ny, nx)) / 3.0
    z = torch.zeros((nrows, nx), dtype=torch.float64, device=device)
    amp, total, o = 1.0, 0.0, 0
    spacing = base
    while spacing >= 2.0:
        z += amp * _value_noise(rows, cols, spacing, seed * 131 + o)
        total += amp
        amp *= 2.0 ** (-hurst)
        spacing /= 2.0
        o += 1
    # fBm roughness is a fraction of the regional trend so that, as on a fluvially eroded
    # landscape, closed depressions come from the pits/bowls below rather than from the noise
    z = z / total * (rough * relief)
    # tilted plane: most of the flow leaves through the south-east edges
    ry = rows.to(torch.float64)[:, None] / max(ny - 1, 1)
    rx = cols.to(torch.float64)[None, :] / max(nx - 1, 1)
    z = z + relief * (1.0 - 0.625 * ry - 0.375 * rx)
    return z


def _bowls(ny, nx, seed, n_bowls, must_rows=()):
    """Gaussian bowl parameters (row, col, sigma, depth); `must_rows` forces one bowl centre per row."""
    rng = np.random.default_rng(seed + 7)
    r = rng.uniform(0, ny, n_bowls)
    c = rng.uniform(0, nx, n_bowls)
    smax = max(5.0, min(80.0, min(ny, nx) / 12.0))
    s = rng.uniform(min(5.0, smax), smax, n_bowls)
    d = rng.uniform(2.0, 40.0, n_bowls)
    for i, mr in enumerate(must_rows):
        if i < n_bowls:
            r[i] = mr
    return np.stack([r, c, s, d], axis=1)


def make_dem(ny, nx, seed, pits=True, pit_frac=0.005, n_bowls=200, nodata_frame=False, quantise=True,
             row0=0, nrows=None, device="cpu", seam_rows=(), plateau_rows=()):
    """float32 DEM (rows [row0, row0+nrows) of an ny x nx raster) on `device`.

    pits: lower pit_frac of the cells by U(0.1, 5) m and add n_bowls Gaussian bowls (sigma 5-80 px,
    depth 2-40 m); seam_rows get a bowl centred on them, plateau_rows a 40-row-high exact plateau
    (configuration C3's seam-straddling depressions and flats)."""
    nrows = ny - row0 if nrows is None else nrows
    z = terrain_band(ny, nx, row0, nrows, seed, device=device)
    if pits:
        g = torch.Generator(device="cpu")
        for (br, bc, bs, bd) in _bowls(ny, nx, seed, n_bowls, must_rows=seam_rows):
            r_lo, r_hi = int(max(row0, br - 4 * bs)), int(min(row0 + nrows, br + 4 * bs + 1))
            c_lo, c_hi = int(max(0, bc - 4 * bs)), int(min(nx, bc + 4 * bs + 1))
            if r_lo >= r_hi or c_lo >= c_hi:
                continue
            rr = torch.arange(r_lo, r_hi, dtype=torch.float64, device=device)[:, None]
            cc = torch.arange(c_lo, c_hi, dtype=torch.float64, device=device)[None, :]
            z[r_lo - row0:r_hi - row0, c_lo:c_hi] -= bd * torch.exp(-((rr - br) ** 2 + (cc - bc) ** 2) / (2 * bs * bs))
        # single-cell pits: hash-selected so that bands agree
        rows = torch.arange(row0, row0 + nrows, dtype=torch.int64, device=device)[:, None]
        cols = torch.arange(0, nx, dtype=torch.int64, device=device)[None, :]
        sel = _hash01(cols, rows, seed * 977 + 1)
        dep = _hash01(cols, rows, seed * 977 + 2)
        z = torch.where(sel < pit_frac, z - (0.1 + 4.9 * dep), z)
        del g
    for pr in plateau_rows:
        r_lo, r_hi = max(row0, pr - 20), min(row0 + nrows, pr + 20)
        if r_lo < r_hi:
            c_lo, c_hi = nx // 4, nx // 4 + min(400, nx // 2)
            lvl = float(terrain_band(ny, nx, pr, 1, seed, device=device)[0, nx // 4])
            z[r_lo - row0:r_hi - row0, c_lo:c_hi] = lvl
    z32 = z.to(torch.float32)
    if quantise:
        z32 = torch.round(z32 * 100.0) / 100.0
    if nodata_frame:
        if row0 == 0:
            z32[0, :] = NODATA
        z32[:, 0] = NODATA
    return z32.contiguous()


def make_dem_numpy(ny, nx, seed, **kw):
    return make_dem(ny, nx, seed, device="cpu", **kw).numpy()


def random_dem(ny, nx, seed, kind="noise", nodata_holes=0, levels=None):
    """Small adversarial DEMs for parity tests (numpy float32)."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        z = rng.random((ny, nx))
    elif kind == "smooth":
        z = make_dem_numpy(ny, nx, seed, n_bowls=max(2, (ny * nx) // 4000), quantise=False).astype(np.float64)
    elif kind == "cone":
        yy, xx = np.mgrid[0:ny, 0:nx]
        z = np.hypot(yy - ny / 2.0, xx - nx / 2.0) + rng.random((ny, nx)) * 0.5
    elif kind == "bowl":
        yy, xx = np.mgrid[0:ny, 0:nx]
        z = -np.hypot(yy - ny / 2.0, xx - nx / 2.0) + rng.random((ny, nx)) * 2.0
    else:
        raise ValueError(kind)
    z = z.astype(np.float32)
    if levels:
        lo, hi = float(z.min()), float(z.max())
        z = (np.floor((z - lo) / max(hi - lo, 1e-9) * levels) / levels * (hi - lo) + lo).astype(np.float32)
    for _ in range(nodata_holes):
        h, w = rng.integers(1, max(2, ny // 4)), rng.integers(1, max(2, nx // 4))
        r, c = rng.integers(0, ny - h + 1), rng.integers(0, nx - w + 1)
        z[r:r + h, c:c + w] = NODATA
    return z


def make_c4(ny, nx, seed, device="cpu", n_lakes=None, n_rivers=4):
    """BASELINE.json configs[3]: the C2 terrain with the large EXACT flats pydrodem's water burning creates —
    lakes set to one constant (tools/water.py:594), meandering rivers 5-40 px wide set to a per-reach constant
    (tools/water.py:1383,1722), everything under water clamped to >= -1.0 (models/burn_rivers.py:754-762).
    Returns (float32 DEM, uint8 water mask with the codes of models/burn_rivers.py:309-324: 1 river, 3 lake)."""
    z = make_dem(ny, nx, seed, device=device, n_bowls=max(20, int(40 * ny * nx / (10801.0 * 10801.0))))
    mask = torch.zeros((ny, nx), dtype=torch.uint8, device=device)
    rng = np.random.default_rng(seed + 11)
    n_lakes = n_lakes if n_lakes is not None else max(3, int(60 * ny * nx / (10801.0 * 10801.0)))
    rows = torch.arange(ny, device=device, dtype=torch.float32)[:, None]
    cols = torch.arange(nx, device=device, dtype=torch.float32)[None, :]
    for _ in range(n_lakes):
        r, c = rng.uniform(0.05, 0.95) * ny, rng.uniform(0.05, 0.95) * nx
        rad = rng.uniform(0.004, 0.03) * min(ny, nx)
        r0, r1 = int(max(0, r - rad)), int(min(ny, r + rad + 1))
        c0, c1 = int(max(0, c - rad)), int(min(nx, c + rad + 1))
        if r0 >= r1 or c0 >= c1:
            continue
        disc = ((rows[r0:r1] - r) ** 2 + (cols[:, c0:c1] - c) ** 2) <= rad * rad
        sub = z[r0:r1, c0:c1]
        lvl = torch.round(sub[disc].min() * 100.0) / 100.0 if bool(disc.any()) else None
        if lvl is None:
            continue
        z[r0:r1, c0:c1] = torch.where(disc, lvl, sub)
        mask[r0:r1, c0:c1] = torch.where(disc, torch.tensor(3, dtype=torch.uint8, device=device), mask[r0:r1, c0:c1])
    for k in range(n_rivers):
        # a meandering polyline from the high (north-west) side to the low (south-east) side, cut into reaches
        t = np.linspace(0, 1, max(50, (ny + nx) // 16))
        rr = (0.05 + 0.9 * t) * ny + 0.06 * ny * np.sin(2 * np.pi * (3 + k) * t + k)
        cc = (0.1 + 0.2 * k + 0.6 * t) * nx % nx + 0.05 * nx * np.cos(2 * np.pi * (2 + k) * t)
        width = rng.uniform(5, 40) * min(1.0, min(ny, nx) / 3000.0) + 2
        for i in range(len(t) - 1):
            r, c = rr[i], cc[i]
            r0, r1 = int(max(0, r - width)), int(min(ny, r + width + 1))
            c0, c1 = int(max(0, c - width)), int(min(nx, c + width + 1))
            if r0 >= r1 or c0 >= c1:
                continue
            disc = ((rows[r0:r1] - r) ** 2 + (cols[:, c0:c1] - c) ** 2) <= (width / 2) ** 2
            if not bool(disc.any()):
                continue
            sub = z[r0:r1, c0:c1]
            lvl = torch.round(sub[disc].min() * 100.0) / 100.0
            z[r0:r1, c0:c1] = torch.where(disc, lvl, sub)
            mask[r0:r1, c0:c1] = torch.where(disc & (mask[r0:r1, c0:c1] == 0), torch.tensor(1, dtype=torch.uint8, device=device),
                                             mask[r0:r1, c0:c1])
    z = torch.where((mask > 0) & (z < -1.0), torch.tensor(-1.0, device=device), z)
    return z.contiguous(), mask
