"""endo.find_sink's raster logic (tools/endo.py:285-420) on the GPU vs a statement-by-statement numpy/scipy replay
of the reference lines (the function itself is wrapped in GDAL / geopandas / shapely I/O that is not installed)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GT = (100.0, 0.5, 0.0, 40.0, 0.0, -0.25)


def replay_find_sink(raster_array, SWO_array=None, gt=GT):
    from scipy.ndimage import label
    low_threshold = np.percentile(raster_array, 1, axis=None)                       # :286
    min_elev_ind = np.where(raster_array <= low_threshold)                          # :287
    if len(min_elev_ind[0]) == 1:
        return int(min_elev_ind[0][0]), int(min_elev_ind[1][0])
    structure = [[1, 1, 1], [1, 1, 1], [1, 1, 1]]
    minmask = np.zeros(raster_array.shape, dtype=int)
    minmask[min_elev_ind] = 1
    label_array, num_features = label(minmask, structure)                           # :323
    count_list = [len(np.where(label_array == f)[0]) for f in np.arange(1, num_features + 1)]
    biggest_feature = np.argmax(count_list) + 1
    min_contig = np.where(label_array == biggest_feature)
    sink_area_mask = np.zeros(raster_array.shape, dtype=int)
    if SWO_array is not None:
        SWO_mask_ind = np.where(SWO_array >= 80)
        a = np.ravel_multi_index(min_contig, raster_array.shape)
        b = np.ravel_multi_index(SWO_mask_ind, raster_array.shape)
        comb = np.unravel_index(np.intersect1d(a, b), raster_array.shape)
        sink_area_mask[comb] = 1
        if len(np.where(sink_area_mask == 1)[0]) > 0:
            min_contig = np.where(sink_area_mask == 1)
        else:
            sink_area_mask = np.zeros(raster_array.shape, dtype=int)
            sink_area_mask[min_contig] = 1
    else:
        sink_area_mask[min_contig] = 1
    min_contig_min = raster_array[min_contig].min()                                 # :379
    rr, cc = np.where((sink_area_mask == 1) & (raster_array == min_contig_min))
    if len(rr) == 1:
        return int(rr[0]), int(cc[0])
    x0, dx, _, y0, _, dy = gt
    xs = np.array([x0 + c * dx + dx / 2.0 for c in cc])
    ys = np.array([y0 + r * dy + dy / 2.0 for r in rr])
    d2 = (xs - xs.sum() / len(xs)) ** 2 + (ys - ys.sum() / len(ys)) ** 2            # nearest_points(centroid, pixels)
    j = int(np.argmin(d2))
    return int(rr[j]), int(cc[j])


def _basin(seed, ny, nx, levels=None):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:ny, 0:nx]
    z = 200 + 0.002 * ((yy - ny * 0.6) ** 2 + (xx - nx * 0.4) ** 2) + rng.normal(0, 0.3, (ny, nx))
    z += 8 * np.exp(-((yy - ny * 0.3) ** 2 + (xx - nx * 0.7) ** 2) / 200.0) * -1
    z = z.astype(np.float32)
    if levels:
        z = (np.round(z * levels) / levels).astype(np.float32)
    z[:3, :] = 99999.0
    z[:, -2:] = 99999.0
    return z


@pytest.mark.parametrize("seed,shape,levels,with_swo", [(1, (120, 150), None, False), (2, (200, 170), 2, False),
                                                       (3, (90, 260), 4, True), (4, (300, 300), 1, True),
                                                       (5, (64, 64), None, True)])
def test_find_sink_pixel_matches_reference_logic(seed, shape, levels, with_swo, cuda):
    from pydrodem_b200.tools import endo
    z = _basin(seed, *shape, levels=levels)
    swo = None
    if with_swo:
        rng = np.random.default_rng(seed + 100)
        swo = (rng.random(shape) * 100).astype(np.uint8)
        if seed == 5:
            swo[:] = 0          # no water at all: falls back to the low-elevation blob
    want = replay_find_sink(z, swo)
    got = endo.find_sink_pixel(z, swo, GT)
    assert got == want


def test_kth_value_and_percentile_exact(cuda):
    import torch
    from pydrodem_b200.tools import endo
    rng = np.random.default_rng(7)
    for n in (1, 2, 37, 1000, 1_000_003):
        x = (rng.normal(0, 100, n)).astype(np.float32)
        x[rng.integers(0, n, max(1, n // 10))] = np.float32(-3.5)     # duplicates
        t = torch.from_numpy(x).cuda()
        s = np.sort(x)
        for k in sorted({0, n // 3, n // 2, n - 1}):
            assert endo.kth_value(t, k) == s[k]
        for q in (1, 50, 99.5):
            got = endo._percentile_from_order_stats(n, q, lambda i: s[i])
            assert got == np.percentile(x, q), (n, q)
