"""TauDEM-compatible command lines on libhydrocore (the flags pydrodem builds at run_aws.py:580-617):

    python -m pydrodem_b200.taudem PitRemove -z dem.tif -fel filled.tif
    python -m pydrodem_b200.taudem D8FlowDir -fel filled.tif -p flowdir.tif -sd8 slope.tif
    python -m pydrodem_b200.taudem AreaD8 [-nc] -p flowdir.tif -ad8 flowacc.tif

`p` is written as int16 with TauDEM's codes 1..8 (1 = E, counter-clockwise) and -32768 where TauDEM
leaves no direction (nodata, raster border, cells next to nodata); `ad8` as float32 cell counts with
-1 as nodata, which is how the consumers read them (models/taudem_check.py:46-47, :105-147).
"""
import argparse
import sys

import numpy as np

from . import core, raster_io
from ._lib import SEED_TAUDEM, TABLE_TAUDEM


def _nodata(meta):
    return float(meta.nodata) if meta.nodata is not None else -9999.0


def pit_remove(z_path, fel_path):
    z, meta = raster_io.read_raster(z_path)
    nd = _nodata(meta)
    raster_io.write_raster(fel_path, core.fill(np.asarray(z, np.float32), nodata=nd, seed_mode=SEED_TAUDEM), meta, nodata=nd)


def d8_flow_dir(fel_path, p_path, sd8_path=None):
    import torch
    z, meta = raster_io.read_raster(fel_path)
    nd = _nodata(meta)
    dx, dy = meta.pixel_size
    zt = torch.from_numpy(np.ascontiguousarray(z, np.float32)).cuda()
    d, s = core.d8(zt, nodata=nd, dx=dx, dy=dy, table=TABLE_TAUDEM, want_slope=True)
    d = core.resolve_flats(zt, d, nodata=nd, table=TABLE_TAUDEM)
    p = d.cpu().numpy().astype(np.int16)
    p[p == 255] = -32768
    raster_io.write_raster(p_path, p, meta, nodata=-32768)
    if sd8_path:
        raster_io.write_raster(sd8_path, s.cpu().numpy(), meta, nodata=-1.0)


def area_d8(p_path, ad8_path, nc=True):
    p, meta = raster_io.read_raster(p_path)
    d = np.where(p == -32768, 255, p).astype(np.uint8)
    a = core.accum(d, table=TABLE_TAUDEM).astype(np.float32)
    a[d == 255] = -1.0
    raster_io.write_raster(ad8_path, a, meta, nodata=-1.0)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    tool = argv.pop(0).lower().replace(".exe", "")
    ap = argparse.ArgumentParser(prog=tool)
    if tool == "pitremove":
        ap.add_argument("-z", required=True)
        ap.add_argument("-fel", required=True)
        a = ap.parse_args(argv)
        pit_remove(a.z, a.fel)
    elif tool == "d8flowdir":
        ap.add_argument("-fel", required=True)
        ap.add_argument("-p", required=True)
        ap.add_argument("-sd8")
        a = ap.parse_args(argv)
        d8_flow_dir(a.fel, a.p, a.sd8)
    elif tool == "aread8":
        ap.add_argument("-p", required=True)
        ap.add_argument("-ad8", required=True)
        ap.add_argument("-nc", action="store_true")
        a = ap.parse_args(argv)
        area_d8(a.p, a.ad8, a.nc)
    else:
        print(f"unknown tool {tool!r}; expected PitRemove, D8FlowDir or AreaD8", file=sys.stderr)
        return 2
    return 0


if __name__ == "__main__":
    sys.exit(main())
