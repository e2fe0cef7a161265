"""begin / end form of the host chain: two units in flight give the results of two synchronous calls."""
import numpy as np
import pytest

from pydrodem_b200 import _lib, synth


@pytest.mark.gpu
def test_two_units_in_flight_match_synchronous_calls():
    import torch
    import pydrodem_b200 as hc
    lib = _lib.load()
    ctx = _lib.context(0)
    units = [synth.make_dem_numpy(300, 411, 31), synth.make_dem_numpy(257, 513, 32), synth.make_dem_numpy(300, 411, 33)]
    want = [hc.fill_d8_accum(z, nodata=-9999.0, dx=30.0, dy=30.0, seed_mode=hc.SEED_TAUDEM, table=hc.TABLE_TAUDEM,
                             flats=True, want_slope=True) for z in units]

    def pinned(z):
        hz = torch.empty(z.shape, dtype=torch.float32, pin_memory=True)
        hz.copy_(torch.from_numpy(z))
        return dict(z=hz, f=torch.zeros(z.shape, dtype=torch.float32, pin_memory=True),
                    d=torch.zeros(z.shape, dtype=torch.uint8, pin_memory=True),
                    s=torch.zeros(z.shape, dtype=torch.float32, pin_memory=True),
                    a=torch.zeros(z.shape, dtype=torch.int32, pin_memory=True))

    bufs = [pinned(z) for z in units]

    def begin(slot, b):
        ny, nx = b["z"].shape
        _lib.check(lib.hc_fill_d8_accum_host_begin(ctx, slot, b["z"].data_ptr(), b["f"].data_ptr(), b["d"].data_ptr(),
                                                   b["s"].data_ptr(), b["a"].data_ptr(), ny, nx, -9999.0, 30.0, 30.0,
                                                   hc.SEED_TAUDEM, hc.TABLE_TAUDEM, 1), "begin")

    def end(slot):
        _lib.check(lib.hc_fill_d8_accum_host_end(ctx, slot), "end")

    begin(0, bufs[0])
    begin(1, bufs[1])
    end(0)
    begin(0, bufs[2])          # slot 0 re-used while unit 1 may still be travelling
    end(1)
    end(0)
    for b, w in zip(bufs, want):
        assert np.array_equal(b["f"].numpy(), w[0])
        assert np.array_equal(b["d"].numpy(), w[1])
        assert np.array_equal(b["s"].numpy(), w[2])
        assert np.array_equal(b["a"].numpy().view(np.uint32), w[3])
    assert lib.hc_fill_d8_accum_host_begin(ctx, 2, None, None, None, None, None, 4, 4, -9999.0, 1.0, 1.0, 0, 0, 1) != 0
    assert lib.hc_fill_d8_accum_host_end(ctx, 0) == 0      # idempotent
