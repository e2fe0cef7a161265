"""Randomised properties (hypothesis) of the host-side pieces: the TIFF LZW codec and the breach."""
import ctypes

import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from oracle import breach_oracle as bo  # noqa: E402
from oracle import hydro_oracle as ho  # noqa: E402
from pydrodem_b200 import _lib  # noqa: E402

ND = -9999.0


def _enc(b):
    lib = _lib.load()
    src = np.frombuffer(b, np.uint8)
    dst = np.empty(2 * len(b) + 16, np.uint8)
    n = ctypes.c_int64()
    assert lib.hc_lzw_encode_host(src.ctypes.data if len(b) else None, len(b), dst.ctypes.data, len(dst), ctypes.byref(n)) == 0
    return dst[:n.value].tobytes()


def _dec(b, nout):
    lib = _lib.load()
    src = np.frombuffer(b, np.uint8)
    dst = np.empty(max(nout, 1), np.uint8)
    n = ctypes.c_int64()
    assert lib.hc_lzw_decode_host(src.ctypes.data, len(b), dst.ctypes.data, nout, ctypes.byref(n)) == 0
    return dst[:n.value].tobytes()


@settings(max_examples=150, deadline=None)
@given(st.binary(max_size=3000), st.integers(1, 6))
def test_lzw_round_trip_any_bytes(data, rep):
    b = data * rep                       # repetition drives long dictionary strings and the code == next case
    assert _dec(_enc(b), len(b)) == b


@settings(max_examples=60, deadline=None)
@given(st.binary(min_size=1, max_size=400))
def test_lzw_decoder_never_overruns_on_garbage(data):
    lib = _lib.load()
    src = np.frombuffer(data, np.uint8)
    dst = np.full(300 + 8, 0xAB, np.uint8)
    n = ctypes.c_int64()
    lib.hc_lzw_decode_host(src.ctypes.data, len(data), dst.ctypes.data, 300, ctypes.byref(n))   # may fail, must not overrun
    assert 0 <= n.value <= 300 and np.all(dst[300:] == 0xAB)


@settings(max_examples=40, deadline=None)
@given(st.integers(5, 18), st.integers(5, 18), st.integers(0, 2 ** 31 - 1), st.integers(0, 6), st.booleans())
def test_breach_equals_restatement_and_drains(ny, nx, seed, max_length, fill_pits):
    rng = np.random.default_rng(seed)
    z = rng.integers(0, 6, (ny, nx)).astype(np.float32)          # tiny range: ties, flats and pits everywhere
    z[rng.random((ny, nx)) < 0.06] = ND
    out = np.empty_like(z)
    st_ = (ctypes.c_int64 * 5)()
    assert _lib.load().hc_breach_depressions_host(z.ctypes.data, out.ctypes.data, ny, nx, ND, max_length, 0.0,
                                                  int(fill_pits), 1, st_) == 0
    want, sw = bo.breach_depressions(z, ND, max_length, None, fill_pits, True)
    assert np.array_equal(out, want)
    assert [st_[0], st_[1], st_[2], st_[3]] == [sw["pits"], sw["breached"], sw["left"], sw["cut"]]
    assert np.array_equal(ho.fill(out, ND, ho.SEED_WBT), out)    # depression free
    assert np.array_equal(out == ND, z == ND)


@settings(max_examples=40, deadline=None)
@given(st.integers(6, 16), st.integers(6, 18), st.integers(0, 2 ** 31 - 1), st.integers(2, 12), st.booleans(),
       st.sampled_from([None, 0.5, 3.0, 40.0]), st.sampled_from([None, 0.01]))
def test_least_cost_breach_specification_properties(ny, nx, seed, levels, min_dist, max_cost, eps):
    """oracle/breach_leastcost.py on random small surfaces: never raises a cell, accounts for every pit, can only reduce
    what a fill has to add, respects the cost limit, and with a window that sees the whole raster and no limit every pit
    drains (a fill adds nothing afterwards)"""
    from oracle import breach_leastcost as lc
    rng = np.random.default_rng(seed)
    z = rng.integers(0, levels, (ny, nx)).astype(np.float32)
    if seed % 3 == 0:
        z[ny // 2, nx // 2] = ND
    out, st_ = lc.breach_least_cost(z, ND, dist=4, max_cost=max_cost, min_dist=min_dist, flat_increment=eps, return_stats=True)
    valid = z != ND
    assert (out[valid] <= z[valid]).all() and np.array_equal(out[~valid], z[~valid])
    assert st_["pits"] == st_["breached"] + st_["left"] and st_["cells_lowered"] == int((out < z).sum())
    v0 = float((ho.fill(z, ND, ho.SEED_TAUDEM) - z)[valid].sum())
    v1 = float((ho.fill(out, ND, ho.SEED_TAUDEM) - out)[valid].sum())
    assert v1 <= v0 + 1e-4
    if max_cost is not None and eps is None:
        # level channels: what one breach removes is at most its cost limit; all of them together at most pits x limit
        assert float((z - out)[valid].sum()) <= st_["breached"] * max_cost * (1 + 1e-6) + 1e-3
    if eps is None:
        full = lc.breach_least_cost(z, ND, dist=max(ny, nx), min_dist=min_dist)
        assert np.array_equal(ho.fill(full, ND, ho.SEED_TAUDEM), full)
