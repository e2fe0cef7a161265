"""The C-ABI library loads and exports every symbol include/hydrocore.h declares (no compute calls:
this suite runs without a GPU), the Python binding types all of them, and the product path fails
loudly instead of falling back when no CUDA device exists."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "hydrocore.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hc_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = header_symbols()
    for need in ("hc_fill_dev", "hc_d8_dev", "hc_resolve_flats_dev", "hc_accum_dev", "hc_fill_d8_accum_dev",
                 "hc_label_dev", "hc_fill_host", "hc_fill_d8_accum_host", "hc_label_host", "hc_create", "hc_destroy",
                 "hc_last_error"):
        assert need in syms


def test_library_exports_every_declared_symbol():
    from pydrodem_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libhydrocore.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in hydrocore.h but not exported"


def test_binding_covers_the_header():
    from pydrodem_b200 import _lib
    assert sorted(_lib.SIGNATURES) == header_symbols()
    _lib.load()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("only meaningful on a box without a GPU")
    import numpy as np
    import pydrodem_b200 as hc
    with pytest.raises(hc.HydrocoreError, match="no CUDA device|CPU"):
        hc.fill(np.zeros((4, 4), np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pydrodem_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            txt = open(os.path.join(dirpath, f), errors="replace").read() if f.endswith((".py", ".cu", ".cuh", ".h")) else ""
            if f.endswith(".py"):
                assert not re.search(r"^\s*(from|import)\s+oracle\b|hydro_oracle|_build/libhydro", txt, re.M), f
            else:
                assert not re.search(r"#include\s+[\"<][^\">]*oracle", txt), f
