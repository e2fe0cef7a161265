import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no libhydrocore.so (built artefacts are git-ignored): build it once, the way
    __graft_entry__.build() does (nvcc cross-compiles sm_100a without a GPU).  An existing library is never rebuilt here."""
    lib = os.path.join(ROOT, "pydrodem_b200", "libhydrocore.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "pydrodem_b200", "csrc"), "-j8"],
                              stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle():
    from oracle import hydro_oracle
    hydro_oracle.build()
    return hydro_oracle


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test running without a CUDA device")
    return torch.device("cuda:0")
