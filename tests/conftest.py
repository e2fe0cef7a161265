import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """Built artefacts are git-ignored, and a stale libhydrocore.so would silently test yesterday's kernels: run the
    (incremental) make at every session start -- an up-to-date library costs a fraction of a second; nvcc cross-compiles
    sm_100a without a GPU.  Without nvcc (a box that only received the built library) an existing .so is used as is."""
    import shutil
    import subprocess
    lib = os.path.join(ROOT, "pydrodem_b200", "libhydrocore.so")
    if shutil.which("nvcc") is None:
        if not os.path.exists(lib):
            print("conftest: nvcc not found and no libhydrocore.so: GPU and ABI tests will fail loudly", file=sys.stderr)
        return
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "pydrodem_b200", "csrc"), "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-4000:] + r.stderr[-8000:])
        raise RuntimeError("building libhydrocore.so failed")


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
