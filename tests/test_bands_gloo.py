"""world_size-2 gloo run (CPU only) of the row-band exchange protocol: tests/run_bands_gloo.py holds a plain
Python model of the hc_band_* phases and talks through the very same bands.SeamComm / exchange_halos the GPU
path uses; both ranks must reproduce the whole-raster oracle for fill and accumulation."""
import os
import subprocess
import sys


def test_bands_protocol_world2_gloo():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29617", os.path.join(root, "tests", "run_bands_gloo.py")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "BANDS_GLOO_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]


def test_split_rows_and_single_rank_comm():
    import torch
    from pydrodem_b200 import bands
    assert bands.split_rows(10, 3) == [4, 3, 3] and sum(bands.split_rows(40000, 7)) == 40000
    comm = bands.SeamComm()
    assert comm.world == 1
    a = [torch.arange(12, dtype=torch.float32).reshape(4, 3).clone(), torch.arange(100, 115, dtype=torch.float32).reshape(5, 3).clone()]
    bands.exchange_halos(a, 0, 2, comm)     # two local bands: halos come from each other
    assert torch.equal(a[0][-1], a[1][1]) and torch.equal(a[1][0], a[0][-2])
    assert comm.any(True, torch.device("cpu")) and not comm.any(False, torch.device("cpu"))
