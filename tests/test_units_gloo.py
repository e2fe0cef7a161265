"""Unit-level parallelism (units.run_units): single process loop, and a world-3 gloo job sharing one dynamic queue."""
import os
import subprocess
import sys


def test_run_units_single_process_order_and_results():
    from pydrodem_b200 import units
    seen = []
    out = units.run_units(lambda u, k, off=0: (seen.append(u), u * k + off)[1], [3, 11, 7], 2, off=1)
    assert seen == [11, 7, 3]                      # descending ids, as the reference sorts its units
    assert out == {11: 23, 7: 15, 3: 7}
    seen.clear()
    units.run_units(lambda u: seen.append(u), ["b", "a"], sort_descending=False)
    assert seen == ["b", "a"]


def test_run_units_world3_gloo():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "3", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(root, "tests", "run_units_gloo.py")]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "UNITS_GLOO_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
