"""world_size-N gloo check of units.run_units: every unit runs exactly once, on one rank, and a slow unit does not
hold the queue (launched by tests/test_units_gloo.py through torch.distributed.run)."""
import os
import sys
import time

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pydrodem_b200 import units  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    def work(u, scale):
        time.sleep(0.25 if u == 19 else 0.01)      # the first unit drawn (descending order) is the slow one
        return (u * scale, rank)

    ids = list(range(20))
    mine = units.run_units(work, ids, 3)
    allr = units.run_units(work, ids, 5, gather=True)
    ok = sorted(allr) == ids and all(allr[u][0] == 5 * u for u in ids)
    counts = [None] * world
    dist.all_gather_object(counts, sorted(mine))
    flat = sorted(u for c in counts for u in c)
    ok = ok and flat == ids and all(v[0] == 3 * u and v[1] == rank for u, v in mine.items())
    # dynamic queue: the rank that drew the slow unit ran fewer units than the others
    slow_rank = [r for r, c in enumerate(counts) if 19 in c][0]
    if world > 1:
        ok = ok and len(counts[slow_rank]) <= min(len(c) for r, c in enumerate(counts) if r != slow_rank)
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if rank == 0:
        print("UNITS_GLOO_OK" if all(flags) else f"UNITS_GLOO_FAIL {flags} {counts}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
