"""Unit-level parallelism (last row): independent units -- tiles or basins of a region -- are spread over
the GPUs of a box, one process per GPU, with no data-path collective.

The reference hands its units to `MPICommExecutor.starmap` or `multiprocessing.Pool.starmap`
(models/depression_handling.py:1606-1635, models/find_sinks.py:171-200): a dynamic queue, units sorted in descending id
order first.  `run_units` is that for a `torch.distributed` job (`python -m torch.distributed.run --nproc-per-node N`):
the queue is one shared counter in the job's key-value store, so a rank that draws a small unit simply comes back for
the next one; with a single process it is a plain loop.
"""
import numpy as np


def run_units(fn, unit_ids, *args, group=None, gather=False, sort_descending=True, **kwargs):
    """Call `fn(unit_id, *args, **kwargs)` once per unit, each unit on exactly one rank.

    Returns {unit_id: result} for the units THIS rank ran; with gather=True every rank gets the merged dictionary
    (results must be picklable).  The caller binds its rank to a GPU (`torch.cuda.set_device(LOCAL_RANK)`) as usual."""
    import torch.distributed as dist
    ids = list(np.sort(np.asarray(list(unit_ids)))[::-1]) if sort_descending else list(unit_ids)
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    mine = {}
    if not multi:
        for u in ids:
            mine[_key(u)] = fn(u, *args, **kwargs)
        return mine
    store = dist.distributed_c10d._get_default_store()
    key = f"pydrodem_b200/run_units/{_call_index(store, group)}"
    while True:
        i = store.add(key, 1) - 1
        if i >= len(ids):
            break
        mine[_key(ids[i])] = fn(ids[i], *args, **kwargs)
    dist.barrier(group)
    if dist.get_rank(group) == 0:
        try:
            store.delete_key(key)      # every member has drawn its last index: the counter is garbage
        except (RuntimeError, NotImplementedError):
            pass                       # (a store without delete support keeps the few bytes)
    if not gather:
        return mine
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, mine, group=group)
    merged = {}
    for p in parts:
        merged.update(p)
    return merged


_calls = {}


def _call_index(store, group):
    """Name of the shared counter of this run_units call: the group's global ranks (identical on all its members, distinct
    between disjoint groups) + how many run_units calls this process has made on that group (every member makes the
    same sequence of calls)."""
    import torch.distributed as dist
    ranks = tuple(dist.get_process_group_ranks(group)) if group is not None else ("world",)
    _calls[ranks] = _calls.get(ranks, 0) + 1
    return "-".join(str(r) for r in ranks) + f"/{_calls[ranks]}"


def _key(u):
    return u.item() if hasattr(u, "item") else u
