"""Time-sharded execution across ranks (one process per GPU, torch.distributed for the plumbing).

Time steps are independent through contours -> indices -> properties -> to_xarray
(utils/index_utils.py:217-258), so every rank owns a contiguous block of the time axis and no data-path
collective is needed.  Three small host-side merges remain:

1. ``exp_lon.max()`` is global over all dates (streamer_index.py:106) -> scalar MAX all-reduce;
2. event ids / row indices are global (streamer_index.py:281-283) -> exclusive prefix sum of the per-rank counts;
3. the event tables are gathered to rank 0 in rank (= time) order.
"""

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(ntime, rank, world_size):
    """Contiguous block [t0, t1) of the time axis owned by ``rank`` (earlier ranks get the remainder)."""
    base, rem = divmod(int(ntime), int(world_size))
    t0 = rank * base + min(rank, rem)
    return t0, t0 + base + (1 if rank < rem else 0)


def _device_for_collectives():
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def global_max(value):
    """MAX of an int over all ranks (the global exp_lon.max() / dlon)."""
    rank, ws = world()
    if ws == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64, device=_device_for_collectives())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def exclusive_offset(count):
    """(offset of this rank, total) for per-rank counts: global event ids = offset + local index."""
    rank, ws = world()
    if ws == 1:
        return 0, int(count)
    t = torch.zeros(ws, dtype=torch.int64, device=_device_for_collectives())
    t[rank] = int(count)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    counts = t.cpu().numpy()
    return int(counts[:rank].sum()), int(counts.sum())


def gather_frames(frame, dst=0):
    """Concatenate per-rank pandas tables on ``dst`` in rank (= time) order; other ranks get None."""
    import pandas as pd

    rank, ws = world()
    if ws == 1:
        return frame.reset_index(drop=True)
    parts = [None] * ws if rank == dst else None
    dist.gather_object(frame, parts, dst=dst)
    if rank != dst:
        return None
    parts = [p for p in parts if p is not None and len(p)]
    return pd.concat(parts).reset_index(drop=True) if parts else frame.iloc[:0]


def run_sharded(detector, raw_full_or_shard, ntime_total=None, is_shard=False):
    """Run ``detector`` on this rank's block of time steps with the global exp_lon maximum.

    Returns ``(t0, t1, BatchResult)``.  If the maximum of another rank exceeds the local one, the index
    stage is re-run with the global value (rare: every step normally has a full-width contour).
    """
    rank, ws = world()
    if is_shard:
        shard = raw_full_or_shard
        t0, t1 = shard_range(ntime_total, rank, ws)
    else:
        t0, t1 = shard_range(raw_full_or_shard.shape[0], rank, ws)
        shard = raw_full_or_shard[t0:t1]
    res = detector.run_batch(shard)
    g = global_max(res.gmax_nx)
    if g != res.gmax_nx:
        res = detector.run_batch(shard, gmax_nx=g)
    return t0, t1, res


def track_sharded(local_events, dst=0, **kwargs):
    """``track_events`` across ranks: event tables are gathered on ``dst`` in time order and tracked there
    (events.py:113-241 links t to (t, t + time_range], i.e. across shard boundaries too); other ranks get None."""
    from . import tracking

    merged = gather_frames(local_events, dst=dst)
    if merged is None:
        return None
    return tracking.track_events(merged, **kwargs)
