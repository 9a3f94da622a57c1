"""Multi-GPU = independent camera+IMU streams sharded over ranks; no collective on the data path.

A frame pair does not shard (every feature reads arbitrary windows of both pyramids and the only
cross-feature step is the per-pair mean pixel error, reference src/gyro_aided_tracker.cpp:294-308), so each
rank owns whole streams on its own GPU ("replicas only" in the driver's vocabulary).  torch.distributed is
used for the rendezvous, the barrier and two scalar reductions of the measurement (time MAX, counts SUM)."""
from __future__ import annotations

from typing import Tuple


def streams_of_rank(n_streams: int, rank: int, world: int) -> range:
    """stream s lives on rank s mod world"""
    return range(rank, n_streams, world)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _device():
    import torch
    d = _dist()
    if d is not None and d.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def reduce_time_max(seconds: float) -> float:
    import torch
    d = _dist()
    if d is None:
        return float(seconds)
    t = torch.tensor([seconds], dtype=torch.float64, device=_device())
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def reduce_counts(*counts: int) -> Tuple[int, ...]:
    import torch
    d = _dist()
    if d is None:
        return tuple(int(c) for c in counts)
    t = torch.tensor([float(c) for c in counts], dtype=torch.float64, device=_device())
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return tuple(int(round(x)) for x in t.tolist())
