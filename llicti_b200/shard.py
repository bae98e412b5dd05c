"""Image-level sharding over one process per GPU and the rate-statistics reduction.

Every image (or 4K image treated as one unit) is coded independently -- own header, own min/max,
own streams (reference: the eval loop of agents/llicti_agent.py:129-149 walks images with batch
size 1) -- so the hot path needs no collective.  The only exchange is the reduction of the
(pixels, bytes, launches, ...) sums and of the per-rank device times (max) at the end.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [start, stop) of rank `rank`: ceil(n/world) items per rank, the last
    ranks may get fewer (or none)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    per = -(-n_items // world)
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def reduce_stats(sums: Sequence[float], maxs: Sequence[float], device=None):
    """Sum `sums` and take the maximum of `maxs` over all ranks (identity without an initialised
    process group).  NCCL on GPUs, gloo on CPU: whatever backend the group was created with."""
    s = torch.tensor(list(sums), dtype=torch.float64, device=device)
    m = torch.tensor(list(maxs), dtype=torch.float64, device=device)
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        torch.distributed.all_reduce(s, op=torch.distributed.ReduceOp.SUM)
        torch.distributed.all_reduce(m, op=torch.distributed.ReduceOp.MAX)
    return s.tolist(), m.tolist()
