"""Image-level sharding over one process per GPU and the rate-statistics reduction.

Every image (or 4K image treated as one unit) is coded independently -- own header, own min/max,
own streams (reference: the eval loop of agents/llicti_agent.py:129-149 walks images with batch
size 1) -- so the hot path needs no collective.  The only exchange is the reduction of the
(pixels, bytes, launches, ...) sums and of the per-rank device times (max) at the end.

The training step is the one place with a real exchange: under `torchrun` every rank takes its share of each
batch and the weight gradients are averaged (one all-reduce of the 24 tensors as one flat buffer, NCCL over NVLink)
before clipping and the optimizer step -- the loss is a mean over the batch, so the average of the ranks' gradients is
the gradient of the global batch.
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


def average_gradients(params) -> int:
    """All-reduce (mean) of the `.grad` of `params` over the ranks as ONE flat buffer; identity without an initialised
    process group.  Returns the number of elements exchanged (0 when nothing was)."""
    dist = torch.distributed
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
    return off


def rank_share(order: Sequence[int], rank: int, world: int):
    """This rank's items of one epoch's (shared) random order: every world-th item, the same count on every rank (the
    remainder of an order that does not divide is dropped for the epoch, so that all ranks take the same number of
    optimizer steps); a set smaller than the world is left whole on every rank."""
    order = list(order)
    if world <= 1 or len(order) < world:
        return order
    per = len(order) // world
    return order[rank:per * world:world]


def bind_host_to_gpu(device_index: int):
    """Restrict this process's host threads to the CPUs NVML reports as local to CUDA device `device_index` (the GPU's NUMA
    node), so that pinned staging buffers allocated afterwards are local to the GPU's PCIe root: with one process per GPU
    on a two-socket box, host<->device copies of half the ranks otherwise cross the socket interconnect.  Returns the CPU
    list bound to, or None when nothing was changed (no NVML, no affinity API, an empty intersection with the CPUs this
    process may use, or LLICTI_NUMA_BIND=0)."""
    import os
    if os.environ.get("LLICTI_NUMA_BIND", "1") == "0" or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:       # noqa: BLE001 -- older bindings want bytes
            h = pynvml.nvmlDeviceGetHandleByUUID((("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid).encode())
        allowed = os.sched_getaffinity(0)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed | {os.cpu_count() or 1}) // 64) + 1)
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus = sorted(local & allowed)
        if not cpus or set(cpus) == set(allowed):
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:           # noqa: BLE001 -- an optimisation, never a requirement
        return None
