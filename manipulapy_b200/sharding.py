"""Multi-GPU sharding of the hot path.

Every unit of work (a trajectory point, a configuration, a whole rollout) is independent,
so the batch is split into contiguous index ranges, one per rank (one process per GPU,
``torch.distributed``); robot constants are replicated (< 4 KB) and each rank computes its
own slice.  The only collective is the optional final gather of result rows (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  No reduction, no exchange step.
"""

from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(units: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous ``[lo, hi)`` of ``units`` owned by ``rank``: ``ceil(units / world)`` per rank."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    per = -(-int(units) // world_size)
    lo = min(units, rank * per)
    return lo, min(units, lo + per)


def gather_rows(local: torch.Tensor, units: int, group: Optional[dist.ProcessGroup] = None,
                dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Concatenate the per-rank row slices produced under ``shard_range`` back into ``units`` rows.

    ``dst=None``: all-gather (every rank gets the result); otherwise only ``dst`` does
    (``None`` elsewhere).  Slices are padded to the common ``ceil(units / world)`` rows so a
    single fixed-size collective moves everything.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-int(units) // world)
    tail = local.shape[1:]
    buf = local
    if local.shape[0] != per:
        buf = local.new_zeros((per, *tail))
        buf[: local.shape[0]] = local
    buf = buf.contiguous()
    if dst is None:
        out = local.new_empty((world * per, *tail))
        dist.all_gather_into_tensor(out, buf, group=group)
        return out[:units]
    pieces = [local.new_empty((per, *tail)) for _ in range(world)] if rank == dst else None
    dist.gather(buf, pieces, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat(pieces, 0)[:units]
