"""Utterance sharding across GPUs: one process per GPU, no collective on the
data path.  The reference has no working multi-GPU inference (SURVEY.md §2a:
``Synthesizer`` placeholders are batch-1); utterances are independent, so rank
``r`` of ``W`` takes a contiguous block and only the OUTPUTS are gathered
(NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of utterances for ``rank``: sizes differ by at most one and
    earlier ranks get the larger blocks.  Empty blocks are allowed (n < world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(arrays: Sequence, world: int, rank: int) -> List:
    """Slice every array of a batch (first axis = utterance) for ``rank``; ``None`` passes through."""
    n = len(next(a for a in arrays if a is not None))
    lo, hi = shard_bounds(n, world, rank)
    return [None if a is None else a[lo:hi] for a in arrays]


def gather_outputs(local: torch.Tensor, n_total: int, group=None, dst: Optional[int] = None) -> Optional[torch.Tensor]:
    """Concatenate the per-rank output blocks along the utterance axis.

    Blocks may have different sizes; each rank pads to the largest block so one
    ``all_gather`` (or ``gather`` to ``dst``) moves everything.  Returns the full
    ``[n_total, ...]`` tensor (on ``dst`` only when ``dst`` is given)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    big = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if dst is None:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
    else:
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
