"""Multi-GPU plumbing: one process per GPU; AR inference shards clips (no data-path collective), training all-reduces
one flat gradient bucket (mint_b200/trainer.py).  The reference's counterpart is tf.distribute.MirroredStrategy
(trainer.py:125-135)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_clips(num_clips: int, rank: int, world_size: int) -> list[int]:
    """Round-robin assignment of evaluation clips to ranks (each clip's AR chain is independent)."""
    return list(range(rank, num_clips, world_size))


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed numbers are reported as the max over ranks."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
