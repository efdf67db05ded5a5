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


class SymmetricArena:
    """One replica's [gradient bucket fp32 | master weights fp32 | bf16 weight mirror] in torch symmetric memory
    (torch.distributed._symmetric_memory: cuMem allocations every rank of the group maps, plus the NVSwitch multicast
    mapping when the fabric has one).  The plumbing is torch's; the kernel that uses the peer mappings is ours
    (fact_dp_adam_step: cross-replica gradient sum + Adam + weight broadcast in one launch over NVLink)."""

    ALIGN = 256

    def __init__(self, numel: int, device, group=None):
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.numel = int(numel)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise RuntimeError("fact_dp_adam_step supports at most 16 replicas")
        up = lambda v: (v + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.grad_off = 0
        self.w_off = up(4 * self.numel)
        self.wb_off = self.w_off + up(4 * self.numel)
        self.nbytes = self.wb_off + up(2 * self.numel)
        self.buf = symm_mem.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        ptrs = list(self.handle.buffer_ptrs)
        if len(ptrs) != self.world or any(p == 0 for p in ptrs):
            raise RuntimeError("symmetric memory rendezvous returned no peer mappings")
        self.peer_ptrs = (C.c_void_p * self.world)(*ptrs)
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.mc_ptr = mc if mc != 0 else None
        self.grad = self.buf[self.grad_off:self.grad_off + 4 * self.numel].view(torch.float32)
        self.w = self.buf[self.w_off:self.w_off + 4 * self.numel].view(torch.float32)
        self.wb = self.buf[self.wb_off:self.wb_off + 2 * self.numel].view(torch.bfloat16)
        self.buf.zero_()

    def barrier(self) -> None:
        """Device-side barrier over the replicas on the current stream (signal pads of the symmetric memory)."""
        self.handle.barrier(channel=0)

    def shard(self) -> tuple[int, int]:
        """(offset, count) of the bucket elements this rank's optimizer state covers (fact_dp_adam_step's split)."""
        per = -(-self.numel // self.world)
        per = (per + 7) // 8 * 8
        lo = min(per * self.rank, self.numel)
        return lo, min(lo + per, self.numel) - lo
