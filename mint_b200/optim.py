"""Keras-semantics Adam on the model's flat parameter bucket + the reference's ManualStepping schedule.

tf.keras.optimizers.Adam(learning_rate) as used at trainer.py:150 (beta_1 0.9, beta_2 0.999, epsilon 1e-7, epsilon
outside the square root, bias correction folded into the step size); ManualStepping: learning_schedules.py:19-67.
"""
from __future__ import annotations

import torch

from . import lib


class ManualStepping:
    """Piecewise-constant rate: rates[i] applies for boundaries[i-1] <= step < boundaries[i]; with warmup the first
    interval ramps linearly from rates[0] to rates[1] (learning_schedules.py:19-67)."""

    def __init__(self, boundaries, rates, warmup: bool = False):
        if any(b < 0 for b in boundaries) or any(int(b) != b for b in boundaries):
            raise ValueError("boundaries must be a list of positive integers")
        if any(b2 <= b1 for b1, b2 in zip(boundaries, boundaries[1:])):
            raise ValueError("Entries in boundaries must be strictly increasing.")
        if len(rates) != len(boundaries) + 1:
            raise ValueError("Number of provided learning rates must exceed number of boundary points by exactly 1.")
        if boundaries and boundaries[0] == 0:
            raise ValueError("First step cannot be zero.")
        self.boundaries, self.rates, self.warmup = list(boundaries), [float(r) for r in rates], warmup

    def __call__(self, step: int) -> float:
        bounds, rates = self.boundaries, self.rates
        if self.warmup and bounds:
            slope = (rates[1] - rates[0]) / bounds[0]
            ramp = [rates[0] + slope * i for i in range(bounds[0])]
            bounds = list(range(bounds[0])) + bounds
            rates = ramp + rates[1:]
        else:
            bounds = [0] + bounds
        idx = 0
        for i, b in enumerate(bounds):
            if step >= b:
                idx = i
        return rates[idx]


def learning_rate_from_config(train_config):
    """train.proto LearningRate oneof -> callable(step) (trainer.py:49-96)."""
    lr = train_config.learning_rate
    kind = lr.WhichOneof("learning_rate")
    if kind == "constant_learning_rate":
        rate = lr.constant_learning_rate.learning_rate
        return lambda step: rate
    if kind == "manual_step_learning_rate":
        cfg = lr.manual_step_learning_rate
        if not cfg.schedule:
            raise ValueError("Empty learning rate schedule.")
        return ManualStepping([s.step for s in cfg.schedule],
                              [cfg.initial_learning_rate] + [s.learning_rate for s in cfg.schedule], cfg.warmup)
    raise ValueError("Learning_rate %s not supported." % kind)


def shard_of_range(offset: int, count: int, rank: int, world: int) -> tuple[int, int]:
    """(first element, element count) of the piece of bucket range [offset, offset + count) that `rank` of `world`
    updates in fact_dp_adam_range: the range is cut into pieces of ceil(count / world) rounded up to a multiple of 8
    elements (16-byte bf16 stores); the last ranks may get a short or an empty piece.  Same arithmetic as the C side."""
    per = (-(-count // world) + 7) // 8 * 8
    lo = offset + min(per * rank, count)
    hi = min(lo + per, offset + count)
    return lo, max(hi - lo, 0)


class Adam:
    def __init__(self, model, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.model = model
        self.learning_rate = learning_rate          # float or callable(step)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.iterations = 0                         # optimizer.iterations (checkpoint step counter, trainer.py:171)
        flat = model.flat_parameters
        self.m = torch.zeros_like(flat)
        self.v = torch.zeros_like(flat)
        self._lib = lib.load()

    def current_lr(self) -> float:
        lr = self.learning_rate
        return float(lr(self.iterations)) if callable(lr) else float(lr)

    def apply_gradients(self, grad_scale: float = 1.0) -> None:
        """w -= lr_t * m / (sqrt(v) + eps) on the whole bucket, then refresh the bf16 operand copies."""
        self.begin_step()
        self.apply_range(0, self.model.flat_parameters.numel(), grad_scale)
        self.end_step()

    # The same update in pieces: begin_step(), apply_range() over disjoint slices that together cover the bucket (any
    # order), end_step().  The data-parallel trainer updates a slice as soon as its all-reduce has landed, while the
    # next slice is still on the wire.
    def begin_step(self) -> None:
        self._lr = self.current_lr()
        self.iterations += 1

    def apply_range(self, offset: int, count: int, grad_scale: float = 1.0) -> None:
        model = self.model
        flat, grad = model.flat_parameters, model.flat_gradients
        if offset % 4 or offset < 0 or offset + count > flat.numel():
            raise ValueError("slice must start on a 16-byte boundary inside the bucket")
        with torch.cuda.device(model.device):
            st = torch.cuda.current_stream(model.device).cuda_stream
            kl = model.flat_bf16_parameters      # bf16 copy of the bucket (Keras-layout operands of the dX GEMMs)
            lib.check(self._lib.fact_adam_step(
                flat.data_ptr() + 4 * offset, grad.data_ptr() + 4 * offset, self.m.data_ptr() + 4 * offset,
                self.v.data_ptr() + 4 * offset, count, self._lr, self.beta_1, self.beta_2, self.epsilon,
                self.iterations, float(grad_scale), kl.data_ptr() + 2 * offset if kl is not None else None, st),
                "fact_adam_step")

    def end_step(self) -> None:
        self.model.repack(bf16_copy_done=self.model.flat_bf16_parameters is not None)

    def dp_fused_step(self, arena, grad_scale: float = 1.0) -> None:
        """Cross-replica gradient sum + Adam + weight broadcast in ONE kernel over peer memory (fact_dp_adam_step).
        The caller brackets it with arena.barrier() and calls end_step() afterwards.  From the first call on this
        rank's m / v are meaningful only inside arena.shard() (sharded optimizer state; state_dict() reassembles)."""
        import ctypes as C
        self.begin_step()
        self._sharded = arena
        self.__dict__.pop("_shard_ranges", None)     # whole-bucket shards from here on
        with torch.cuda.device(self.model.device):
            st = torch.cuda.current_stream(self.model.device).cuda_stream
            lib.check(self._lib.fact_dp_adam_step(
                arena.peer_ptrs, arena.mc_ptr, arena.grad_off, arena.w_off, arena.wb_off, self.m.data_ptr(),
                self.v.data_ptr(), arena.numel, arena.rank, arena.world, self._lr, self.beta_1, self.beta_2,
                self.epsilon, self.iterations, float(grad_scale), st), "fact_dp_adam_step")

    def dp_fused_range(self, arena, offset: int, count: int, max_blocks: int = 0, grad_scale: float = 1.0) -> None:
        """The fused step on elements [offset, offset + count) of the bucket (fact_dp_adam_range): the data-parallel
        trainer calls it per slice, on a side stream, as the backward finishes the slices.  begin_step() first, a
        cross-replica barrier before every call, one after the last, then end_step().  This rank's m / v are meaningful
        inside its shard of every slice it has been called with."""
        self._sharded = arena
        ranges = self.__dict__.setdefault("_shard_ranges", {})
        ranges[(offset, count)] = shard_of_range(offset, count, arena.rank, arena.world)
        with torch.cuda.device(self.model.device):
            st = torch.cuda.current_stream(self.model.device).cuda_stream
            lib.check(self._lib.fact_dp_adam_range(
                arena.peer_ptrs, arena.mc_ptr, arena.grad_off, arena.w_off, arena.wb_off, self.m.data_ptr(),
                self.v.data_ptr(), offset, count, arena.rank, arena.world, self._lr, self.beta_1, self.beta_2,
                self.epsilon, self.iterations, float(grad_scale), int(max_blocks), st), "fact_dp_adam_range")

    def _mask_to_shard(self, t: torch.Tensor) -> torch.Tensor:
        out = torch.zeros_like(t)
        ranges = getattr(self, "_shard_ranges", None)
        spans = list(ranges.values()) if ranges else [self._sharded.shard()]
        for lo, cnt in spans:
            out[lo:lo + cnt] = t[lo:lo + cnt]
        return out

    def state_dict(self):
        """COLLECTIVE when the state is sharded (after dp_fused_step): every rank calls it, every rank gets the full
        moments (each shard comes from its owner: a SUM all-reduce of the masked arrays)."""
        arena = getattr(self, "_sharded", None)
        if arena is None:
            return {"iterations": self.iterations, "m": self.m, "v": self.v}
        import torch.distributed as dist
        m, v = self._mask_to_shard(self.m), self._mask_to_shard(self.v)
        dist.all_reduce(m)
        dist.all_reduce(v)
        return {"iterations": self.iterations, "m": m, "v": v}

    def load_state_dict(self, sd):
        self.iterations = int(sd["iterations"])
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
