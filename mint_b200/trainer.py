"""SingleTaskTrainer -- torch/B200 counterpart of mint/ctl/single_task_trainer.py:28-211.

Same step semantics: pop the label, forward, `loss / num_replicas`, gradients, optional per-replica
clip_by_global_norm BEFORE the cross-replica sum (single_task_trainer.py:180-183), one SUM all-reduce of all gradients
(what MirroredStrategy does inside apply_gradients, :186-187), Adam.  One process per GPU; the all-reduce is a single
NCCL call on the model's flat gradient bucket.
"""
from __future__ import annotations

import math
import time

import torch
import torch.distributed as dist


class SingleTaskTrainer:
    def __init__(self, train_dataset, label_key, model, loss_fn=None, optimizer=None, metrics=None,
                 trainer_options=None, summary_fn=None, grad_clip_norm: float = 0.0):
        """train_dataset: iterable of dict batches (torch / numpy); label_key: 'target' (trainer.py:157)."""
        self.train_dataset = train_dataset
        self.label_key = label_key
        self.model = model
        self.optimizer = optimizer
        self.grad_clip_norm = grad_clip_norm
        self.summary_fn = summary_fn
        self.train_loss = 0.0
        self._steps = 0
        self._iter = None
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def train_loop_begin(self):
        self.train_loss, self._steps = 0.0, 0
        self._t0 = time.perf_counter()

    def train_step(self, inputs: dict) -> torch.Tensor:
        inputs = dict(inputs)
        target = inputs.pop(self.label_key)                                     # :145
        loss = self.model.forward_backward(inputs, target, loss_scale=1.0 / self.world)   # :151-158, 178
        grad_scale = 1.0
        if self.grad_clip_norm > 0:                                             # :180-183, per replica, before the sum
            grad_scale = clip_scale(self.model, self.grad_clip_norm)
            if grad_scale != 1.0:
                self.model.flat_gradients.mul_(grad_scale)
        if self.world > 1:                                                      # :186-187 (cross-replica SUM)
            dist.all_reduce(self.model.flat_gradients, op=dist.ReduceOp.SUM)
        self.optimizer.apply_gradients()
        self.model.global_step = self.optimizer.iterations
        return loss

    def train(self, num_steps: int):
        """orbit StandardTrainer.train: num_steps consecutive train_step calls on the dataset iterator."""
        if self._iter is None:
            self._iter = iter(self.train_dataset)
        self.train_loop_begin()
        total = None
        for _ in range(num_steps):
            loss = self.train_step(next(self._iter))
            total = loss if total is None else total + loss
            self._steps += 1
        return self.train_loop_end(total)

    def train_loop_end(self, total_loss=None):
        """Same scalar names as single_task_trainer.py:201-211."""
        n = max(self._steps, 1)
        mean = float(total_loss) / n if total_loss is not None else float("nan")
        dt = time.perf_counter() - self._t0
        return {"training_loss": mean, "task_loss": mean, "regularization_loss": 0.0,
                "learning_rate": self.optimizer.current_lr(), "steps_per_second": n / dt if dt > 0 else float("nan")}


def clip_scale(model, clip_norm: float) -> float:
    """tf.clip_by_global_norm factor: clip_norm / max(global_norm, clip_norm)."""
    from . import lib
    g = model.flat_gradients
    out = torch.zeros((), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        lib.check(lib.load().fact_sum_squares(g.data_ptr(), g.numel(), out.data_ptr(),
                                              torch.cuda.current_stream(g.device).cuda_stream), "fact_sum_squares")
    norm = math.sqrt(float(out))
    return clip_norm / max(norm, clip_norm)
