"""SingleTaskTrainer -- torch/B200 counterpart of mint/ctl/single_task_trainer.py:28-211.

Same step semantics: pop the label, forward, `loss / num_replicas`, gradients, optional per-replica
clip_by_global_norm BEFORE the cross-replica sum (single_task_trainer.py:180-183), one SUM all-reduce of all gradients
(what MirroredStrategy does inside apply_gradients, :186-187), Adam.  One process per GPU.

How the cross-replica sum happens is chosen by `overlap`:

  "fused"     (default wherever torch symmetric memory can map the replicas' buffers into each other: NVLink /
              NVSwitch boxes).  No all-reduce at all: the gradient sum, the Adam update and the mirroring of the new
              weights are ONE hand-written kernel over peer memory (fact_dp_adam_range: each rank reduces and updates a
              1/world shard -- one multimem.ld_reduce through the NVSwitch multicast per 16 bytes, or P2P loads -- and
              stores the new fp32 / bf16 weights into every replica).  It runs PER SLICE of the bucket, on a side
              stream, as the backward finishes the slices (fact_train_step's stage events; a device-side barrier over
              the replicas in front of every slice): a slice's weights are no longer read by the rest of the backward,
              so its sum + update + broadcast hide under the layers below and only the last slice (the audio
              encoder's first layer and embedding) is exposed.  The slice kernels use a small grid (one block per SM at
              most), which shares the SMs with the backward's persistent GEMMs instead of displacing their CTAs -- the
              reason overlapping NCCL with the backward did not pay.  With gradient clipping (the norm of the whole
              gradient is needed first) the step is one launch behind the backward ("fused_tail").  The optimizer
              moments are sharded (each rank keeps its shards only; Adam.state_dict() reassembles them).
  "fused_tail" the same kernel once on the whole bucket after the backward, bracketed by two barriers (A/B timing).

Otherwise the sum is ONE logical NCCL all-reduce of the model's flat gradient bucket, issued as a few contiguous slices,
hidden in one of two ways ("auto" times one all-reduce of the bucket at construction):

  "adam"      (fast links: NVLink / NVSwitch, the all-reduce is about as long as the optimizer pass).  All slices are
              issued right after the backward; the Adam update of slice i runs as soon as slice i has landed, while
              slice i + 1 is on the wire, so the update hides most of the transfer.  Measured on 8 x B200 (NV18): the
              all-reduce alone takes 1.4 ms; overlapping it with the BACKWARD instead made the step 2.1 ms longer,
              because every GEMM of the backward is a persistent one-CTA-per-SM kernel with a static tile map and the
              NCCL kernel takes SMs away from it (the displaced CTAs run as a second wave).
  "backward"  (slow links: the all-reduce is several times the optimizer pass).  Slices are reduced on a side stream in
              the order the backward finishes them (FACTModel.gradient_stages: head, cross layers top to bottom, motion
              encoder, audio encoder) behind the events fact_train_step records; only the last slice is exposed.
              Measured on a 2-GPU box whose pair ran at 75 GB/s: 6.4 ms all-reduce, 2.8 ms exposed.

Clipping needs the norm of the whole gradient before any slice may be summed, so with grad_clip_norm > 0 the
"backward" mode is not available; the clip factor is computed and applied on the device (no host synchronisation).
"""
from __future__ import annotations

import time

import torch
import torch.distributed as dist


class SingleTaskTrainer:
    def __init__(self, train_dataset, label_key, model, loss_fn=None, optimizer=None, metrics=None,
                 trainer_options=None, summary_fn=None, grad_clip_norm: float = 0.0, overlap: str = "auto",
                 allreduce: bool = True, allreduce_chunks: int = 4, sync_only: bool = False, arena=None):
        """train_dataset: iterable of dict batches (torch / numpy); label_key: 'target' (trainer.py:157).
        overlap: "auto" | "adam" | "backward" | "none" (see the module docstring).
        allreduce=False skips the cross-replica sum; sync_only=True then still makes the replicas meet once per step
        (a 4-byte all-reduce): bench.py subtracts that step from the data-parallel one, so what is left is the cost of
        moving the gradients, not the skew of replicas whose clocks differ under their power caps.
        arena: a SymmetricArena the model already lives in (another trainer of the same model made it)."""
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
        self.allreduce = allreduce
        self.chunks = max(1, int(allreduce_chunks))
        if overlap not in ("auto", "fused", "fused_tail", "adam", "backward", "none"):
            raise ValueError("overlap must be auto, fused, fused_tail, adam, backward or none")
        staged_ok = hasattr(model, "gradient_stages")
        sliced_ok = hasattr(optimizer, "apply_range")
        self.calibration_ms = None
        self.sync_only = bool(sync_only) and not allreduce
        self._sync_word = None
        self.arena, self.fused_error = arena, None
        if overlap in ("auto", "fused", "fused_tail") and self.world > 1 and allreduce \
                and hasattr(model, "adopt_symmetric") and hasattr(optimizer, "dp_fused_step"):
            if self.arena is None:
                self.arena, self.fused_error = _make_arena(model)
            if self.arena is not None and overlap == "auto":
                overlap = "fused"
        if overlap in ("fused", "fused_tail") and self.arena is None:
            overlap = "auto"                                   # symmetric memory unavailable: NCCL paths below
        if overlap == "fused" and not (staged_ok and hasattr(optimizer, "dp_fused_range")):
            overlap = "fused_tail"
        if overlap == "auto":
            overlap = "adam" if sliced_ok else ("backward" if staged_ok else "none")
            if self.world > 1 and allreduce and staged_ok and sliced_ok:
                self.calibration_ms = self._time_allreduce()
                # the optimizer pass over the bucket is ~1 ms; a transfer several times that is better hidden behind the
                # backward even though that slows the backward's persistent kernels
                overlap = "backward" if self.calibration_ms > 3.0 else "adam"
        if overlap == "backward" and not staged_ok:
            overlap = "none"
        if overlap == "adam" and not sliced_ok:
            overlap = "none"
        self.overlap = overlap
        self._comm = None
        self._events = None
        self._sumsq = None
        staged_modes = ("backward", "fused")
        self._plan = plan_allreduce(model.gradient_stages(), 2 * self.chunks if overlap == "fused" else self.chunks) \
            if overlap in staged_modes else None
        dev = getattr(model, "device", torch.device("cpu"))
        if self.world > 1 and overlap in staged_modes and dev.type == "cuda":
            self._comm = torch.cuda.Stream(dev)
            wanted = {ev for _, _, ev in self._plan[:-1]}
            self._events = [torch.cuda.Event() if i in wanted else None
                            for i in range(model.num_gradient_stage_events)]
            for e in self._events:
                if e is not None:
                    e.record(torch.cuda.current_stream(dev))     # creates the cudaEvent_t the C ABI re-records

    def _time_allreduce(self) -> float:
        """Milliseconds of one all-reduce of the gradient bucket (max over ranks), after one warm-up call."""
        g = self.model.flat_gradients
        dist.all_reduce(g)
        if g.device.type == "cuda":
            torch.cuda.synchronize(g.device)
        dist.barrier()
        t0 = time.perf_counter()
        dist.all_reduce(g)
        if g.device.type == "cuda":
            torch.cuda.synchronize(g.device)
        ms = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=g.device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        g.zero_()
        return float(ms)

    def even_slices(self):
        """The bucket in `chunks` contiguous slices of equal size (16-byte aligned): the "adam" mode's transfer units."""
        total = self.model.flat_gradients.numel()
        step = -(-total // self.chunks)
        step = (step + 7) // 8 * 8
        return [(o, min(step, total - o)) for o in range(0, total, step)]

    def train_loop_begin(self):
        self.train_loss, self._steps = 0.0, 0
        self._t0 = time.perf_counter()

    def train_step(self, inputs: dict) -> torch.Tensor:
        inputs = dict(inputs)
        target = inputs.pop(self.label_key)                                     # :145
        clip = self.grad_clip_norm > 0
        reduce = self.world > 1 and self.allreduce
        staged = reduce and self.overlap in ("backward", "fused") and not clip
        kw = {"stage_events": self._events} if staged and self._events else {}
        loss = self.model.forward_backward(inputs, target, loss_scale=1.0 / self.world, **kw)   # :151-158, 178
        grads = self.model.flat_gradients
        if clip:                                                                # :180-183, per replica, before the sum
            clip_by_global_norm_(self.model, self.grad_clip_norm, self)
        if reduce and self.overlap == "fused" and staged and self._comm is not None:
            # :186-187 + Adam + re-mirroring, fused, slice by slice under the backward (module docstring)
            cur = torch.cuda.current_stream(grads.device)
            self.optimizer.begin_step()
            sms = torch.cuda.get_device_properties(grads.device).multi_processor_count
            with torch.cuda.stream(self._comm):
                for off, cnt, ev in self._plan[:-1]:
                    self._comm.wait_event(self._events[ev])
                    self.arena.barrier()            # every replica has finished this slice (and with its weights)
                    self.optimizer.dp_fused_range(self.arena, off, cnt, max_blocks=sms)
                self._comm.wait_stream(cur)         # the last slice: behind the whole backward
                off, cnt, _ = self._plan[-1]
                self.arena.barrier()
                self.optimizer.dp_fused_range(self.arena, off, cnt)
                self.arena.barrier()                # every replica's stores into my weights have landed
            cur.wait_stream(self._comm)
            self.optimizer.end_step()
        elif reduce and self.overlap in ("fused", "fused_tail"):
            self.arena.barrier()                    # every replica's gradients are final
            self.optimizer.dp_fused_step(self.arena)
            self.arena.barrier()                    # every replica's stores into my weights have landed
            self.optimizer.end_step()
        elif staged:                                                            # :186-187 (cross-replica SUM)
            if self._comm is not None:
                cur = torch.cuda.current_stream(grads.device)
                with torch.cuda.stream(self._comm):
                    for off, cnt, ev in self._plan[:-1]:
                        self._comm.wait_event(self._events[ev])
                        dist.all_reduce(grads[off:off + cnt], op=dist.ReduceOp.SUM)
                off, cnt, _ = self._plan[-1]
                dist.all_reduce(grads[off:off + cnt], op=dist.ReduceOp.SUM)     # behind the whole backward
                cur.wait_stream(self._comm)
            else:
                for off, cnt, _ in self._plan:
                    dist.all_reduce(grads[off:off + cnt], op=dist.ReduceOp.SUM)
            self.optimizer.apply_gradients()
        elif reduce and self.overlap in ("adam", "backward") and hasattr(self.optimizer, "apply_range"):
            # all slices go out back to back; the update of a slice follows its own transfer, not the whole bucket's
            slices = self.even_slices()
            works = [dist.all_reduce(grads[o:o + c], op=dist.ReduceOp.SUM, async_op=True) for o, c in slices]
            self.optimizer.begin_step()
            for (o, c), w in zip(slices, works):
                w.wait()                                                        # stream-side wait for THIS slice only
                self.optimizer.apply_range(o, c)
            self.optimizer.end_step()
        else:
            if reduce:
                dist.all_reduce(grads, op=dist.ReduceOp.SUM)
            elif self.sync_only and self.world > 1:
                if self._sync_word is None:
                    self._sync_word = torch.zeros(1, dtype=torch.float32, device=grads.device)
                dist.all_reduce(self._sync_word)                               # replicas meet, nothing moves
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
        """Same scalar names as single_task_trainer.py:201-211.  The reference feeds `loss / num_replicas` to a Keras
        Mean metric on every replica (:157-158, :189-190); under MirroredStrategy its total and count are summed over
        replicas on read, so the logged value is the mean over replicas AND steps of the SCALED loss (= mean loss /
        num_replicas).  Reproduced literally: one all-reduce per logging loop, not per step."""
        n = max(self._steps, 1)
        if total_loss is not None and self.world > 1 and torch.is_tensor(total_loss):
            total_loss = total_loss.detach().clone() / self.world
            dist.all_reduce(total_loss, op=dist.ReduceOp.SUM)
            total_loss = total_loss / self.world
        mean = float(total_loss) / n if total_loss is not None else float("nan")
        dt = time.perf_counter() - self._t0
        return {"training_loss": mean, "task_loss": mean, "regularization_loss": 0.0,
                "learning_rate": self.optimizer.current_lr(), "steps_per_second": n / dt if dt > 0 else float("nan")}


def _make_arena(model):
    """(arena, None) with the model's buckets moved into symmetric memory on EVERY rank, or (None, reason)."""
    from .parallel import SymmetricArena
    arena, err = None, None
    try:
        arena = SymmetricArena(model.flat_parameters.numel(), model.device)
    except Exception as exc:                                   # no symmetric memory on this box / build: NCCL instead
        err = repr(exc)[:300]
    ok = torch.tensor([0 if arena is None else 1], dtype=torch.int32, device=model.device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)                  # all or nobody
    if int(ok) == 0:
        return None, err or "symmetric memory unavailable on another rank"
    model.adopt_symmetric(arena)
    arena.barrier()
    return arena, None


def plan_allreduce(stages, chunks: int):
    """Merge the backward's stages [(offset, count, event)] (completion order) into at most ~`chunks` contiguous slices
    [(offset, count, event)] of similar size; a merged slice waits for the LAST event of its members.  Only slices that
    touch in memory are merged, and the slices whose event is the final one stay apart from the rest (they are the
    exposed tail, so they are kept as small as the stages allow) -- but are merged with each other when adjacent."""
    total = sum(c for _, c, _ in stages)
    last_ev = stages[-1][2]
    target = total / max(1, chunks)
    out = []
    for off, cnt, ev in stages:
        if out:
            o, c, e = out[-1]
            touching = o + c == off or off + cnt == o
            same_side = (e == last_ev) == (ev == last_ev)
            if touching and same_side and (c < target or ev == last_ev):
                out[-1] = (min(o, off), c + cnt, max(e, ev))
                continue
        out.append((off, cnt, ev))
    # the final-event slices go last, as ONE call when they touch, else in bucket order
    tail = sorted(s for s in out if s[2] == last_ev)
    head = [s for s in out if s[2] != last_ev]
    return head + tail


def clip_by_global_norm_(model, clip_norm: float, owner=None) -> None:
    """tf.clip_by_global_norm in place on the flat gradient bucket: g *= clip_norm / max(||g||, clip_norm).  The squared
    norm stays in device memory (fact_sum_squares -> fact_clip_scale): no host round trip in the step."""
    g = model.flat_gradients
    if g.device.type != "cuda":                      # host-logic tests (gloo): same arithmetic in torch
        norm = g.double().pow(2).sum().sqrt()
        g.mul_(float(clip_norm) / max(float(norm), float(clip_norm)))
        return
    from . import lib
    buf = getattr(owner, "_sumsq", None)
    if buf is None:
        buf = torch.zeros((), dtype=torch.float32, device=g.device)
        if owner is not None:
            owner._sumsq = buf
    with torch.cuda.device(g.device):
        st = torch.cuda.current_stream(g.device).cuda_stream
        lib.check(lib.load().fact_sum_squares(g.data_ptr(), g.numel(), buf.data_ptr(), st), "fact_sum_squares")
        lib.check(lib.load().fact_clip_scale(g.data_ptr(), g.numel(), buf.data_ptr(), float(clip_norm), st),
                  "fact_clip_scale")


def clip_scale(model, clip_norm: float) -> float:
    """The clip factor as a host float (diagnostics / tests; synchronises)."""
    from . import lib
    g = model.flat_gradients
    out = torch.zeros((), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        lib.check(lib.load().fact_sum_squares(g.data_ptr(), g.numel(), out.data_ptr(),
                                              torch.cuda.current_stream(g.device).cuda_stream), "fact_sum_squares")
    norm = float(out) ** 0.5
    return clip_norm / max(norm, clip_norm)
