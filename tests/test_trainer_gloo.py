"""N>1 host logic on CPU (gloo, world_size 2): the trainer's loss scaling, clip-before-reduce order and the single
SUM all-reduce of the flat gradient bucket (mint/ctl/single_task_trainer.py:157-158, 180-187); clip sharding helper."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mint_b200 import parallel
from mint_b200.trainer import SingleTaskTrainer


class _FakeModel:
    """Stands in for FACTModel: per-rank gradient = (rank + 1) * base * loss_scale."""

    def __init__(self, rank):
        self.rank = rank
        self.flat_parameters = torch.zeros(8)
        self.flat_gradients = torch.zeros(8)
        self.global_step = None
        self.device = torch.device("cpu")

    def forward_backward(self, inputs, target, loss_scale=1.0):
        assert "target" not in inputs                      # the label is popped before the call (:145)
        self.flat_gradients.copy_((self.rank + 1) * torch.arange(1.0, 9.0) * loss_scale)
        return torch.tensor(float(self.rank + 1))


class _StagedModel(_FakeModel):
    """Adds the staged-reduce contract of FACTModel: three contiguous slices covering the bucket."""

    num_gradient_stage_events = 4

    def gradient_stages(self):
        # (offset, count, event) in completion order: not in memory order, and the last event covers two ranges
        return [(3, 2, 0), (0, 3, 1), (6, 1, 2), (5, 1, 3), (7, 1, 3)]


class _FakeOpt:
    def __init__(self, model):
        self.model, self.iterations, self.seen = model, 0, None

    def apply_gradients(self, grad_scale=1.0):
        self.seen = self.model.flat_gradients.clone()
        self.iterations += 1

    def current_lr(self):
        return 1e-4


class _SlicedOpt(_FakeOpt):
    """Optimizer with the begin_step / apply_range / end_step protocol: records which slices were applied, in order,
    and a copy of each slice AS IT WAS when its update ran (it must already hold the cross-replica sum)."""

    def __init__(self, model):
        super().__init__(model)
        self.ranges, self.seen_parts, self.ended = [], {}, 0

    def begin_step(self):
        self.iterations += 1

    def apply_range(self, offset, count, grad_scale=1.0):
        self.ranges.append((offset, count))
        self.seen_parts[offset] = self.model.flat_gradients[offset:offset + count].clone()

    def end_step(self):
        self.ended += 1
        self.seen = torch.cat([self.seen_parts[o] for o in sorted(self.seen_parts)])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = _FakeModel(rank)
    opt = _FakeOpt(model)
    tr = SingleTaskTrainer([], "target", model, optimizer=opt)
    loss = tr.train_step({"motion_input": 0, "audio_input": 0, "target": 1})
    # SUM over ranks of (rank+1)/world * base = mean over ranks = 1.5 * base
    ok = torch.allclose(opt.seen, 1.5 * torch.arange(1.0, 9.0)) and float(loss) == rank + 1 and model.global_step == 1
    # staged reduce (three slices of the same bucket) must give the same sums as the single call
    m2 = _StagedModel(rank)
    o2 = _FakeOpt(m2)
    SingleTaskTrainer([], "target", m2, optimizer=o2).train_step({"motion_input": 0, "target": 1})
    ok = ok and torch.allclose(o2.seen, 1.5 * torch.arange(1.0, 9.0))
    # "adam" mode: even slices, each updated right after its own all-reduce (async_op work.wait), covering the bucket
    m6 = _FakeModel(rank)
    o6 = _SlicedOpt(m6)
    tr6 = SingleTaskTrainer([], "target", m6, optimizer=o6, overlap="adam", allreduce_chunks=3)
    assert tr6.overlap == "adam"
    tr6.train_step({"motion_input": 0, "target": 1})
    ok = ok and o6.ended == 1 and sum(c for _, c in o6.ranges) == 8 and o6.ranges == sorted(o6.ranges)
    ok = ok and torch.allclose(o6.seen, 1.5 * torch.arange(1.0, 9.0)) and m6.global_step == 1
    # clip BEFORE the sum (:180-183): each replica scales its own gradient to norm <= clip, then SUM
    m3 = _StagedModel(rank)
    o3 = _FakeOpt(m3)
    SingleTaskTrainer([], "target", m3, optimizer=o3, grad_clip_norm=1.0).train_step({"motion_input": 0, "target": 1})
    base = torch.arange(1.0, 9.0)
    per = [(r + 1) * base / world for r in range(world)]
    expect = sum(p * (1.0 / max(float(p.norm()), 1.0)) for p in per)
    ok = ok and torch.allclose(o3.seen, expect, atol=1e-6)
    # allreduce=False leaves the local gradient (bench.py's "exposed all-reduce" baseline)
    m4 = _StagedModel(rank)
    o4 = _FakeOpt(m4)
    SingleTaskTrainer([], "target", m4, optimizer=o4, allreduce=False).train_step({"motion_input": 0, "target": 1})
    ok = ok and torch.allclose(o4.seen, (rank + 1) * base / world)
    # logged loss = Keras Mean of loss / num_replicas over replicas and steps (single_task_trainer.py:157-158, 189)
    tr5 = SingleTaskTrainer([{"motion_input": 0, "target": 1}] * 2, "target", _StagedModel(rank),
                            optimizer=_FakeOpt(_StagedModel(rank)))
    logs = tr5.train(2)
    ok = ok and abs(logs["training_loss"] - 0.75) < 1e-6
    shard = parallel.shard_clips(11, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, shard)
    flat = sorted(i for s in gathered for i in s)
    ok = ok and flat == list(range(11)) and abs(len(gathered[0]) - len(gathered[1])) <= 1
    t = parallel.max_over_ranks(float(rank + 3))
    ok = ok and t == 4.0
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(30)
    assert sorted(results) == [(0, True), (1, True)]


def test_allreduce_plan_merges_touching_stages_and_keeps_the_tail_small():
    from mint_b200.trainer import plan_allreduce
    stages = [(100, 10, 0)] + [(100 - 20 * (i + 1), 20, i + 1) for i in range(5)] + [(130, 20, 6), (110, 20, 7), (150, 5, 7)]
    total = sum(c for _, c, _ in stages)
    for chunks in (1, 3, 8, 100):
        plan = plan_allreduce(stages, chunks)
        covered = sorted((o, o + c) for o, c, _ in plan)
        assert covered[0][0] == 0 and covered[-1][1] == 155 and sum(c for _, c, _ in plan) == total
        assert all(a[1] <= b[0] for a, b in zip(covered, covered[1:]))          # disjoint
        for o, c, e in plan:                                                     # a slice waits for its last member
            assert e == max(ev for so, sc, ev in stages if so >= o and so + sc <= o + c)
        assert [e for _, _, e in plan] == sorted(e for _, _, e in plan)          # issue order = completion order
        tail = [p for p in plan if p[2] == 7]
        assert sum(c for _, c, _ in tail) == 25                                  # the exposed tail is only the last stage
    assert len(plan_allreduce(stages, 100)) == len(stages)
    assert len(plan_allreduce(stages, 1)) <= 4


def test_shard_clips_single_process():
    assert parallel.shard_clips(5, 0, 1) == [0, 1, 2, 3, 4]
    assert [len(parallel.shard_clips(10, r, 4)) for r in range(4)] == [3, 3, 2, 2]
    assert parallel.max_over_ranks(2.5) == 2.5
