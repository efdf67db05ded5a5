#!/usr/bin/env python
"""BASELINE.json configs[2]/[3]: one data-parallel training step (bf16 products, L2 motion loss, Keras Adam) at
per-GPU batch B; N>1: launch under torchrun, ONE NCCL all-reduce of the flat fp32 gradient bucket per step.

    python scripts/bench_train.py [--batch 128] [--steps 6] [--warmup 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py ...
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mint_b200 import config_util, model_builder, optim  # noqa: E402
from mint_b200.trainer import SingleTaskTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG)
    model = model_builder.build(cfg["model"], is_training=True, device=dev, mode="bf16", seed=0)
    d = model.dims
    opt = optim.Adam(model, learning_rate=optim.learning_rate_from_config(cfg["train_config"]))
    B = args.batch
    g = torch.Generator(device="cpu").manual_seed(rank)
    batch = {"motion_input": (0.5 * torch.randn(B, d.motion.seq_len, d.motion.feature_dim, generator=g)).to(dev),
             "audio_input": torch.randn(B, d.audio.seq_len, d.audio.feature_dim, generator=g).to(dev),
             "target": (0.5 * torch.randn(B, 20, d.out_dim, generator=g)).to(dev)}
    trainer = SingleTaskTrainer([batch] * (args.steps + args.warmup), "target", model, optimizer=opt)
    stream = torch.cuda.Stream(dev)
    with torch.cuda.stream(stream):
        losses = [float(trainer.train_step(batch)) for _ in range(args.warmup)]
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        # all-reduce alone (the one collective of the step)
        ar_ms = 0.0
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                dist.all_reduce(model.flat_gradients)
            e1.record(stream)
            torch.cuda.synchronize(dev)
            ar_ms = e0.elapsed_time(e1) / 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            loss = trainer.train_step(batch)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
    if rank == 0:
        flops = 3 * 80.97e9 * B * world
        print(json.dumps({
            "metric": "training step time", "config": "fact_v5_deeper_t10_cm12, bf16 products, L2 loss, Keras Adam",
            "n_gpus": world, "batch_per_gpu": B, "ms_per_step": ms, "samples_per_s": world * B * 1e3 / ms,
            "algorithmic_tflops_per_gpu": flops / world / (ms * 1e-3) / 1e12, "allreduce_ms": ar_ms,
            "grad_bucket_mb": model.flat_gradients.numel() * 4 / 1e6, "loss_first": losses[0], "loss_last": float(loss),
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
