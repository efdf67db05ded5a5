#!/usr/bin/env python
"""Generate motion with a trained FACT on B200 -- counterpart of the reference's evaluator.py (evaluator.py:28-71).

    python evaluator.py --config_path configs/fact_v5_deeper_t10_cm12.config --model_dir /tmp/fact --output_dir out/

Restores the latest checkpoint of --model_dir (random init if none), runs infer_auto_regressive(steps=1200) per clip and
writes outputs/{motion_name}_{audio_name}.npy.  Clips come from --data_npz {motion_input [C,120,225], audio_input
[C,T,35]} or from the synthetic generator; N processes (torchrun) shard the clips, no collective.
"""
import argparse
import os

import numpy as np
import torch
import torch.distributed as dist

import glob

from mint_b200 import config_util, inputs, model_builder, parallel
from mint_b200.evaluator import SingleTaskEvaluator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_path", default=config_util.DEFAULT_CONFIG)
    ap.add_argument("--model_dir", default="/tmp/fact_b200")
    ap.add_argument("--output_dir", default="outputs")
    ap.add_argument("--data_npz", default="")
    ap.add_argument("--tf_checkpoint", default="", help="TensorFlow checkpoint prefix (ckpt-N) to evaluate; default: "
                    "whichever of the newest TF checkpoint and the newest ckpt-*.pt in --model_dir has the higher step")
    ap.add_argument("--num_clips", type=int, default=4)
    ap.add_argument("--steps", type=int, default=1200)
    ap.add_argument("--mode", default="precise", choices=["precise", "bf16"])
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    cfg = config_util.get_configs_from_pipeline_file(args.config_path)
    model = model_builder.build(cfg["model"], True, device=dev, mode=args.mode)      # is_training=True: evaluator.py:56
    from mint_b200 import tf_checkpoint
    tf_prefix = args.tf_checkpoint or tf_checkpoint.latest_checkpoint(args.model_dir)   # evaluator.py:56-60 restores TF
    ckpts = sorted(f for f in os.listdir(args.model_dir) if f.startswith("ckpt-") and f.endswith(".pt")) \
        if os.path.isdir(args.model_dir) else []
    if tf_prefix and ckpts and not args.tf_checkpoint:
        # both kinds present: the newer step wins (training may have continued without --export_tf_checkpoint)
        import re
        m = re.search(r"-(\d+)$", tf_prefix)
        tf_step = int(m.group(1)) if m else -1
        pt_step = int(ckpts[-1][len("ckpt-"):-len(".pt")])
        if pt_step > tf_step:
            tf_prefix = None
    if tf_prefix:
        model.set_weights(tf_checkpoint.load_fact_weights(tf_prefix, model.dims))
    elif ckpts:
        sd = torch.load(os.path.join(args.model_dir, ckpts[-1]), map_location=dev)
        model.flat_parameters.copy_(sd["flat_parameters"])
        model.repack()
    d = model.dims
    if glob.glob(cfg["eval_dataset"].data_files):            # evaluator.py:50-54: the reference's TFRecords
        clips = [b for b in inputs.create_input(cfg["eval_config"], cfg["eval_dataset"], is_training=False)]
        mine = parallel.shard_clips(len(clips), rank, world)
        ev = SingleTaskEvaluator((clips[i] for i in mine), model, output_dir=args.output_dir, steps=args.steps)
        with torch.cuda.stream(torch.cuda.Stream(dev)):
            ev.evaluate(-1)
            torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    if args.data_npz:
        data = np.load(args.data_npz)
        motion, audio = data["motion_input"], data["audio_input"]
    else:
        rng = np.random.default_rng(0)
        motion = 0.5 * rng.standard_normal((args.num_clips, d.motion.seq_len, d.motion.feature_dim)).astype(np.float32)
        audio = rng.standard_normal((args.num_clips, d.audio.seq_len + args.steps - 1, d.audio.feature_dim)).astype(np.float32)
    bs = cfg["eval_config"].batch_size
    mine = parallel.shard_clips(motion.shape[0], rank, world)

    def batches():
        for i in range(0, len(mine), bs):
            idx = mine[i:i + bs]
            yield {"motion_input": torch.from_numpy(motion[idx]), "audio_input": torch.from_numpy(audio[idx]),
                   "motion_name": ["clip%04d" % j for j in idx], "audio_name": ["audio%04d" % j for j in idx]}

    ev = SingleTaskEvaluator(batches(), model, output_dir=args.output_dir, steps=args.steps)
    with torch.cuda.stream(torch.cuda.Stream(dev)):
        ev.evaluate(-1)
        torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
