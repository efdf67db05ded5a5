#!/usr/bin/env python
"""BASELINE.json configs[1]/[4]: AR frames/s vs batch on one B200 (device-resident inputs, CUDA events).

    python scripts/sweep_batch.py [--modes precise,bf16] [--batches 1,2,4,...,512] [--steps 8] [--out FILE]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from mint_b200 import config_util, model_builder  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--modes", default="precise,bf16")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG)
    stream = torch.cuda.Stream(dev)
    rows = []
    for mode in args.modes.split(","):
        model = model_builder.build(cfg["model"], is_training=False, device=dev, mode=mode)
        d = model.dims
        for b in [int(x) for x in args.batches.split(",")]:
            k, w = args.steps, args.warmup
            motion = 0.5 * torch.randn(b, d.motion.seq_len, d.motion.feature_dim, device=dev)
            audio = torch.randn(b, d.audio.seq_len + w + k - 1, d.audio.feature_dim, device=dev)
            with torch.cuda.stream(stream):
                hist = model.new_history(motion, w + k)
                model.generate_into(hist, audio, 0, w)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                stream.synchronize()
                e0.record(stream)
                model.generate_into(hist, audio, w, k)
                e1.record(stream)
                stream.synchronize()
            ms = e0.elapsed_time(e1) / k
            rows.append({"mode": mode, "batch": b, "ms_per_frame_step": ms, "frames_per_s": b * 1e3 / ms,
                         "algorithmic_tflops": b * 80.97e9 / (ms * 1e-3) / 1e12})
            print(json.dumps(rows[-1]), flush=True)
            del hist
        del model
        torch.cuda.empty_cache()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
