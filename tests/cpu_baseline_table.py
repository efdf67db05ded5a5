#!/usr/bin/env python
"""CPU baseline table of BASELINE.md section 3 on the host cores of the GPU box: the torch-CPU fp32 port of the
reference math (oracle/fact_oracle_torch.py; TensorFlow is not installable in this image).

  (1) single forward, B=1, median of 5 after 1 warm-up
  (2) infer_auto_regressive, 16 frames, B=1 -> s/frame (x1200 for a clip, linear extrapolation)
  (3) one forward + backward + Adam step at B=8
"""
import json
import os
import statistics
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import fact_oracle as O, fact_oracle_torch as OT  # noqa: E402


def best_threads(w, dims):
    one = {k: torch.from_numpy(v).float() for k, v in O.synthetic_inputs(dims, 1, seed=1).items()}
    best = (1e9, 1)
    for cand in sorted({c for c in (8, 16, 32, 64, os.cpu_count() or 1) if c <= (os.cpu_count() or 1)}):
        torch.set_num_threads(cand)
        with torch.no_grad():
            OT.call(w, dims, one)
            t0 = time.perf_counter()
            OT.call(w, dims, one)
            dt = time.perf_counter() - t0
        if dt < best[0]:
            best = (dt, cand)
    return best[1]


def main():
    dims = O.FACT_V5
    w = OT.to_torch(O.init_weights(dims, seed=0))
    threads = best_threads(w, dims)
    torch.set_num_threads(threads)
    inp = {k: torch.from_numpy(v).float() for k, v in O.synthetic_inputs(dims, 1, audio_len=dims.audio_seq + 15).items()}
    win = {"motion_input": inp["motion_input"], "audio_input": inp["audio_input"][:, :dims.audio_seq]}
    with torch.no_grad():
        OT.call(w, dims, win)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter()
            OT.call(w, dims, win)
            ts.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        OT.infer_auto_regressive(w, dims, inp, steps=16)
        ar = (time.perf_counter() - t0) / 16
    wt = OT.to_torch(O.init_weights(dims, seed=0), requires_grad=True)
    b8 = {k: torch.from_numpy(v).float() for k, v in O.synthetic_inputs(dims, 8).items()}
    opt = torch.optim.Adam(list(wt.values()), lr=1e-4, eps=1e-7)
    t0 = time.perf_counter()
    loss = OT.loss(b8["target"], OT.call(wt, dims, b8))
    loss.backward()
    opt.step()
    step = time.perf_counter() - t0
    cpu = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            cpu = next(line.split(":", 1)[1].strip() for line in f if line.startswith("model name"))
    except Exception:
        pass
    print(json.dumps({"cpu": cpu, "logical_cores": os.cpu_count(), "threads_used": threads, "torch": torch.__version__,
                      "forward_b1_s_median": statistics.median(ts), "ar_s_per_frame_b1": ar,
                      "ar_frames_per_s_b1": 1 / ar, "clip_1200_frames_s_extrapolated": 1200 * ar,
                      "train_step_b8_s": step, "train_samples_per_s_b8": 8 / step}))


if __name__ == "__main__":
    main()
