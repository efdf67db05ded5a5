#!/usr/bin/env python
"""Where does the end-to-end call (host tensors in, host frames out) spend its time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mint_b200 import config_util, model_builder

dev = torch.device("cuda", 0)
cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG)
model = model_builder.build(cfg["model"], is_training=False, device=dev, mode="precise")
d = model.dims
B, K = 128, 12
motion_h = (0.5 * torch.randn(B, d.motion.seq_len, d.motion.feature_dim)).pin_memory()
audio_h = torch.randn(B, d.audio.seq_len + K - 1, d.audio.feature_dim).pin_memory()
out_host = torch.empty(B, K, d.out_dim).pin_memory()
stream = torch.cuda.Stream(dev)
def sync():
    torch.cuda.synchronize(dev)
with torch.cuda.stream(stream):
    for it in range(4):
        sync(); t0 = time.perf_counter()
        frames = model.infer_auto_regressive({"motion_input": motion_h, "audio_input": audio_h}, steps=K)
        t_call = time.perf_counter() - t0
        out_host.copy_(frames, non_blocking=True)
        sync(); t_all = time.perf_counter() - t0
        print(f"call {it}: host returned after {t_call*1e3:.2f} ms, complete after {t_all*1e3:.2f} ms")
    # phases
    sync(); t0 = time.perf_counter()
    m = model._to_dev(motion_h, d.motion.feature_dim, "motion_input"); a = model._to_dev(audio_h, d.audio.feature_dim, "audio_input")
    sync(); t1 = time.perf_counter()
    hist = model.new_history(m, K)
    sync(); t2 = time.perf_counter()
    model.generate_into(hist, a, 0, K)
    t3h = time.perf_counter(); sync(); t3 = time.perf_counter()
    fr = hist[:, d.motion.seq_len:].contiguous(); out_host.copy_(fr, non_blocking=True)
    sync(); t4 = time.perf_counter()
    print(f"h2d {1e3*(t1-t0):.2f}  new_history {1e3*(t2-t1):.2f}  generate host {1e3*(t3h-t2):.2f} total {1e3*(t3-t2):.2f}  d2h {1e3*(t4-t3):.2f} ms")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream); model.generate_into(hist, a, 0, K); e1.record(stream); sync()
    print(f"generate device time {e0.elapsed_time(e1):.2f} ms")
