"""A/B timing of the attention-core backward kernels at the training shape (B = 128, N = 360, 10 heads x 80)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mint_b200 import lib as L
lib = L.load()
dev = torch.device("cuda", 0)
batch, n, heads, dh = 128, 360, 10, 80
d = heads * dh
BF = torch.bfloat16
qkv = (torch.randn(batch * n, 3 * d, device=dev) * 0.5).to(BF)
o = torch.zeros(batch * n, d, device=dev, dtype=BF)
lse = torch.zeros(batch, heads, n, device=dev)
st = torch.cuda.current_stream().cuda_stream
L.check(lib.fact_sdpa_lse(qkv.data_ptr(), None, o.data_ptr(), None, lse.data_ptr(), batch, n, heads, dh, st))
d_o = torch.randn(batch * n, d, device=dev).to(BF)
dqkv = torch.zeros(batch * n, 3 * d, device=dev, dtype=BF)
dscr = torch.zeros(batch * heads * n, device=dev)
dq = torch.zeros(batch * n * d, device=dev)
res = {}
for flag in (1, 2, 1, 2):
    lib.fact_set_flag(b"sdpa_bwd_tc", flag)
    def run():
        L.check(lib.fact_sdpa_backward(qkv.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(), dscr.data_ptr(),
                                       dq.data_ptr(), dqkv.data_ptr(), batch, n, heads, dh, d ** -0.5, st))
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    print("sdpa_bwd_tc =", flag, "us per call (prep + core + finish + memset):", round(e0.elapsed_time(e1) * 100, 1), flush=True)
    res[flag] = dqkv.float().clone()
print("max abs diff pipelined vs first generation:", float((res[1] - res[2]).abs().max()), "of", float(res[2].abs().max()))
