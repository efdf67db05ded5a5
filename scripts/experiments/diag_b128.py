"""Diagnostic: gradients of k-times duplicated clips vs the 2 clips alone, per tensor (which batch size / kernel path
introduces differences beyond fp32 summation order)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from mint_b200.fact_model import FACTModel
from oracle import fact_oracle as O
from tests.helpers import make_config, oracle_dims

dims = oracle_dims()
w = O.init_weights(dims, seed=2)
two = O.synthetic_inputs(dims, batch=2, seed=11)
t2 = {k: torch.from_numpy(v).float() for k, v in two.items()}
m = FACTModel(make_config(), is_training=True, mode="bf16")
m.set_weights(w)
m.forward_backward(t2, t2["target"])
g2 = m.flat_gradients.clone()
for B in (4, 8, 16, 32, 128):
    idx = torch.arange(B) % 2
    big = {k: v[idx].contiguous() for k, v in t2.items()}
    m.forward_backward(big, big["target"])
    g = m.flat_gradients
    errs = []
    for name, (off, cnt) in m._offsets.items():
        a, b = g[off:off + cnt].double(), g2[off:off + cnt].double()
        nb = float(b.norm())
        if nb > 1e-12:
            errs.append((float((a - b).norm()) / nb, name))
    errs.sort(reverse=True)
    print(B, "worst:", [(round(e, 6), n) for e, n in errs[:4]], "median:", round(errs[len(errs) // 2][0], 7), flush=True)
