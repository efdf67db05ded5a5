"""Shared helpers for the parity tests (oracle side is test infrastructure only)."""
import numpy as np
import torch

from mint_b200 import protos
from oracle import fact_oracle as O


def make_config(d=800, heads=10, ff=3072, layers=(2, 2, 12), motion_seq=120, audio_seq=240, motion_dim=225,
                out_dim=225):
    """A mint.protos.FACTModel message with the given dims (fact_v5 by default)."""
    cfg = protos.FACTModel()
    for name, seq, fdim, nl in (("audio", audio_seq, None, layers[1]), ("motion", motion_seq, motion_dim, layers[0])):
        m = cfg.modality.add()
        m.feature_name, m.sequence_length = name, seq
        if fdim is not None:
            m.feature_dim = fdim
        t = m.model.add().transformer
        t.hidden_size, t.num_attention_heads, t.num_hidden_layers, t.intermediate_size = d, heads, nl, ff
    cm = cfg.cross_modal_model
    cm.modality_a, cm.modality_b = "motion", "audio"
    t = cm.transformer
    t.hidden_size, t.num_attention_heads, t.num_hidden_layers, t.intermediate_size = d, heads, layers[2], ff
    cm.output_layer.out_dim = out_dim
    return cfg


def oracle_dims(d=800, heads=10, ff=3072, layers=(2, 2, 12), motion_seq=120, audio_seq=240, motion_dim=225,
                audio_dim=35, out_dim=225):
    return O.Dims(d, heads, ff, layers[0], layers[1], layers[2], motion_seq, audio_seq, motion_dim, audio_dim, out_dim)


def split_ref(x: torch.Tensor):
    """bf16 hi/lo split as the kernels do it."""
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi, lo


def join(hi: torch.Tensor, lo: torch.Tensor | None):
    return hi.float() + (lo.float() if lo is not None else 0.0)


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
