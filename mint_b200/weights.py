"""Variable naming and Keras-default initialisation of FACTModel's weights.

Flat dict of fp32 tensors in Keras layout (Dense `kernel` is [in, out]); names follow the reference's
sub-layer tree (mint/core/fact_model.py:43-70, mint/core/base_models.py:22-202):

  cross_modal_layer/transformer/layer_{i}/attn/{norm/gamma,norm/beta,to_qkv/kernel,to_out/kernel,to_out/bias}
  cross_modal_layer/transformer/layer_{i}/mlp/{norm/gamma,norm/beta,dense_0/kernel,dense_0/bias,dense_1/kernel,dense_1/bias}
  cross_modal_layer/output/{kernel,bias}
  {motion,audio}_transformer/layer_{i}/...   {motion,audio}_pos_embedding   {motion,audio}_linear_embedding/{kernel,bias}
"""
from __future__ import annotations

import math

import torch

from .config_util import FactDims

LAYER_SUFFIXES = (
    "attn/norm/gamma", "attn/norm/beta", "attn/to_qkv/kernel", "attn/to_out/kernel", "attn/to_out/bias",
    "mlp/norm/gamma", "mlp/norm/beta", "mlp/dense_0/kernel", "mlp/dense_0/bias", "mlp/dense_1/kernel",
    "mlp/dense_1/bias",
)


def _layer_shapes(d: int, ff: int):
    return ((d,), (d,), (d, 3 * d), (d, d), (d,), (d,), (d,), (d, ff), (ff,), (ff, d), (d,))


def variable_shapes(dims: FactDims) -> dict:
    """name -> shape, in the reference's creation order (cross-modal layer first, fact_model.py:43-70)."""
    out = {}
    d = dims.cross_hidden

    def stack(prefix, layers, dd, ff):
        for i in range(layers):
            for suf, shp in zip(LAYER_SUFFIXES, _layer_shapes(dd, ff)):
                out[f"{prefix}/layer_{i}/{suf}"] = shp

    stack("cross_modal_layer/transformer", dims.cross_layers, d, dims.cross_ff)
    out["cross_modal_layer/output/kernel"] = (d, dims.out_dim)
    out["cross_modal_layer/output/bias"] = (dims.out_dim,)
    for name, enc in (("motion", dims.motion), ("audio", dims.audio)):
        stack(f"{name}_transformer", enc.layers, enc.hidden, enc.ff)
        out[f"{name}_pos_embedding"] = (enc.seq_len, enc.hidden)
        out[f"{name}_linear_embedding/kernel"] = (enc.feature_dim, enc.hidden)
        out[f"{name}_linear_embedding/bias"] = (enc.hidden,)
    return out


def keras_default_init(dims: FactDims, seed: int = 0) -> dict:
    """Keras defaults: Dense kernels glorot-uniform, biases zero, LayerNorm gamma 1 / beta 0; position tables and the
    output kernel TruncatedNormal(stddev) resampled beyond 2 sigma (mint/core/base_model_util.py:89-91,
    base_models.py:148-152, 176-180).  The Transformer ignores its initializer_range (base_models.py:94-108)."""
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    w = {}
    for name, shp in variable_shapes(dims).items():
        if name.endswith("pos_embedding") or name == "cross_modal_layer/output/kernel":
            std = dims.out_init_range if name.endswith("kernel") else 0.02
            t = torch.empty(shp, dtype=torch.float32)
            torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
            w[name] = t
        elif name.endswith("/kernel"):
            lim = math.sqrt(6.0 / (shp[0] + shp[1]))
            w[name] = (torch.rand(shp, generator=gen, dtype=torch.float32) * 2 - 1) * lim
        elif name.endswith("/gamma"):
            w[name] = torch.ones(shp, dtype=torch.float32)
        else:
            w[name] = torch.zeros(shp, dtype=torch.float32)
    return w


GEMM_KERNELS = ("attn/to_qkv/kernel", "attn/to_out/kernel", "mlp/dense_0/kernel", "mlp/dense_1/kernel")
