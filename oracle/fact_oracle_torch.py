"""Second, independent CPU restatement of the FACT hot path in torch (fp32 by default) -- TEST INFRASTRUCTURE.

Parity pin: see fact_oracle.py (pinned to the reference's own model code run over a NumPy shim of TF; real-TF
numerics unpinned).  Written separately from the NumPy oracle, against the same reference
lines, so the two can cross-check each other; it is also the "port" CPU baseline that bench.py times (the
reference's TF-CPU path cannot run: TensorFlow is absent from the image) and, being differentiable, the
autograd ground truth for the backward kernels.

Reference: mint/core/fact_model.py:72-148, mint/core/base_models.py:22-202, mint/core/base_model_util.py:94-107.
Weights: the flat Keras-layout dict of fact_oracle.weight_shapes().
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def to_torch(weights: dict, dtype=torch.float32, requires_grad: bool = False) -> dict:
    out = {}
    for k, v in weights.items():
        t = torch.as_tensor(v).to(dtype).clone()
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def _gelu(x):
    # base_model_util.py:94-107
    return x * (0.5 * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x * x * x))))


def _block(x, w, p, heads):
    # Residual(Norm(Attention)) -- base_models.py:22-42, 60-88
    d = x.shape[-1]
    h = F.layer_norm(x, (d,), w[p + "/attn/norm/gamma"], w[p + "/attn/norm/beta"], 1e-5)
    qkv = h.matmul(w[p + "/attn/to_qkv/kernel"])
    q, k, v = qkv.split(d, dim=-1)                       # (qkv h d) column order: q | k | v
    split = lambda t: t.unflatten(-1, (heads, d // heads)).transpose(1, 2)   # b h n dh
    q, k, v = split(q), split(k), split(v)
    p_attn = torch.softmax(q.matmul(k.transpose(-1, -2)) * (float(d) ** -0.5), dim=-1)
    o = p_attn.matmul(v).transpose(1, 2).flatten(-2)      # b n (h dh)
    x = x + o.matmul(w[p + "/attn/to_out/kernel"]) + w[p + "/attn/to_out/bias"]
    # Residual(Norm(MLP)) -- base_models.py:45-57
    h = F.layer_norm(x, (d,), w[p + "/mlp/norm/gamma"], w[p + "/mlp/norm/beta"], 1e-5)
    h = _gelu(h.matmul(w[p + "/mlp/dense_0/kernel"]) + w[p + "/mlp/dense_0/bias"])
    return x + h.matmul(w[p + "/mlp/dense_1/kernel"]) + w[p + "/mlp/dense_1/bias"]


# ---- the same math with the product path's bf16 OPERAND rounding (mint_b200/csrc/engine_train.cu), for gradient parity:
# every tensor-core product reads bf16 copies of both operands, forward and backward (the incoming gradient is rounded
# too), accumulation and everything else stay in the working precision (fp64 here).  Comparing the CUDA gradients with
# autograd through THIS graph isolates accumulation / kernel errors from the bf16 rounding the configuration asks for.
def _r(t):
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundST(torch.autograd.Function):
    """Round to bf16, identity gradient (a stored bf16 activation)."""

    @staticmethod
    def forward(ctx, x):
        return _r(x)

    @staticmethod
    def backward(ctx, g):
        return g


class _QMatmul(torch.autograd.Function):
    """y = bf16(a) @ bf16(b); backward: da = bf16(dy) @ bf16(b)^T, db = bf16(a)^T @ bf16(dy)."""

    @staticmethod
    def forward(ctx, a, b):
        ar, br = _r(a), _r(b)
        ctx.save_for_backward(ar, br)
        return ar.matmul(br)

    @staticmethod
    def backward(ctx, g):
        ar, br = ctx.saved_tensors
        gr = _r(g)
        da = gr.matmul(br.transpose(-1, -2))
        db = ar.transpose(-1, -2).matmul(gr)
        while db.dim() > br.dim():            # a batched left operand against a 2-D weight: sum over the batch
            db = db.sum(0)
        return da, db


class _GeluSavedZ(torch.autograd.Function):
    """h = gelu(z); the backward evaluates gelu' at the bf16 copy of z the forward saved (FACT_EPI_GELU_GRAD)."""

    @staticmethod
    def forward(ctx, z):
        ctx.save_for_backward(_r(z))
        return _gelu(z)

    @staticmethod
    def backward(ctx, g):
        (z,) = ctx.saved_tensors
        c = math.sqrt(2.0 / math.pi)
        u = c * (z + 0.044715 * z * z * z)
        t = torch.tanh(u)
        return g * (0.5 * (1.0 + t) + 0.5 * z * (1.0 - t * t) * c * (1.0 + 3 * 0.044715 * z * z))


def _block_bf16(x, w, p, heads):
    d = x.shape[-1]
    qm = _QMatmul.apply
    h = F.layer_norm(x, (d,), w[p + "/attn/norm/gamma"], w[p + "/attn/norm/beta"], 1e-5)
    qkv = qm(h, w[p + "/attn/to_qkv/kernel"])
    q, k, v = qkv.split(d, dim=-1)
    # the QKV epilogue stores bf16 q * (d^-0.5 * log2 e), bf16 k, bf16 v; scores live in the log2 domain
    q = _RoundST.apply(q * (float(d) ** -0.5 * 1.4426950408889634))
    k, v = _RoundST.apply(k), _RoundST.apply(v)
    split = lambda t: t.unflatten(-1, (heads, d // heads)).transpose(1, 2)
    q, k, v = split(q), split(k), split(v)
    s = qm(q, k.transpose(-1, -2)) * 0.6931471805599453          # back to the natural-log domain for softmax
    o = qm(torch.softmax(s, dim=-1), v).transpose(1, 2).flatten(-2)
    x = x + qm(_RoundST.apply(o), w[p + "/attn/to_out/kernel"]) + w[p + "/attn/to_out/bias"]
    h = F.layer_norm(x, (d,), w[p + "/mlp/norm/gamma"], w[p + "/mlp/norm/beta"], 1e-5)
    h = _GeluSavedZ.apply(qm(h, w[p + "/mlp/dense_0/kernel"]) + w[p + "/mlp/dense_0/bias"])
    return x + qm(h, w[p + "/mlp/dense_1/kernel"]) + w[p + "/mlp/dense_1/bias"]


def _stack(x, w, prefix, layers, heads, block=_block):
    for i in range(layers):
        x = block(x, w, f"{prefix}/layer_{i}", heads)
    return x


def call(w: dict, dims, inputs: dict, bf16_operands: bool = False):
    """FACTModel.call, fact_model.py:72-101.  bf16_operands=True: same graph with the product path's operand rounding
    (see _QMatmul) -- the gradient-parity reference of tests/test_train_gpu.py, never a parity source for `call`."""
    ref = next(iter(w.values()))
    mo = torch.as_tensor(inputs["motion_input"]).to(ref.dtype)
    au = torch.as_tensor(inputs["audio_input"]).to(ref.dtype)
    assert mo.shape[1] == dims.motion_seq and au.shape[1] == dims.audio_seq
    mm = _QMatmul.apply if bf16_operands else torch.matmul
    block = _block_bf16 if bf16_operands else _block
    m = mm(mo, w["motion_linear_embedding/kernel"]) + w["motion_linear_embedding/bias"] + w["motion_pos_embedding"]
    m = _stack(m, w, "motion_transformer", dims.motion_layers, dims.heads, block)
    a = mm(au, w["audio_linear_embedding/kernel"]) + w["audio_linear_embedding/bias"] + w["audio_pos_embedding"]
    a = _stack(a, w, "audio_transformer", dims.audio_layers, dims.heads, block)
    x = _stack(torch.cat([m, a], dim=1), w, "cross_modal_layer/transformer", dims.cross_layers, dims.heads, block)
    return mm(x, w["cross_modal_layer/output/kernel"]) + w["cross_modal_layer/output/bias"]


@torch.no_grad()
def infer_auto_regressive(w: dict, dims, inputs: dict, steps: int = 1200):
    """fact_model.py:103-132."""
    ref = next(iter(w.values()))
    motion = torch.as_tensor(inputs["motion_input"]).to(ref.dtype)
    audio_all = torch.as_tensor(inputs["audio_input"]).to(ref.dtype)
    frames = []
    for i in range(steps):
        window = audio_all[:, i:i + dims.audio_seq]
        if window.shape[1] < dims.audio_seq:
            break
        first = call(w, dims, {"motion_input": motion, "audio_input": window})[:, :1]
        frames.append(first)
        motion = torch.cat([motion[:, 1:], first], dim=1)
    return torch.cat(frames, dim=1)


def loss(target, pred):
    """fact_model.py:134-148."""
    target = torch.as_tensor(target).to(pred.dtype)
    return ((target - pred[:, :target.shape[1]]) ** 2).mean()
