"""CPU oracle for the FACT hot path -- TEST INFRASTRUCTURE ONLY.

PARITY PIN: the reference ships no numeric golden vector for this path (its tests assert shapes only:
mint/core/fact_model_test.py:23-54, base_models_test.py:20-40) and TensorFlow is not installable in this image, so
real-TF outputs cannot be produced.  What IS pinned: tests/golden/fact_reference_code_small.npz holds outputs of the
reference's OWN model code -- /root/reference/mint/core/{fact_model,base_models,base_model_util,model_builder}.py
imported in the build container with `tensorflow` resolving to oracle/tf_shim (NumPy semantics of the ~12 TF/Keras
primitives those files call) -- for FACTModel.call, .infer_auto_regressive (incl. the early stop) and .loss; this
restatement reproduces them to 1e-11 in float64 (tests/test_oracle.py).  So the COMPOSITION (layer order, einsum and
Rearrange patterns, the dim**-0.5 scale, concat order, residual wiring, kept row, shift) is pinned to the reference's
code; the primitives (Dense = x@W+b, LayerNormalization eps-inside-sqrt biased variance, softmax, tanh) carry their
documented Keras semantics and remain UNPINNED against real TensorFlow numerics.  Further checks: an independent
second restatement (`fact_oracle_torch.py`), the reference's own shape tests, and structural self-checks
(tests/test_oracle.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (mint_b200/) never does and fails loudly without its CUDA library.

Everything is float64 NumPy unless `dtype` is given.  Weight container: flat dict, Keras layout
(`kernel` is [in, out] so y = x @ kernel + bias), names in `weight_names()`.

Reference files restated (read for behaviour; no code copied):
  mint/core/fact_model.py:72-101   FACTModel.call
  mint/core/fact_model.py:103-132  FACTModel.infer_auto_regressive
  mint/core/fact_model.py:134-148  FACTModel.loss / compute_motion_generation_loss
  mint/core/base_models.py:22-31   Norm (LayerNormalization eps 1e-5, pre-norm)
  mint/core/base_models.py:34-42   Residual
  mint/core/base_models.py:45-57   MLP
  mint/core/base_models.py:60-88   Attention (fused qkv, scale = dim**-0.5, no mask)
  mint/core/base_models.py:91-110  Transformer (no final LN)
  mint/core/base_models.py:130-156 LinearEmbedding, PositionEmbedding
  mint/core/base_models.py:159-202 CrossModalLayer
  mint/core/base_model_util.py:94-107 gelu (tanh form)
"""
from __future__ import annotations

import math

import numpy as np

EPS_LN = 1e-5  # base_models.py:27


# --------------------------------------------------------------------------- dims / names
class Dims:
    """Plain shape record (the oracle does not depend on the product's config code)."""

    def __init__(self, d=800, heads=10, ff=3072, motion_layers=2, audio_layers=2, cross_layers=12,
                 motion_seq=120, audio_seq=240, motion_dim=225, audio_dim=35, out_dim=225):
        self.d, self.heads, self.ff = d, heads, ff
        self.motion_layers, self.audio_layers, self.cross_layers = motion_layers, audio_layers, cross_layers
        self.motion_seq, self.audio_seq = motion_seq, audio_seq
        self.motion_dim, self.audio_dim, self.out_dim = motion_dim, audio_dim, out_dim


FACT_V5 = Dims()


def layer_names(prefix: str):
    return [
        f"{prefix}/attn/norm/gamma", f"{prefix}/attn/norm/beta",
        f"{prefix}/attn/to_qkv/kernel",
        f"{prefix}/attn/to_out/kernel", f"{prefix}/attn/to_out/bias",
        f"{prefix}/mlp/norm/gamma", f"{prefix}/mlp/norm/beta",
        f"{prefix}/mlp/dense_0/kernel", f"{prefix}/mlp/dense_0/bias",
        f"{prefix}/mlp/dense_1/kernel", f"{prefix}/mlp/dense_1/bias",
    ]


def weight_shapes(dims: Dims) -> dict:
    """name -> shape, in the reference's sub-layer creation order (fact_model.py:43-70)."""
    d, ff = dims.d, dims.ff
    out = {}

    def layer(prefix):
        shp = [(d,), (d,), (d, 3 * d), (d, d), (d,), (d,), (d,), (d, ff), (ff,), (ff, d), (d,)]
        for n, s in zip(layer_names(prefix), shp):
            out[n] = s

    for i in range(dims.cross_layers):
        layer(f"cross_modal_layer/transformer/layer_{i}")
    out["cross_modal_layer/output/kernel"] = (d, dims.out_dim)
    out["cross_modal_layer/output/bias"] = (dims.out_dim,)
    for i in range(dims.motion_layers):
        layer(f"motion_transformer/layer_{i}")
    out["motion_pos_embedding"] = (dims.motion_seq, d)
    out["motion_linear_embedding/kernel"] = (dims.motion_dim, d)
    out["motion_linear_embedding/bias"] = (d,)
    for i in range(dims.audio_layers):
        layer(f"audio_transformer/layer_{i}")
    out["audio_pos_embedding"] = (dims.audio_seq, d)
    out["audio_linear_embedding/kernel"] = (dims.audio_dim, d)
    out["audio_linear_embedding/bias"] = (d,)
    return out


def init_weights(dims: Dims, seed: int = 0, randomize_affine: bool = False) -> dict:
    """Keras-default initialisation, deterministic in `seed` (NumPy PCG64).

    Dense kernels glorot-uniform, biases 0, LN gamma 1 / beta 0 (Keras defaults, base_models.py:27,51-53,
    68-69,135); position tables and the output kernel TruncatedNormal(0.02) re-sampled outside +-2 sigma
    (base_model_util.py:89-91; base_models.py:148-152,176-180).
    `randomize_affine=True` perturbs biases / LN affine so tests exercise them (they are all-trivial at init).
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    w = {}
    for name, shp in weight_shapes(dims).items():
        if name.endswith("pos_embedding") or name == "cross_modal_layer/output/kernel":
            a = rng.standard_normal(shp)
            bad = np.abs(a) > 2.0
            while bad.any():
                a[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(a) > 2.0
            w[name] = 0.02 * a
        elif name.endswith("/kernel"):
            lim = math.sqrt(6.0 / (shp[0] + shp[1]))
            w[name] = rng.uniform(-lim, lim, shp)
        elif name.endswith("/gamma"):
            w[name] = np.ones(shp) + (0.1 * rng.standard_normal(shp) if randomize_affine else 0.0)
        else:  # bias / beta
            w[name] = 0.05 * rng.standard_normal(shp) if randomize_affine else np.zeros(shp)
    return w


def synthetic_inputs(dims: Dims, batch: int, audio_len: int | None = None, seed: int = 0,
                     target_len: int = 20) -> dict:
    """SURVEY.md 8(d) synthetic tensors: motion 0.5*N(0,1) with 6 leading zero dims, audio N(0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000))
    T = dims.audio_seq if audio_len is None else audio_len
    motion = 0.5 * rng.standard_normal((batch, dims.motion_seq, dims.motion_dim))
    motion[..., :6] = 0.0  # the 219 -> 225 zero padding (inputs_util.py:70-73)
    audio = rng.standard_normal((batch, T, dims.audio_dim))
    target = 0.5 * rng.standard_normal((batch, target_len, dims.motion_dim))
    target[..., :6] = 0.0
    return {"motion_input": motion, "audio_input": audio, "target": target}


# --------------------------------------------------------------------------- blocks
def layer_norm(x, gamma, beta):
    """base_models.py:27-31 -- biased variance over the last axis, eps inside the sqrt."""
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + EPS_LN) * gamma + beta


def gelu_tanh(x):
    """base_model_util.py:94-107."""
    return x * 0.5 * (1.0 + np.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def attention(x, wqkv, wo, bo, heads):
    """base_models.py:60-88.  qkv columns are ordered (qkv, head, dh): einops
    "b n (qkv h d) -> qkv b h n d"; scale is d_model**-0.5 (NOT head_dim**-0.5)."""
    b, n, d = x.shape
    dh = d // heads
    qkv = (x @ wqkv).reshape(b, n, 3, heads, dh)
    q = qkv[:, :, 0].transpose(0, 2, 1, 3)
    k = qkv[:, :, 1].transpose(0, 2, 1, 3)
    v = qkv[:, :, 2].transpose(0, 2, 1, 3)
    dots = np.matmul(q, k.transpose(0, 1, 3, 2)) * (d ** -0.5)   # "bhid,bhjd->bhij"
    dots = dots - dots.max(-1, keepdims=True)
    e = np.exp(dots)
    attn = e / e.sum(-1, keepdims=True)
    out = np.matmul(attn, v)                                    # "bhij,bhjd->bhid"
    out = out.transpose(0, 2, 1, 3).reshape(b, n, d)
    return out @ wo + bo


def mlp(x, w1, b1, w2, b2):
    """base_models.py:45-57."""
    return gelu_tanh(x @ w1 + b1) @ w2 + b2


def transformer_layer(x, w, prefix, heads):
    """One [Residual(Norm(Attention)), Residual(Norm(MLP))] pair, base_models.py:102-106."""
    g = lambda s: w[f"{prefix}/{s}"]
    x = x + attention(layer_norm(x, g("attn/norm/gamma"), g("attn/norm/beta")),
                      g("attn/to_qkv/kernel"), g("attn/to_out/kernel"), g("attn/to_out/bias"), heads)
    x = x + mlp(layer_norm(x, g("mlp/norm/gamma"), g("mlp/norm/beta")),
                g("mlp/dense_0/kernel"), g("mlp/dense_0/bias"), g("mlp/dense_1/kernel"), g("mlp/dense_1/bias"))
    return x


def transformer(x, w, prefix, layers, heads):
    for i in range(layers):
        x = transformer_layer(x, w, f"{prefix}/layer_{i}", heads)
    return x  # no final LayerNorm (base_models.py:109-110)


# --------------------------------------------------------------------------- model
def encode_modalities(w, dims: Dims, motion_input, audio_input):
    """fact_model.py:88-96 -> (motion_features [B,120,d], audio_features [B,240,d])."""
    if motion_input.shape[1] != dims.motion_seq or audio_input.shape[1] != dims.audio_seq:
        raise ValueError("sequence length must equal the position table length (base_models.py:154-156)")
    m = motion_input @ w["motion_linear_embedding/kernel"] + w["motion_linear_embedding/bias"]
    m = m + w["motion_pos_embedding"]
    m = transformer(m, w, "motion_transformer", dims.motion_layers, dims.heads)
    a = audio_input @ w["audio_linear_embedding/kernel"] + w["audio_linear_embedding/bias"]
    a = a + w["audio_pos_embedding"]
    a = transformer(a, w, "audio_transformer", dims.audio_layers, dims.heads)
    return m, a


def call(w, dims: Dims, inputs: dict, dtype=np.float64):
    """FACTModel.call (fact_model.py:72-101): {"motion_input","audio_input"} -> [B, 360, out_dim]."""
    w = {k: np.asarray(v, dtype) for k, v in w.items()}
    m, a = encode_modalities(w, dims, np.asarray(inputs["motion_input"], dtype),
                             np.asarray(inputs["audio_input"], dtype))
    x = np.concatenate([m, a], axis=1)  # motion first (base_models.py:192-193)
    x = transformer(x, w, "cross_modal_layer/transformer", dims.cross_layers, dims.heads)
    return x @ w["cross_modal_layer/output/kernel"] + w["cross_modal_layer/output/bias"]


def infer_auto_regressive(w, dims: Dims, inputs: dict, steps: int = 1200, dtype=np.float64):
    """fact_model.py:103-132: slide the audio window by one, keep row 0, shift it into the motion window."""
    motion = np.asarray(inputs["motion_input"], dtype)
    audio_all = np.asarray(inputs["audio_input"], dtype)
    outs = []
    for i in range(steps):
        audio = audio_all[:, i:i + dims.audio_seq]
        if audio.shape[1] < dims.audio_seq:
            break
        out = call(w, dims, {"motion_input": motion, "audio_input": audio}, dtype)[:, 0:1]
        outs.append(out)
        motion = np.concatenate([motion[:, 1:], out], axis=1)
    if not outs:
        raise ValueError("audio shorter than one window")  # tf.concat([]) raises in the reference
    return np.concatenate(outs, axis=1)


def loss(target, pred):
    """fact_model.py:134-148: mean squared error on the first target_seq_len rows of pred."""
    target = np.asarray(target, np.float64)
    pred = np.asarray(pred, np.float64)
    t = target.shape[1]
    return float(np.mean((target - pred[:, :t]) ** 2))


def per_joint_l2(a, b):
    """Parity metric: on dims 6.., L2 of the 3-dim translation and of each 9-dim rotmat;
    returns the max over batch, time and the 25 groups.  Inputs [..., 225]."""
    d = np.asarray(a, np.float64)[..., 6:] - np.asarray(b, np.float64)[..., 6:]
    if d.shape[-1] < 12 or (d.shape[-1] - 3) % 9:
        return float(np.abs(d).max()) if d.size else 0.0  # toy dims: no joint structure
    groups = [np.linalg.norm(d[..., :3], axis=-1)]
    rot = d[..., 3:].reshape(*d.shape[:-1], -1, 9)
    groups.append(np.linalg.norm(rot, axis=-1).max(-1))
    return float(np.maximum(groups[0], groups[1]).max())
