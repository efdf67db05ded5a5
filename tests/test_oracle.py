"""CPU checks that pin the oracle as far as the reference allows (no numeric golden exists upstream:
mint/core/fact_model_test.py:23-54 and base_models_test.py:20-40 assert shapes only)."""
import json
import os

import einops
import numpy as np
import pytest
import torch

from oracle import fact_oracle as O
from oracle import fact_oracle_torch as OT
from tests.helpers import oracle_dims

SMALL = oracle_dims(d=32, heads=4, ff=48, layers=(1, 1, 2), motion_seq=6, audio_seq=10, motion_dim=225, out_dim=225)
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_reference_shape_test_default_768():
    """fact_model_test.py:23-54: hidden 768 / 12 heads defaults, ones input -> (2, 360, 225)."""
    dims = oracle_dims(d=768, heads=12)
    w = OT.to_torch(O.init_weights(dims, 0))
    out = OT.call(w, dims, {"motion_input": np.ones((2, 120, 225), np.float32),
                            "audio_input": np.ones((2, 240, 35), np.float32)})
    assert tuple(out.shape) == (2, 360, 225)


def test_transformer_shape_like_base_models_test():
    """base_models_test.py:20-40: Transformer(hidden 20, heads 10) on ones [4,128,20] keeps the shape."""
    dims = oracle_dims(d=20, heads=10, ff=3072, layers=(2, 1, 1))
    w = O.init_weights(dims, 0)
    y = O.transformer(np.ones((4, 128, 20)), w, "motion_transformer", 2, 10)
    assert y.shape == (4, 128, 20)


def test_parameter_count_v5():
    assert sum(int(np.prod(s)) for s in O.weight_shapes(O.FACT_V5).values()) == 120406977


def test_numpy_and_torch_restatements_agree():
    dims = O.FACT_V5
    w = O.init_weights(dims, 0)
    inp = O.synthetic_inputs(dims, 1)
    y64 = O.call(w, dims, inp)
    y32 = OT.call(OT.to_torch(w), dims, inp).numpy()
    assert O.per_joint_l2(y64, y32) < 1e-4
    yt64 = OT.call(OT.to_torch(w, torch.float64), dims, inp).numpy()
    assert np.abs(yt64 - y64).max() < 1e-10


def test_qkv_column_order_matches_einops():
    """base_models.py:71-72 "b n (qkv h d) -> qkv b h n d" with qkv=3, h=heads."""
    b, n, h, dh = 2, 5, 4, 8
    x = np.arange(b * n * 3 * h * dh, dtype=np.float64).reshape(b, n, 3 * h * dh)
    q, k, v = einops.rearrange(x, "b n (qkv h d) -> qkv b h n d", qkv=3, h=h)
    mine = x.reshape(b, n, 3, h, dh)
    assert np.array_equal(q, mine[:, :, 0].transpose(0, 2, 1, 3))
    assert np.array_equal(k, mine[:, :, 1].transpose(0, 2, 1, 3))
    assert np.array_equal(v, mine[:, :, 2].transpose(0, 2, 1, 3))
    o = np.arange(b * h * n * dh, dtype=np.float64).reshape(b, h, n, dh)
    assert np.array_equal(einops.rearrange(o, "b h n d -> b n (h d)"), o.transpose(0, 2, 1, 3).reshape(b, n, h * dh))


def test_attention_scale_is_model_dim():
    """scale = dim ** -0.5 with dim = d_model (base_models.py:66,104): differs from the head_dim scaling."""
    rng = np.random.default_rng(0)
    d, heads = 32, 4
    x = rng.standard_normal((1, 7, d))
    wqkv, wo, bo = rng.standard_normal((d, 3 * d)), rng.standard_normal((d, d)), np.zeros(d)
    y = O.attention(x, wqkv, wo, bo, heads)
    qkv = (x @ wqkv).reshape(1, 7, 3, heads, d // heads).transpose(2, 0, 3, 1, 4)
    s = qkv[0] @ qkv[1].transpose(0, 1, 3, 2) * d ** -0.5
    p = np.exp(s - s.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    assert np.allclose(p.sum(-1), 1.0)
    ref = (p @ qkv[2]).transpose(0, 2, 1, 3).reshape(1, 7, d) @ wo
    assert np.allclose(y, ref)
    s2 = qkv[0] @ qkv[1].transpose(0, 1, 3, 2) * (d // heads) ** -0.5
    p2 = np.exp(s2 - s2.max(-1, keepdims=True)); p2 /= p2.sum(-1, keepdims=True)
    assert not np.allclose(y, (p2 @ qkv[2]).transpose(0, 2, 1, 3).reshape(1, 7, d) @ wo)


def test_layer_is_permutation_equivariant_without_positions():
    w = O.init_weights(SMALL, 1, randomize_affine=True)
    x = np.random.default_rng(2).standard_normal((2, 9, SMALL.d))
    perm = np.random.default_rng(3).permutation(9)
    y = O.transformer_layer(x, w, "motion_transformer/layer_0", SMALL.heads)
    yp = O.transformer_layer(x[:, perm], w, "motion_transformer/layer_0", SMALL.heads)
    assert np.allclose(y[:, perm], yp, atol=1e-12)


def test_gelu_is_tanh_form():
    x = np.linspace(-6, 6, 101)
    assert np.allclose(O.gelu_tanh(x), torch.nn.functional.gelu(torch.from_numpy(x), approximate="tanh").numpy())
    assert not np.allclose(O.gelu_tanh(x), torch.nn.functional.gelu(torch.from_numpy(x)).numpy(), atol=1e-5)


def test_layernorm_eps_and_biased_variance():
    x = np.random.default_rng(0).standard_normal((3, 16)) * 1e-2
    ref = torch.nn.functional.layer_norm(torch.from_numpy(x), (16,), eps=1e-5).numpy()
    assert np.allclose(O.layer_norm(x, np.ones(16), np.zeros(16)), ref, atol=1e-12)


def test_ar_invariants_and_early_stop():
    """fact_model.py:123-131: n = min(steps, T - audio_seq + 1); frame i = row 0 of call() on the shifted window."""
    w = O.init_weights(SMALL, 5, randomize_affine=True)
    inp = O.synthetic_inputs(SMALL, batch=2, audio_len=SMALL.audio_seq + 4, seed=5)
    out = O.infer_auto_regressive(w, SMALL, inp, steps=1200)
    assert out.shape == (2, 5, 225)
    assert O.infer_auto_regressive(w, SMALL, inp, steps=3).shape == (2, 3, 225)
    motion = inp["motion_input"]
    for i in range(3):
        row0 = O.call(w, SMALL, {"motion_input": motion, "audio_input": inp["audio_input"][:, i:i + SMALL.audio_seq]})[:, :1]
        assert np.allclose(row0, out[:, i:i + 1], atol=1e-12)
        motion = np.concatenate([motion[:, 1:], row0], 1)
    ot = OT.infer_auto_regressive(OT.to_torch(w, torch.float64), SMALL, inp, steps=1200).numpy()
    assert np.abs(ot - out).max() < 1e-10
    with pytest.raises(ValueError):
        O.infer_auto_regressive(w, SMALL, O.synthetic_inputs(SMALL, 1, audio_len=SMALL.audio_seq - 1), steps=2)


def test_loss_is_mse_on_first_target_rows():
    rng = np.random.default_rng(0)
    pred, target = rng.standard_normal((3, 16, 225)), rng.standard_normal((3, 5, 225))
    assert np.isclose(O.loss(target, pred), np.mean((target - pred[:, :5]) ** 2))
    assert np.isclose(float(OT.loss(target, torch.from_numpy(pred))), O.loss(target, pred))


def test_committed_golden_vectors():
    """Vectors generated by tests/golden/make_golden.py from the fp64 oracle (small config): guards the oracle
    itself against silent edits (the GPU parity tests lean on it)."""
    path = os.path.join(GOLDEN, "fact_small_fp64.npz")
    g = np.load(path)
    meta = json.loads(str(g["meta"]))
    dims = oracle_dims(**meta["dims"])
    w = O.init_weights(dims, meta["seed"], randomize_affine=True)
    inp = O.synthetic_inputs(dims, meta["batch"], audio_len=meta["audio_len"], seed=meta["seed"],
                             target_len=meta["target_len"])
    assert np.allclose(O.call(w, dims, {"motion_input": inp["motion_input"],
                                        "audio_input": inp["audio_input"][:, :dims.audio_seq]}), g["call"], atol=1e-12)
    assert np.allclose(O.infer_auto_regressive(w, dims, inp, steps=meta["steps"]), g["ar"], atol=1e-12)
    assert np.isclose(O.loss(inp["target"], g["call"]), float(g["loss"]))


def _load_reference_golden():
    g = np.load(os.path.join(GOLDEN, "fact_reference_code_small.npz"))
    meta = json.loads(str(g["meta"]))
    dims = O.Dims(meta["d"], meta["heads"], meta["ff"], *meta["layers"], meta["motion_seq"], meta["audio_seq"],
                  meta["motion_dim"], meta["audio_dim"], meta["out_dim"])
    w = {k[2:]: g[k].astype(np.float64) for k in g.files if k.startswith("w:")}
    return g, meta, dims, w


def test_oracle_matches_the_references_own_model_code():
    """tests/golden/fact_reference_code_small.npz was produced by /root/reference's FACTModel (call,
    infer_auto_regressive, loss) executed over a NumPy shim of the TF primitives (make_reference_golden.py): the
    restatement must reproduce the reference code's composition exactly (float64 both sides)."""
    g, meta, dims, w = _load_reference_golden()
    assert set(w) == set(O.weight_shapes(dims)) and all(w[k].shape == s for k, s in O.weight_shapes(dims).items())
    motion, audio = g["motion"].astype(np.float64), g["audio"].astype(np.float64)
    call = O.call(w, dims, {"motion_input": motion, "audio_input": audio[:, :dims.audio_seq]})
    assert np.abs(call - g["call"]).max() < 1e-11
    ar = O.infer_auto_regressive(w, dims, {"motion_input": motion, "audio_input": audio}, steps=meta["steps"])
    assert ar.shape == g["ar"].shape == (meta["batch"], meta["audio_len"] - meta["audio_seq"] + 1, 225)
    assert np.abs(ar - g["ar"]).max() < 1e-10
    assert abs(O.loss(g["target"], g["call"]) - float(g["loss"])) < 1e-13
    tcall = OT.call(OT.to_torch(w, torch.float64), dims,
                    {"motion_input": motion, "audio_input": audio[:, :dims.audio_seq]}).numpy()
    assert np.abs(tcall - g["call"]).max() < 1e-11


def test_oracle_matches_reference_code_at_fact_v5_size():
    """Same pin at the real dims: the reference code (over the shim) on oracle.init_weights(FACT_V5, seed 0) rounded to
    fp32 and synthetic_inputs(seed 0, batch 1) -- only the outputs are stored (fact_reference_code_v5.npz)."""
    g = np.load(os.path.join(GOLDEN, "fact_reference_code_v5.npz"))
    dims = O.FACT_V5
    w = {k: v.astype(np.float32).astype(np.float64) for k, v in O.init_weights(dims, seed=0).items()}
    inp = O.synthetic_inputs(dims, batch=1, audio_len=dims.audio_seq + 1, seed=0)
    motion = inp["motion_input"].astype(np.float32).astype(np.float64)
    audio = inp["audio_input"].astype(np.float32).astype(np.float64)
    call = O.call(w, dims, {"motion_input": motion, "audio_input": audio[:, :dims.audio_seq]})
    assert np.abs(call - g["call"]).max() < 1e-9
    assert O.per_joint_l2(call, g["call"]) < 1e-9
