"""Model-level parity: FACTModel (CUDA, through the C ABI) vs the fp64 oracle on the same seeded inputs.

Bar (BASELINE.json north_star): max per-joint L2 on the 219-dim motion vector <= 1e-3 in precise mode."""
import numpy as np
import pytest
import torch

from mint_b200 import config_util, model_builder, protos
from mint_b200.fact_model import FACTModel
from oracle import fact_oracle as O
from tests.helpers import make_config, oracle_dims

pytestmark = pytest.mark.gpu

PARITY_TOL = 1e-3          # per-joint L2, precise mode (the north_star tolerance)
BF16_TOL = 0.25            # throughput mode: single bf16 products; reported, not a parity claim

SMALL = dict(d=64, heads=4, ff=128, layers=(1, 1, 2), motion_seq=12, audio_seq=20, motion_dim=225, out_dim=225)


def _model(cfg, w, mode, **kw):
    m = FACTModel(cfg, is_training=False, mode=mode, **kw)
    m.set_weights(w)
    return m


@pytest.fixture(scope="module")
def v5():
    dims = oracle_dims()
    w = O.init_weights(dims, seed=0)
    inp = O.synthetic_inputs(dims, batch=2, seed=0)
    ref = O.call(w, dims, inp)
    return dims, w, inp, ref


def test_forward_shape_like_reference_test(cuda, fact_lib):
    """mint/core/fact_model_test.py:23-54: default hidden 768 / 12 heads, ones input -> (2, 360, 225)."""
    cfg = make_config(d=768, heads=12, ff=3072, layers=(2, 2, 12))
    m = FACTModel(cfg, is_training=True)
    out = m({"motion_input": torch.ones(2, 120, 225), "audio_input": torch.ones(2, 240, 35)})
    assert tuple(out.shape) == (2, 360, 225)
    assert torch.isfinite(out).all()


def test_forward_parity_precise(cuda, fact_lib, v5):
    dims, w, inp, ref = v5
    m = _model(make_config(), w, "precise")
    out = m({k: torch.from_numpy(v).float() for k, v in inp.items()}).cpu().numpy()
    err = O.per_joint_l2(out, ref)
    print("precise per-joint L2:", err)
    assert err <= PARITY_TOL, err


def test_forward_fp32_simt_crosscheck(cuda, fact_lib, v5):
    """CUDA-core GEMMs over the same split operands: separates tcgen05 plumbing errors from numerics."""
    dims, w, inp, ref = v5
    m = _model(make_config(), w, "fp32_simt")
    out = m({k: torch.from_numpy(v).float() for k, v in inp.items()}).cpu().numpy()
    err = O.per_joint_l2(out, ref)
    print("fp32_simt per-joint L2:", err)
    assert err <= PARITY_TOL, err


def test_forward_bf16_mode_reported(cuda, fact_lib, v5):
    dims, w, inp, ref = v5
    m = _model(make_config(), w, "bf16")
    out = m({k: torch.from_numpy(v).float() for k, v in inp.items()}).cpu().numpy()
    err = O.per_joint_l2(out, ref)
    print("bf16 per-joint L2:", err)
    assert err <= BF16_TOL, err


@pytest.mark.parametrize("mode,tol", [("precise", 2e-4), ("fp32_simt", 2e-4)])
def test_small_config_with_affine(cuda, fact_lib, mode, tol):
    """Non-trivial biases / LN affine (all trivial at Keras init), odd sequence lengths, 16-dim heads."""
    dims = oracle_dims(audio_dim=35, **SMALL)
    w = O.init_weights(dims, seed=3, randomize_affine=True)
    inp = O.synthetic_inputs(dims, batch=3, seed=3)
    ref = O.call(w, dims, inp)
    m = _model(make_config(**SMALL), w, mode)
    out = m({k: torch.from_numpy(v).float() for k, v in inp.items()}).cpu().numpy()
    assert np.abs(out - ref).max() <= tol * max(1.0, np.abs(ref).max()), np.abs(out - ref).max()


def test_loss_matches_oracle(cuda, fact_lib, v5):
    dims, w, inp, ref = v5
    m = _model(make_config(), w, "precise")
    pred = torch.from_numpy(ref).float()
    got = float(m.loss(torch.from_numpy(inp["target"]).float(), pred))
    want = O.loss(inp["target"], ref)
    assert abs(got - want) <= 1e-5 * want


@pytest.mark.parametrize("use_graph", [False, True])
def test_ar_matches_oracle_small(cuda, fact_lib, use_graph):
    dims = oracle_dims(audio_dim=35, **SMALL)
    w = O.init_weights(dims, seed=4, randomize_affine=True)
    steps = 9
    inp = O.synthetic_inputs(dims, batch=2, audio_len=dims.audio_seq + 6, seed=4)   # early stop after 7 frames
    ref = O.infer_auto_regressive(w, dims, inp, steps=steps)
    assert ref.shape[1] == 7
    m = _model(make_config(**SMALL), w, "precise", use_graph=use_graph)
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
    out = m.infer_auto_regressive(tin, steps=steps).cpu().numpy()
    assert out.shape == ref.shape
    assert np.abs(out - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())
    # and again (graph cache hit, step counter reset)
    out2 = m.infer_auto_regressive(tin, steps=steps).cpu().numpy()
    assert np.array_equal(out, out2)


def test_ar_parity_v5(cuda, fact_lib):
    """fact_v5, 4 generated frames, batch 2: each frame within the per-joint bar; frame i equals row 0 of a
    single forward on the shifted window (the invariant of fact_model.py:123-131)."""
    dims = oracle_dims()
    w = O.init_weights(dims, seed=0)
    inp = O.synthetic_inputs(dims, batch=2, audio_len=dims.audio_seq + 3, seed=1)
    ref = O.infer_auto_regressive(w, dims, inp, steps=4)
    m = _model(make_config(), w, "precise")
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
    out = m.infer_auto_regressive(tin, steps=1200)
    assert tuple(out.shape) == (2, 4, 225)
    err = O.per_joint_l2(out.cpu().numpy(), ref)
    print("AR per-joint L2:", err)
    assert err <= PARITY_TOL
    motion1 = torch.cat([tin["motion_input"][:, 1:], out[:, :1].cpu()], 1)
    single = m({"motion_input": motion1, "audio_input": tin["audio_input"][:, 1:241]})[:, 0]
    assert (single - out[:, 1]).abs().max() < 1e-4


def test_ar_row0_pruning_is_exact(cuda, fact_lib):
    """The pruned last layer (row 0 only) must reproduce the full last layer on the kept row: same products, only the
    fp32 summation order of the small-M split-K GEMMs differs (bit-identical with split-K off)."""
    dims = oracle_dims()
    w = O.init_weights(dims, seed=2)
    inp = O.synthetic_inputs(dims, batch=3, audio_len=dims.audio_seq + 2, seed=2)
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
    outs = []
    for flag in (1, 0):
        fact_lib.fact_set_flag(b"ar_prune", flag)
        m = _model(make_config(), w, "precise")
        outs.append(m.infer_auto_regressive(tin, steps=3).cpu())
    fact_lib.fact_set_flag(b"ar_prune", 1)
    assert (outs[0] - outs[1]).abs().max() <= 2e-5 * outs[1].abs().max()
    fact_lib.fact_set_flag(b"gemm_splitk", 0)
    try:
        exact = []
        for flag in (1, 0):
            fact_lib.fact_set_flag(b"ar_prune", flag)
            m = _model(make_config(), w, "precise")
            exact.append(m.infer_auto_regressive(tin, steps=2).cpu())
    finally:
        fact_lib.fact_set_flag(b"ar_prune", 1)
        fact_lib.fact_set_flag(b"gemm_splitk", 1)
    assert torch.equal(exact[0], exact[1])


def test_model_builder_and_config(cuda, fact_lib):
    cfg = config_util.get_configs_from_pipeline_file(config_util.DEFAULT_CONFIG)
    m = model_builder.build(cfg["model"], True)
    assert isinstance(m, FACTModel) and m.get_metrics(cfg["eval_config"]) == []
    assert sum(v.numel() for v in m.trainable_variables) == 120406977
    bad = protos.MultiModalModel()
    with pytest.raises(ValueError):
        model_builder.build(bad, True)


def test_input_validation(cuda, fact_lib):
    m = FACTModel(make_config(**SMALL), False)
    with pytest.raises(ValueError):
        m({"motion_input": torch.zeros(1, 11, 225), "audio_input": torch.zeros(1, 20, 35)})
    with pytest.raises(ValueError):
        m.infer_auto_regressive({"motion_input": torch.zeros(1, 12, 225), "audio_input": torch.zeros(1, 19, 35)})


def test_full_size_batch_invariance_and_determinism(cuda, fact_lib):
    """Size-independent properties at the BASELINE batch (128 clips, fact_v5): clips are independent, so (i) a clip's
    output does not depend on its position in the batch or on its neighbours (bitwise: every tile sees the same K
    order), (ii) two runs are bitwise identical, (iii) the batch-128 result equals the batch-2 result of the same
    clips up to fp32 summation order (different kernels: CTA-pair vs 1-SM / split-K)."""
    dims = oracle_dims()
    w = O.init_weights(dims, seed=0)
    m = _model(make_config(), w, "precise")
    two = O.synthetic_inputs(dims, batch=2, seed=9)
    motion = torch.from_numpy(two["motion_input"]).float()
    audio = torch.from_numpy(two["audio_input"]).float()
    idx = torch.arange(128) % 2
    idx[5], idx[6] = 1, 0                                    # break the regular pattern
    big = {"motion_input": motion[idx], "audio_input": audio[idx]}
    out = m(big)
    assert torch.isfinite(out).all()
    ref0, ref1 = out[0], out[1]
    for b in range(128):
        assert torch.equal(out[b], ref0 if idx[b] == 0 else ref1), b
    assert torch.equal(m(big), out)
    small = m({"motion_input": motion, "audio_input": audio})
    assert (small - out[:2]).abs().max() <= 3e-5 * out[:2].abs().max()


@pytest.mark.parametrize("mode", ["precise", "bf16"])
def test_layernorm_inside_the_gemm_launch_is_bitwise_neutral(cuda, fact_lib, mode):
    """Batch 32 (CTA-pair GEMMs): the forward with the LayerNorm done by the residual GEMMs' own warps
    (gemm_fuse_ln = 1, the default) equals, bit for bit, the forward with a LayerNorm launch after each of them, and it
    launches 29 kernels fewer (every out-projection, and every FF2 that is followed by another layer of its stack)."""
    m = FACTModel(make_config(), is_training=False, mode=mode, seed=5)
    g = torch.Generator().manual_seed(6)
    x = {"motion_input": 0.5 * torch.randn(32, 120, 225, generator=g), "audio_input": torch.randn(32, 240, 35, generator=g)}
    outs, counts = [], []
    for flag in (1, 0):
        fact_lib.fact_set_flag(b"gemm_fuse_ln", flag)
        try:
            n0 = fact_lib.fact_launch_count()
            outs.append(m(x).clone())
            counts.append(fact_lib.fact_launch_count() - n0)
        finally:
            fact_lib.fact_set_flag(b"gemm_fuse_ln", 1)
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])
    assert counts[1] - counts[0] == 29, counts


def test_full_size_ar_frames_feed_back(cuda, fact_lib):
    """Batch 128, 3 frames: frame i is row 0 of a plain forward on the window shifted by i (fact_model.py:123-131)."""
    dims = oracle_dims()
    m = FACTModel(make_config(), is_training=False, mode="precise", seed=3)
    g = torch.Generator().manual_seed(4)
    motion = 0.5 * torch.randn(128, 120, 225, generator=g)
    audio = torch.randn(128, 242, 35, generator=g)
    frames = m.infer_auto_regressive({"motion_input": motion, "audio_input": audio}, steps=3)
    assert tuple(frames.shape) == (128, 3, 225)
    win = motion.cuda()
    for i in range(3):
        row0 = m({"motion_input": win, "audio_input": audio[:, i:i + 240]})[:, :1]
        assert (row0[:, 0] - frames[:, i]).abs().max() <= 3e-5 * frames.abs().max()
        win = torch.cat([win[:, 1:], frames[:, i:i + 1]], 1)


def test_cuda_path_matches_the_references_own_model_code(cuda, fact_lib):
    """CUDA (C ABI) vs golden vectors produced by the reference's FACTModel code itself (NumPy shim of TF, float64):
    call, infer_auto_regressive with the early stop, loss."""
    import json
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fact_reference_code_small.npz"))
    meta = json.loads(str(g["meta"]))
    cfg = make_config(d=meta["d"], heads=meta["heads"], ff=meta["ff"], layers=tuple(meta["layers"]),
                      motion_seq=meta["motion_seq"], audio_seq=meta["audio_seq"])
    w = {k[2:]: g[k] for k in g.files if k.startswith("w:")}
    for mode, tol in (("precise", 2e-4), ("fp32_simt", 2e-4)):
        m = _model(cfg, w, mode)
        motion, audio = torch.from_numpy(g["motion"]), torch.from_numpy(g["audio"])
        out = m({"motion_input": motion, "audio_input": audio[:, :meta["audio_seq"]]}).cpu().numpy()
        scale = max(1.0, np.abs(g["call"]).max())
        assert np.abs(out - g["call"]).max() <= tol * scale, (mode, np.abs(out - g["call"]).max())
        ar = m.infer_auto_regressive({"motion_input": motion, "audio_input": audio}, steps=meta["steps"]).cpu().numpy()
        assert ar.shape == g["ar"].shape and np.abs(ar - g["ar"]).max() <= tol * scale
        loss = float(m.loss(torch.from_numpy(g["target"]), torch.from_numpy(g["call"]).float()))
        assert abs(loss - float(g["loss"])) <= 1e-5 * float(g["loss"])


def test_cuda_path_matches_reference_code_at_fact_v5_size(cuda, fact_lib):
    """The headline parity claim against the reference's own code: fact_v5 dims, Keras-default seeded weights,
    call + 2 AR frames, per-joint L2 <= 1e-3 (precise mode)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fact_reference_code_v5.npz"))
    dims = oracle_dims()
    w = {k: v.astype(np.float32) for k, v in O.init_weights(dims, seed=0).items()}
    inp = O.synthetic_inputs(dims, batch=1, audio_len=dims.audio_seq + 1, seed=0)
    motion = torch.from_numpy(inp["motion_input"]).float()
    audio = torch.from_numpy(inp["audio_input"]).float()
    m = _model(make_config(), w, "precise")
    out = m({"motion_input": motion, "audio_input": audio[:, :dims.audio_seq]}).cpu().numpy()
    err = O.per_joint_l2(out, g["call"])
    ar = m.infer_auto_regressive({"motion_input": motion, "audio_input": audio}, steps=1200).cpu().numpy()
    err_ar = O.per_joint_l2(ar, g["ar"])
    print("vs reference code, fact_v5: call", err, "AR", err_ar)
    assert ar.shape == (1, 2, 225) and err <= PARITY_TOL and err_ar <= PARITY_TOL


def test_ar_graphs_belong_to_the_model_session(cuda, fact_lib):
    """Captured frame graphs are owned by the model's session: repeated calls of one shape reuse ONE graph (the staging
    buffers make the key stable by construction), a second model has its own session, the default session stays empty,
    and a model that is freed takes its graphs with it."""
    dims = oracle_dims(audio_dim=35, **SMALL)
    w = O.init_weights(dims, seed=3)
    inp = O.synthetic_inputs(dims, batch=2, audio_len=dims.audio_seq + 5, seed=3)
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items() if k != "target"}
    a = _model(make_config(**SMALL), w, "precise")
    base_default = fact_lib.fact_ar_session_graphs(None)
    first = a.infer_auto_regressive(tin, steps=6)
    assert fact_lib.fact_ar_session_graphs(a._session) == 1
    for _ in range(3):
        again = a.infer_auto_regressive(tin, steps=6)
        assert torch.equal(again, first)
    assert fact_lib.fact_ar_session_graphs(a._session) == 1            # same staging buffers -> same key
    a.infer_auto_regressive(tin, steps=3)                               # another shape: a second graph
    assert fact_lib.fact_ar_session_graphs(a._session) == 2
    b = _model(make_config(**SMALL), O.init_weights(dims, seed=4), "precise")
    other = b.infer_auto_regressive(tin, steps=6)
    assert fact_lib.fact_ar_session_graphs(b._session) == 1 and not torch.equal(other, first)
    assert fact_lib.fact_ar_session_graphs(None) == base_default        # nothing leaked into the default session
    # new weights in the same model: set_weights repacks in place (same pointers, new values) -> the graph still valid
    a.set_weights(O.init_weights(dims, seed=4))
    assert torch.equal(a.infer_auto_regressive(tin, steps=6), other)
    # a session keeps at most 12 graphs, evicting one at a time
    for steps in range(1, 6):
        for extra in range(3):
            t2 = dict(tin)
            t2["audio_input"] = torch.cat([tin["audio_input"], tin["audio_input"][:, :extra + 1]], dim=1)
            a.infer_auto_regressive(t2, steps=steps)
    assert fact_lib.fact_ar_session_graphs(a._session) <= 12
    assert torch.equal(a.infer_auto_regressive(tin, steps=6), other)
    del a, b
