"""Training step (forward + backward + Adam) through the C ABI vs torch autograd on the fp64 oracle."""
import numpy as np
import pytest
import torch

from mint_b200.fact_model import FACTModel
from mint_b200.optim import Adam
from mint_b200.trainer import SingleTaskTrainer, clip_scale
from oracle import fact_oracle as O
from oracle import fact_oracle_torch as OT
from tests.helpers import make_config, oracle_dims

pytestmark = pytest.mark.gpu

SMALL = dict(d=64, heads=4, ff=128, layers=(1, 1, 2), motion_seq=12, audio_seq=20, motion_dim=225, out_dim=225)


def _oracle_grads(dims, w, inp, t_len):
    wt = OT.to_torch(w, torch.float64, requires_grad=True)
    pred = OT.call(wt, dims, inp)
    loss = OT.loss(torch.from_numpy(inp["target"][:, :t_len]), pred)
    loss.backward()
    return float(loss), {k: v.grad.numpy() for k, v in wt.items()}


def _compare(model, ref_grads, tol, cos_tol=0.999):
    worst = (0.0, None)
    for name, g in model.gradients().items():
        got = g.detach().double().cpu().numpy().ravel()
        ref = ref_grads[name].ravel()
        nr = np.linalg.norm(ref)
        if nr < 1e-12:
            assert np.linalg.norm(got) < 1e-6, name
            continue
        err = np.linalg.norm(got - ref) / nr
        cos = float(got @ ref / (np.linalg.norm(got) * nr))
        if err > worst[0]:
            worst = (err, name)
        assert err < tol and cos > cos_tol, (name, err, cos)
    return worst


def test_gradients_small_config(cuda, fact_lib):
    dims = oracle_dims(audio_dim=35, **SMALL)
    w = O.init_weights(dims, seed=5, randomize_affine=True)
    inp = O.synthetic_inputs(dims, batch=3, seed=5, target_len=5)
    loss_ref, gref = _oracle_grads(dims, w, inp, 5)
    m = FACTModel(make_config(**SMALL), is_training=True, mode="bf16")
    m.set_weights(w)
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
    loss = float(m.forward_backward(tin, tin["target"], loss_scale=1.0))
    assert abs(loss - loss_ref) < 2e-2 * loss_ref, (loss, loss_ref)
    worst = _compare(m, gref, tol=6e-2)
    print("small config: loss", loss, loss_ref, "worst grad rel err", worst)
    # loss_scale scales the gradients, not the reported loss
    g1 = m.flat_gradients.clone()
    loss2 = float(m.forward_backward(tin, tin["target"], loss_scale=0.25))
    assert abs(loss2 - loss) < 1e-6 * abs(loss)
    assert torch.allclose(m.flat_gradients, 0.25 * g1, rtol=1e-3, atol=1e-7)


def test_gradients_fact_v5(cuda, fact_lib):
    dims = oracle_dims()
    w = O.init_weights(dims, seed=1)
    inp = O.synthetic_inputs(dims, batch=2, seed=1)
    loss_ref, gref = _oracle_grads(dims, w, inp, 20)
    m = FACTModel(make_config(), is_training=True, mode="bf16")
    m.set_weights(w)
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
    loss = float(m.forward_backward(tin, tin["target"]))
    assert abs(loss - loss_ref) < 2e-2 * loss_ref, (loss, loss_ref)
    worst = _compare(m, gref, tol=8e-2, cos_tol=0.997)
    print("fact_v5: loss", loss, loss_ref, "worst grad rel err", worst)


def test_training_reduces_loss_and_matches_keras_adam_step(cuda, fact_lib):
    dims = oracle_dims(audio_dim=35, **SMALL)
    w = O.init_weights(dims, seed=6, randomize_affine=True)
    inp = O.synthetic_inputs(dims, batch=4, seed=6, target_len=5)
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
    m = FACTModel(make_config(**SMALL), is_training=True, mode="bf16")
    m.set_weights(w)
    opt = Adam(m, learning_rate=1e-3)
    trainer = SingleTaskTrainer([tin] * 100, "target", m, optimizer=opt)
    w0 = m.flat_parameters.clone()
    first = float(trainer.train_step(tin))
    # first Adam step: m/(sqrt(v)+eps) = g/(|g|+eps') -> every touched weight moves by ~lr
    g = m.flat_gradients
    moved = (m.flat_parameters - w0)
    gd = g.double()
    lr_t = 1e-3 * (1 - 0.999) ** 0.5 / (1 - 0.9)
    expect = -lr_t * (0.1 * gd) / ((0.001 * gd * gd).sqrt() + 1e-7)       # Keras: epsilon outside the root
    assert torch.allclose(moved.double(), expect, rtol=1e-3, atol=1e-7)
    assert opt.iterations == 1 and m.global_step == 1
    logs = trainer.train(15)
    assert logs["training_loss"] < first and set(logs) == {"training_loss", "task_loss", "regularization_loss",
                                                           "learning_rate", "steps_per_second"}
    last = float(m.forward_backward(tin, tin["target"]))
    assert last < 0.8 * first, (first, last)


def test_clip_scale(cuda, fact_lib):
    dims = oracle_dims(audio_dim=35, **SMALL)
    m = FACTModel(make_config(**SMALL), is_training=True, mode="bf16")
    inp = O.synthetic_inputs(dims, batch=2, seed=7, target_len=5)
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
    m.forward_backward(tin, tin["target"])
    norm = float(m.flat_gradients.double().norm())
    assert abs(clip_scale(m, norm / 2) - 0.5) < 1e-3 and clip_scale(m, norm * 2) == 1.0


def test_evaluator_writes_reference_layout(cuda, fact_lib, tmp_path):
    """single_task_evaluator.py:64-89: [seed | generated] per clip saved as {motion_name}_{audio_name}.npy."""
    from mint_b200.evaluator import SingleTaskEvaluator
    dims = oracle_dims(audio_dim=35, **SMALL)
    m = FACTModel(make_config(**SMALL), is_training=True)
    inp = O.synthetic_inputs(dims, batch=2, audio_len=dims.audio_seq + 4, seed=8)
    batch = {"motion_input": torch.from_numpy(inp["motion_input"]).float(),
             "audio_input": torch.from_numpy(inp["audio_input"]).float(),
             "motion_name": [b"gBR_sBM_c01", b"gPO_sFM_c02"], "audio_name": ["mBR0", "mPO1"]}
    ev = SingleTaskEvaluator([batch], m, output_dir=str(tmp_path), steps=1200)
    assert ev.evaluate(-1) == {"clips_batches": 1}
    a = np.load(tmp_path / "gBR_sBM_c01_mBR0.npy")
    assert a.shape == (dims.motion_seq + 5, 225) and np.allclose(a[:dims.motion_seq], inp["motion_input"][0], atol=1e-6)
    assert (tmp_path / "gPO_sFM_c02_mPO1.npy").exists()
