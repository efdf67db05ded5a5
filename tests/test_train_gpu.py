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

SAME_OPERANDS_TOL = 8e-3       # vs the oracle on the same bf16-rounded operands (see _oracle_grads)

SMALL = dict(d=64, heads=4, ff=128, layers=(1, 1, 2), motion_seq=12, audio_seq=20, motion_dim=225, out_dim=225)


def _oracle_grads(dims, w, inp, t_len, bf16_operands=False):
    """fp64 autograd of the oracle; bf16_operands=True rounds every tensor-core operand (forward and backward) the way
    the product's bf16 training path does, so what is left is accumulation / kernel error, not operand rounding."""
    wt = OT.to_torch(w, torch.float64, requires_grad=True)
    pred = OT.call(wt, dims, inp, bf16_operands=bf16_operands)
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
    worst = _compare(m, gref, tol=2e-2)
    print("small config: loss", loss, loss_ref, "worst grad rel err", worst)
    # against the oracle on the SAME bf16-rounded operands the tolerance is accumulation error, not bf16 rounding
    loss_q, gq = _oracle_grads(dims, w, inp, 5, bf16_operands=True)
    assert abs(loss - loss_q) < 2e-3 * loss_q, (loss, loss_q)
    worst_q = _compare(m, gq, tol=SAME_OPERANDS_TOL, cos_tol=0.9999)
    print("small config vs same-operand oracle: loss", loss, loss_q, "worst grad rel err", worst_q)
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
    worst = _compare(m, gref, tol=2e-2, cos_tol=0.9995)
    print("fact_v5: loss", loss, loss_ref, "worst grad rel err", worst)
    loss_q, gq = _oracle_grads(dims, w, inp, 20, bf16_operands=True)
    assert abs(loss - loss_q) < 2e-3 * loss_q, (loss, loss_q)
    worst_q = _compare(m, gq, tol=SAME_OPERANDS_TOL, cos_tol=0.9999)
    print("fact_v5 vs same-operand oracle: loss", loss, loss_q, "worst grad rel err", worst_q)


def test_batch128_training_step_equals_the_batch2_step(cuda, fact_lib):
    """The B = 128 step is where gemm_wgrad2_kernel (256x256 pair tiles, bulk reductions), the CTA-pair dgrad GEMMs
    and the persistent LayerNorm-backward ring run inside fact_train_step.  No 128-clip oracle run is needed: the
    loss is a mean, so 2 clips duplicated 64 times have the SAME loss and gradients as the 2 clips alone (1/64 is a
    power of two, so the scaling itself is exact).  What differs is the kernel path (tile shapes, fp32 summation
    order): last-bit differences before a bf16 store flip a rounding now and then, and those flips compound through
    ~100 rounding points -- measured on B200: median 2e-4 per tensor, worst 2.6e-3 on the deepest tensors (motion
    layer 0 / embedding), the same for 4, 8, 16, 32 and 128 clips (scripts/experiments/diag_b128.py).  A wrong tile,
    a dropped token slice or a mis-scaled reduction would show as >= 1e-2."""
    dims = oracle_dims()
    w = O.init_weights(dims, seed=2)
    two = O.synthetic_inputs(dims, batch=2, seed=11)
    t2 = {k: torch.from_numpy(v).float() for k, v in two.items()}
    m = FACTModel(make_config(), is_training=True, mode="bf16")
    m.set_weights(w)
    loss2 = float(m.forward_backward(t2, t2["target"]))
    g2 = m.flat_gradients.clone()
    idx = torch.arange(128) % 2
    idx[5], idx[6] = 1, 0
    # equal numbers of both clips, so the mean over the batch is the batch-2 mean
    assert int((idx == 0).sum()) == 64
    big = {k: v[idx].contiguous() for k, v in t2.items()}
    loss128 = float(m.forward_backward(big, big["target"]))
    g128 = m.flat_gradients.clone()
    assert abs(loss128 - loss2) <= 2e-6 * abs(loss2), (loss128, loss2)
    worst, errs = (0.0, None), []
    for name, (off, cnt) in m._offsets.items():
        a, b = g128[off:off + cnt].double(), g2[off:off + cnt].double()
        nb = float(b.norm())
        if nb < 1e-12:
            assert float(a.norm()) < 1e-7, name
            continue
        err = float((a - b).norm()) / nb
        worst = max(worst, (err, name))
        errs.append(err)
        assert err < 6e-3, (name, err)
    errs.sort()
    print("batch 128 vs batch 2 gradients: worst rel err", worst, "median", errs[len(errs) // 2])
    assert errs[len(errs) // 2] < 1e-3
    # run to run: the loss is bitwise stable, gradients agree to fp32 reduction order (bulk reductions are unordered)
    loss_again = float(m.forward_backward(big, big["target"]))
    assert loss_again == loss128
    again = m.flat_gradients
    assert float((again - g128).double().norm()) <= 2e-3 * float(g128.double().norm())
    assert torch.isfinite(again).all()


def test_stage_events_and_device_side_clip(cuda, fact_lib):
    """fact_train_step's stage events fire in bucket order and the staged slices cover the bucket; the on-device clip
    (fact_sum_squares -> fact_clip_scale) equals tf.clip_by_global_norm."""
    from mint_b200.trainer import clip_by_global_norm_
    dims = oracle_dims(audio_dim=35, **SMALL)
    m = FACTModel(make_config(**SMALL), is_training=True, mode="bf16")
    inp = O.synthetic_inputs(dims, batch=2, seed=7, target_len=5)
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
    stages = m.gradient_stages()
    total = m.flat_parameters.numel()
    spans = sorted((o, o + c) for o, c, _ in stages)
    assert spans[0][0] == 0 and spans[-1][1] == total and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    n_ev = m.num_gradient_stage_events
    assert [e for _, _, e in stages] == sorted(e for _, _, e in stages) and stages[-1][2] == n_ev - 1
    head_off = m._offsets["cross_modal_layer/output/kernel"][0]
    assert stages[0][:2] == (head_off, m._offsets["motion_transformer/layer_0/attn/norm/gamma"][0] - head_off)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_ev)]
    end = torch.cuda.Event(enable_timing=True)
    for e in ev:
        e.record()
    m.forward_backward(tin, tin["target"])
    ref = m.flat_gradients.clone()
    # every slice must already hold its final value when its event fires: snapshot each slice on a side stream that
    # only waits for that event, while the main stream keeps running the backward
    side = torch.cuda.Stream()
    snaps = []
    m.forward_backward(tin, tin["target"], stage_events=ev)
    end.record()
    with torch.cuda.stream(side):
        for off, cnt, e in stages:
            side.wait_event(ev[e])
            snaps.append(m.flat_gradients[off:off + cnt].clone())
    torch.cuda.synchronize()
    assert all(ev[i].elapsed_time(ev[i + 1]) >= 0 for i in range(n_ev - 1)) and ev[-1].elapsed_time(end) >= 0
    for (off, cnt, _), snap in zip(stages, snaps):
        assert float((snap - ref[off:off + cnt]).abs().max()) <= 1e-6 * float(ref.abs().max())
    g = m.flat_gradients
    norm = float(g.double().norm())
    before = g.clone()
    clip_by_global_norm_(m, norm * 2)                       # above the norm: untouched
    assert torch.equal(g, before)
    clip_by_global_norm_(m, norm / 4)
    torch.cuda.synchronize()
    assert torch.allclose(g, before * 0.25, rtol=2e-6, atol=0)
    assert abs(float(g.double().norm()) - norm / 4) < 1e-5 * norm


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


def test_adam_in_slices_equals_adam_on_the_bucket(cuda, fact_lib):
    """begin_step / apply_range over disjoint slices / end_step (what the data-parallel trainer does as all-reduce
    slices land) is bit-identical to one apply_gradients over the bucket, including the bf16 operand copies."""
    dims = oracle_dims(audio_dim=35, **SMALL)
    inp = O.synthetic_inputs(dims, batch=2, seed=9, target_len=5)
    tin = {k: torch.from_numpy(v).float() for k, v in inp.items()}
    models, opts = [], []
    for _ in range(2):
        m = FACTModel(make_config(**SMALL), is_training=True, mode="bf16", seed=3)
        models.append(m)
        opts.append(Adam(m, learning_rate=3e-3))
    for step in range(3):
        for m in models:
            m.forward_backward(tin, tin["target"])
        # (the two backward passes agree only to fp32 reduction order: bulk reductions are unordered)
        models[1].flat_gradients.copy_(models[0].flat_gradients)       # identical inputs to both optimizers
        opts[0].apply_gradients()
        total = models[1].flat_parameters.numel()
        cuts = [0, (total // 3) // 8 * 8, (2 * total // 3) // 8 * 8, total]
        opts[1].begin_step()
        for a, b in reversed(list(zip(cuts, cuts[1:]))):                # any order
            opts[1].apply_range(a, b - a)
        opts[1].end_step()
        assert opts[0].iterations == opts[1].iterations == step + 1
        assert torch.equal(models[0].flat_parameters, models[1].flat_parameters)
        assert torch.equal(opts[0].m, opts[1].m) and torch.equal(opts[0].v, opts[1].v)
        assert torch.equal(models[0].flat_bf16_parameters, models[1].flat_bf16_parameters)
        models[1].flat_parameters.copy_(models[0].flat_parameters)
    with pytest.raises(ValueError):
        opts[1].apply_range(2, 8)
