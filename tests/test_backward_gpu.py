"""Backward kernels through the C ABI vs torch autograd (fp64) on the same bf16-rounded operands."""
import ctypes as C
import math

import pytest
import torch

from mint_b200 import lib as L
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("tokens,in_dim,out_dim", [(64, 128, 128), (300, 800, 800), (1000, 800, 2400),
                                                    (777, 3072, 800), (513, 800, 256), (2000, 72, 136)])
def test_wgrad_gemm(fact_lib, cuda, tokens, in_dim, out_dim):
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(tokens, in_dim, generator=g).to(cuda).to(BF)
    dy = torch.randn(tokens, out_dim, generator=g).to(cuda).to(BF)
    dw = torch.full((in_dim, out_dim), 0.5, device=cuda)      # accumulates on top
    L.check(fact_lib.fact_wgrad_gemm(x.data_ptr(), in_dim, dy.data_ptr(), out_dim, dw.data_ptr(), out_dim, tokens,
                                     in_dim, out_dim, _st()), "fact_wgrad_gemm")
    ref = x.double().t() @ dy.double() + 0.5
    assert rel_err(dw, ref) < 2e-5, rel_err(dw, ref)


@pytest.mark.parametrize("tokens,in_dim,out_dim", [(4608, 800, 2400), (5000, 3072, 800), (4100, 304, 520),
                                                    (4096, 256, 256), (9001, 800, 3072), (12000, 800, 800)])
def test_wgrad_gemm_pair(fact_lib, cuda, tokens, in_dim, out_dim):
    """Problems large enough for the CTA-pair kernel (256 x NT tiles, bulk fp32 reductions into dW): partial tiles on
    both matrix edges, N tiles that are not a multiple of the 32-column epilogue chunk (800 -> 4 x 208, 520 -> 3 x 176),
    a token count that is not a multiple of the 64-token block, accumulation on top of dW, nothing written outside
    [in, out]; the automatic orientation (flag 1), the plain (2) and the transposed (3) one and the 1-SM kernel (0)
    agree."""
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(tokens, in_dim, generator=g).to(cuda).to(BF)
    dy = torch.randn(tokens, out_dim, generator=g).to(cuda).to(BF)
    ref = x.double().t() @ dy.double() + 0.5
    outs = []
    for flag in (1, 2, 3, 0):
        fact_lib.fact_set_flag(b"wgrad_pair", flag)
        buf = torch.full((in_dim + 8, out_dim), 0.5, device=cuda)     # 8 guard rows behind the gradient
        try:
            L.check(fact_lib.fact_wgrad_gemm(x.data_ptr(), in_dim, dy.data_ptr(), out_dim, buf.data_ptr(), out_dim,
                                             tokens, in_dim, out_dim, _st()), "fact_wgrad_gemm")
            torch.cuda.synchronize()
        finally:
            fact_lib.fact_set_flag(b"wgrad_pair", 1)
        assert (buf[in_dim:] == 0.5).all(), f"flag {flag}: wrote past row {in_dim}"
        assert rel_err(buf[:in_dim], ref) < 2e-5, (flag, rel_err(buf[:in_dim], ref))
        outs.append(buf[:in_dim].clone())
    for o in outs[1:]:
        assert rel_err(outs[0], o.double()) < 1e-5


def test_wgrad_gemm_padded_operand(fact_lib, cuda):
    """dy with a padded pitch (the head's 225 -> 256 columns) and a gradient narrower than the pitch."""
    tokens, in_dim, out_dim, ldy = 500, 800, 225, 256
    x = torch.randn(tokens, in_dim, device=cuda).to(BF)
    dy = torch.zeros(tokens, ldy, device=cuda, dtype=BF)
    dy[:, :out_dim] = torch.randn(tokens, out_dim, device=cuda).to(BF)
    dw = torch.zeros(in_dim, out_dim, device=cuda)
    L.check(fact_lib.fact_wgrad_gemm(x.data_ptr(), in_dim, dy.data_ptr(), ldy, dw.data_ptr(), out_dim, tokens, in_dim,
                                     out_dim, _st()))
    assert rel_err(dw, x.double().t() @ dy[:, :out_dim].double()) < 2e-5


def test_gelu_save_and_grad_epilogues(fact_lib, cuda):
    m, n, k = 300, 3072, 800
    a = torch.randn(m, k, device=cuda).to(BF)
    w = (torch.randn(k, n, device=cuda) / math.sqrt(k))
    bias = 0.1 * torch.randn(n, device=cuda)
    w_fwd = w.t().contiguous().to(BF)                         # [n, k]
    h = torch.empty(m, n, device=cuda, dtype=BF)
    z = torch.empty(m, n, device=cuda, dtype=BF)
    e = L.GemmEpilogue(kind=L.EPI_BIAS_GELU_SAVE, out_hi=h.data_ptr(), out_lo=z.data_ptr(), ldo=n,
                       bias=bias.data_ptr())
    L.check(fact_lib.fact_gemm(a.data_ptr(), None, k, w_fwd.data_ptr(), None, k, m, n, k, C.byref(e), _st()))
    zr = a.double() @ w_fwd.double().t() + bias.double()
    assert rel_err(z.float(), zr) < 6e-3 and rel_err(h.float(), torch.nn.functional.gelu(zr, approximate="tanh")) < 6e-3
    # backward: dz = (dh . W^T-free check) -> use dh as the GEMM "accumulator" through an identity-like product
    dh = torch.randn(m, k, device=cuda).to(BF)                # reuse shapes: dz[m, n] = (dh . w_fwd^T) * gelu'(z)
    dz = torch.empty(m, n, device=cuda, dtype=BF)
    e2 = L.GemmEpilogue(kind=L.EPI_GELU_GRAD, out_hi=dz.data_ptr(), ldo=n, aux=z.data_ptr(), ldaux=n)
    L.check(fact_lib.fact_gemm(dh.data_ptr(), None, k, w_fwd.data_ptr(), None, k, m, n, k, C.byref(e2), _st()))
    zz = z.double().requires_grad_(True)
    torch.nn.functional.gelu(zz, approximate="tanh").sum().backward()
    ref = (dh.double() @ w_fwd.double().t()) * zz.grad
    assert rel_err(dz.float(), ref) < 6e-3


@pytest.mark.parametrize("legacy", [0, 1])
@pytest.mark.parametrize("batch,n,heads,dh", [(2, 120, 10, 80), (1, 360, 10, 80), (2, 37, 2, 16), (1, 130, 3, 64),
                                              (3, 240, 4, 80), (3, 360, 10, 80), (5, 200, 2, 80), (2, 384, 1, 80),
                                              (40, 360, 10, 80), (2, 68, 3, 80), (3, 130, 2, 80)])
def test_sdpa_forward_lse_and_backward(fact_lib, cuda, batch, n, heads, dh, legacy):
    if legacy and dh != 80:
        pytest.skip("already the mma.sync path")
    d = heads * dh
    scale = d ** -0.5
    g = torch.Generator(device="cpu").manual_seed(3)
    qkv_f = torch.randn(batch * n, 3 * d, generator=g).to(cuda)
    qkv_f[:, :d] *= 3.0                                        # make the softmax non-trivial
    q_scaled = qkv_f.clone()
    q_scaled[:, :d] *= scale * math.log2(math.e)               # what the QKV epilogue stores
    qkv = q_scaled.to(BF)
    o = torch.zeros(batch * n, d, device=cuda, dtype=BF)
    lse = torch.zeros(batch, heads, n, device=cuda)
    fact_lib.fact_set_flag(b"sdpa_legacy", legacy)
    try:
        L.check(fact_lib.fact_sdpa_lse(qkv.data_ptr(), None, o.data_ptr(), None, lse.data_ptr(), batch, n, heads, dh,
                                       _st()))
    finally:
        fact_lib.fact_set_flag(b"sdpa_legacy", 0)
    # reference on the bf16-rounded stored operands, in terms of the UNSCALED q (= q' / (scale log2 e))
    eff = qkv.double().view(batch, n, 3, heads, dh)
    qp = eff[:, :, 0].permute(0, 2, 1, 3)                       # q' (log2 domain)
    q = (qp / (scale * math.log2(math.e))).detach().requires_grad_(True)
    k = eff[:, :, 1].permute(0, 2, 1, 3).detach().requires_grad_(True)
    v = eff[:, :, 2].permute(0, 2, 1, 3).detach().requires_grad_(True)
    s = (q @ k.transpose(-1, -2)) * scale
    p = torch.softmax(s, -1)
    out = (p @ v).permute(0, 2, 1, 3).reshape(batch * n, d)
    assert (o.double() - out).abs().max() < 2e-2 * max(1.0, float(out.abs().max()))
    lse_ref = torch.logsumexp(s, -1) * math.log2(math.e)
    assert (lse.double() - lse_ref).abs().max() < 1e-3
    d_o = torch.randn(batch * n, d, generator=g).to(cuda).to(BF)
    out.backward(d_o.double())
    dref = torch.stack([t.grad.permute(0, 2, 1, 3).reshape(batch * n, d) for t in (q, k, v)], 1).reshape(batch * n, 3 * d)
    # pipelined tcgen05 kernel (1: 64-query steps, two score buffers), first-generation tcgen05 kernel (2), mma.sync (0)
    for bwd_tc in ((1, 2, 0) if dh == 80 else (1,)):
        dqkv = torch.zeros(batch * n, 3 * d, device=cuda, dtype=BF)
        dscr = torch.zeros(batch * heads * n, device=cuda)
        dq_scr = torch.zeros(batch * n * d, device=cuda)
        fact_lib.fact_set_flag(b"sdpa_bwd_tc", bwd_tc)
        try:
            L.check(fact_lib.fact_sdpa_backward(qkv.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(),
                                                dscr.data_ptr(), dq_scr.data_ptr(), dqkv.data_ptr(), batch, n, heads,
                                                dh, scale, _st()))
            torch.cuda.synchronize()
        finally:
            fact_lib.fact_set_flag(b"sdpa_bwd_tc", 1)
        for name, sl in (("dq", slice(0, d)), ("dk", slice(d, 2 * d)), ("dv", slice(2 * d, 3 * d))):
            err = rel_err(dqkv[:, sl].float(), dref[:, sl])
            assert err < 3e-2, (name, err, bwd_tc)


@pytest.mark.parametrize("rows,d", [(37, 800), (1000, 800), (64, 64)])
@pytest.mark.parametrize("with_res", [True, False])
def test_layernorm_backward(fact_lib, cuda, rows, d, with_res):
    x = (torch.randn(rows, d, device=cuda) * 2 + 0.3)
    gamma = 1 + 0.1 * torch.randn(d, device=cuda)
    dy = torch.randn(rows, d, device=cuda)
    dres = torch.randn(rows, d, device=cuda)
    dx = torch.empty(rows, d, device=cuda)
    dg = torch.zeros(d, device=cuda)
    db = torch.zeros(d, device=cuda)
    stats = torch.zeros(rows, 4, device=cuda)
    if with_res:
        dx = dres.clone()          # in place on the residual gradient, as the engine uses it
    L.check(fact_lib.fact_layernorm_backward(x.data_ptr(), gamma.data_ptr(), dy.data_ptr(),
                                             dx.data_ptr() if with_res else None, dx.data_ptr(), dg.data_ptr(),
                                             db.data_ptr(), stats.data_ptr(), rows, d, _st()))
    xx = x.double().requires_grad_(True)
    gg = gamma.double().requires_grad_(True)
    bb = torch.zeros(d, device=cuda, dtype=torch.float64, requires_grad=True)
    torch.nn.functional.layer_norm(xx, (d,), gg, bb, 1e-5).backward(dy.double())
    ref = xx.grad + (dres.double() if with_res else 0)
    assert rel_err(dx, ref) < 2e-5 and rel_err(dg, gg.grad) < 2e-5 and rel_err(db, bb.grad) < 2e-5


def test_embed_backward_and_cast_colsum(fact_lib, cuda):
    batch, n_tok, f, d, x_len = 3, 12, 35, 800, 20
    x = torch.randn(batch, x_len, f, device=cuda)
    dy = torch.randn(batch * n_tok, d, device=cuda)
    dw = torch.zeros(f, d, device=cuda)
    dbias = torch.zeros(d, device=cuda)
    dpos = torch.zeros(n_tok, d, device=cuda)
    L.check(fact_lib.fact_embed_backward(x.data_ptr(), x_len * f, dy.data_ptr(), dw.data_ptr(), dbias.data_ptr(),
                                         dpos.data_ptr(), batch, n_tok, f, d, _st()))
    xw = x[:, :n_tok].reshape(batch * n_tok, f).double()
    assert rel_err(dw, xw.t() @ dy.double()) < 1e-5
    assert rel_err(dbias, dy.double().sum(0)) < 1e-5
    assert rel_err(dpos, dy.double().view(batch, n_tok, d).sum(0)) < 1e-5
    # cast + colsum with a padded pitch
    src = torch.randn(100, 225, device=cuda)
    y = torch.full((100, 256), 7.0, device=cuda, dtype=BF)
    cs = torch.zeros(225, device=cuda)
    L.check(fact_lib.fact_cast_colsum(src.data_ptr(), 225, y.data_ptr(), 256, cs.data_ptr(), 100, 225, _st()))
    assert torch.equal(y[:, :225], src.to(BF)) and float(y[:, 225:].float().abs().max()) == 0.0
    assert rel_err(cs, src.double().sum(0)) < 1e-5


def test_adam_matches_keras_formula(fact_lib, cuda):
    n = 10007
    w = torch.randn(n, device=cuda)
    g = torch.randn(n, device=cuda)
    m = torch.zeros(n, device=cuda)
    v = torch.zeros(n, device=cuda)
    w0 = w.double().clone()
    mm = torch.zeros(n, dtype=torch.float64, device=cuda)
    vv = torch.zeros_like(mm)
    lr, b1, b2, eps = 1e-4, 0.9, 0.999, 1e-7
    ref = w0.clone()
    w_b = torch.zeros(n, device=cuda, dtype=BF)
    for step in (1, 2, 3):
        L.check(fact_lib.fact_adam_step(w.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, lr, b1, b2, eps,
                                        step, 0.5, w_b.data_ptr() if step > 1 else None, _st()))
        gd = g.double() * 0.5
        mm = b1 * mm + (1 - b1) * gd
        vv = b2 * vv + (1 - b2) * gd * gd
        lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        ref = ref - lr_t * mm / (vv.sqrt() + eps)
    assert (w.double() - ref).abs().max() < 1e-6
    assert torch.equal(w_b, w.to(BF))                          # the bf16 operand copy written by the same pass
    ss = torch.zeros((), device=cuda)
    L.check(fact_lib.fact_sum_squares(g.data_ptr(), n, ss.data_ptr(), _st()))
    assert abs(float(ss) - float((g.double() ** 2).sum())) < 1e-3 * float((g.double() ** 2).sum())
