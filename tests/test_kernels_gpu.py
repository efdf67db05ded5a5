"""Per-kernel parity tests on the GPU, each through the C ABI, against plain torch fp32/fp64 math."""
import ctypes as C
import math

import pytest
import torch

from mint_b200 import lib as L
from tests.helpers import join, rel_err, split_ref

pytestmark = pytest.mark.gpu


def _st():
    return torch.cuda.current_stream().cuda_stream


def _epi(**kw):
    e = L.GemmEpilogue()
    for k, v in kw.items():
        setattr(e, k, v.data_ptr() if isinstance(v, torch.Tensor) else v)
    return e


@pytest.mark.parametrize("rows,d", [(7, 800), (257, 800), (33, 64), (5, 1024)])
@pytest.mark.parametrize("norm", [True, False])
def test_layernorm_split(fact_lib, cuda, rows, d, norm):
    g = torch.Generator(device="cpu").manual_seed(1)
    x = (torch.randn(rows, d, generator=g) * 3 + 0.5).to(cuda)
    gamma = (1 + 0.1 * torch.randn(d, generator=g)).to(cuda)
    beta = (0.1 * torch.randn(d, generator=g)).to(cuda)
    hi = torch.empty(rows, d, dtype=torch.bfloat16, device=cuda)
    lo = torch.empty_like(hi)
    L.check(fact_lib.fact_layernorm_split(x.data_ptr(), gamma.data_ptr() if norm else None,
                                          beta.data_ptr() if norm else None, hi.data_ptr(), lo.data_ptr(), rows, d,
                                          _st()))
    ref = torch.nn.functional.layer_norm(x.double(), (d,), gamma.double(), beta.double(), 1e-5) if norm else x.double()
    got = join(hi, lo).double()
    assert (got - ref).abs().max() < 2e-5 * max(1.0, float(ref.abs().max()))
    # hi alone is the bf16 rounding of the value
    assert (hi.float().double() - ref).abs().max() < 2 ** -8 * float(ref.abs().max()) + 1e-6


@pytest.mark.parametrize("k,n", [(800, 2400), (3072, 800), (35, 800), (800, 225)])
def test_pack_weight(fact_lib, cuda, k, n):
    w = torch.randn(k, n, device=cuda) * 0.05
    hi = torch.empty(n, k, dtype=torch.bfloat16, device=cuda)
    lo = torch.empty_like(hi)
    L.check(fact_lib.fact_pack_weight(w.data_ptr(), hi.data_ptr(), lo.data_ptr(), k, n, _st()))
    rh, rl = split_ref(w.t().contiguous())
    assert torch.equal(hi, rh) and torch.equal(lo, rl)


def _gemm_case(fact_lib, cuda, m, n, k, kind, precise, seed=0, remap=None, splitk=False):
    g = torch.Generator(device="cpu").manual_seed(seed)
    a = torch.randn(m, k, generator=g).to(cuda)
    w = (torch.randn(k, n, generator=g) * (1.0 / math.sqrt(k))).to(cuda)     # Keras layout
    bias = (0.1 * torch.randn(n, generator=g)).to(cuda)
    resid = torch.randn(m, n, generator=g).to(cuda)
    a_hi, a_lo = split_ref(a)
    w_hi, w_lo = split_ref(w.t().contiguous())
    if precise:
        a_eff, w_eff = join(a_hi, a_lo).double(), join(w_hi, w_lo).double()
    else:
        a_eff, w_eff = a_hi.double(), w_hi.double()
    acc = a_eff @ w_eff.t()
    scale, scale_cols = 0.37, n // 3
    e = L.GemmEpilogue()
    e.kind = kind
    out_rows = m
    if kind in (L.EPI_SPLIT, L.EPI_BIAS_GELU_SPLIT):
        o_hi = torch.zeros(m, n, dtype=torch.bfloat16, device=cuda)
        o_lo = torch.zeros_like(o_hi)
        e.out_hi, e.out_lo, e.ldo = o_hi.data_ptr(), o_lo.data_ptr(), n
        e.bias = bias.data_ptr()
        e.scale, e.scale_cols = scale, scale_cols
    else:
        if remap:
            seq_in, seq_out, seq_off = remap
            out_rows = (m // seq_in) * seq_out
            e.seq_in, e.seq_out, e.seq_off = seq_in, seq_out, seq_off
        out = torch.full((out_rows, n), float("nan"), device=cuda)
        e.out_f32, e.ldo = out.data_ptr(), n
        e.bias = bias.data_ptr()
        e.resid, e.ldr = resid.data_ptr(), n
    if splitk:
        scratch = torch.full((8 * m * n,), float("nan"), device=cuda)
        e.splitk_scratch, e.splitk_scratch_bytes = scratch.data_ptr(), scratch.numel() * 4
    L.check(fact_lib.fact_gemm(a_hi.data_ptr(), a_lo.data_ptr() if precise else None, k, w_hi.data_ptr(),
                               w_lo.data_ptr() if precise else None, k, m, n, k, C.byref(e), _st()), "fact_gemm")
    torch.cuda.synchronize()
    if kind == L.EPI_SPLIT:
        ref = acc.clone()
        ref[:, :scale_cols] *= scale
        got = join(o_hi, o_lo).double()
    elif kind == L.EPI_BIAS_GELU_SPLIT:
        ref = torch.nn.functional.gelu(acc + bias.double(), approximate="tanh")
        got = join(o_hi, o_lo).double()
    elif kind == L.EPI_BIAS_RESID_F32:
        ref = acc + bias.double() + resid.double()
        got = out.double()
        if remap:
            idx = torch.arange(m, device=cuda)
            rows = (idx // seq_in) * seq_out + seq_off + idx % seq_in
            untouched = torch.ones(out_rows, dtype=torch.bool, device=cuda)
            untouched[rows] = False
            assert torch.isnan(out[untouched]).all(), "remapped epilogue wrote outside its rows"
            got = got[rows]
    else:
        ref = acc + bias.double()
        got = out.double()
    # vs the exact product of the operands the kernel was given: only fp32 accumulation + output rounding remain
    tol = 3e-5 if precise else 2e-5
    assert rel_err(got, ref) < tol, (rel_err(got, ref), m, n, k, kind, precise)
    if precise:  # and the split operands reproduce the fp32 GEMM
        full = a.double() @ w.double()
        if kind == L.EPI_BIAS_F32:
            assert rel_err(got, full + bias.double()) < 1e-4


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("kind", [L.EPI_SPLIT, L.EPI_BIAS_GELU_SPLIT, L.EPI_BIAS_RESID_F32, L.EPI_BIAS_F32])
@pytest.mark.parametrize("m,n,k", [(128, 160, 64), (360, 2400, 800), (300, 800, 3072), (250, 3072, 800),
                                   (240, 226, 800), (77, 128, 72)])
def test_gemm_tc(fact_lib, cuda, m, n, k, kind, precise):
    _gemm_case(fact_lib, cuda, m, n, k, kind, precise)


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("kind", [L.EPI_SPLIT, L.EPI_BIAS_GELU_SPLIT, L.EPI_BIAS_RESID_F32, L.EPI_BIAS_F32])
@pytest.mark.parametrize("m,n,k", [(256, 160, 64), (360, 2400, 800), (300, 800, 3072), (700, 3072, 800),
                                   (1000, 320, 136)])
def test_gemm_tc_pair_forced(fact_lib, cuda, m, n, k, kind, precise):
    """cta_group::2 kernel forced on small problems: partial pair tiles (rows of the second CTA out of range)."""
    fact_lib.fact_set_flag(b"gemm_pair", 2)
    try:
        _gemm_case(fact_lib, cuda, m, n, k, kind, precise, seed=7)
    finally:
        fact_lib.fact_set_flag(b"gemm_pair", 1)


def test_gemm_pair_vs_single_bitwise(fact_lib, cuda):
    """Same operands through the 1-SM and the CTA-pair kernels: identical accumulation order -> identical bits."""
    m, n, k = 5000, 2400, 800
    a = torch.randn(m, k, device=cuda)
    w = torch.randn(k, n, device=cuda) / math.sqrt(k)
    a_hi, a_lo = split_ref(a)
    w_hi, w_lo = split_ref(w.t().contiguous())
    outs = []
    for flag in (0, 2):
        fact_lib.fact_set_flag(b"gemm_pair", flag)
        o_hi = torch.zeros(m, n, dtype=torch.bfloat16, device=cuda)
        o_lo = torch.zeros_like(o_hi)
        e = _epi(kind=L.EPI_SPLIT, out_hi=o_hi, out_lo=o_lo, ldo=n, scale=1.0, scale_cols=0)
        L.check(fact_lib.fact_gemm(a_hi.data_ptr(), a_lo.data_ptr(), k, w_hi.data_ptr(), w_lo.data_ptr(), k, m, n, k,
                                   C.byref(e), _st()))
        torch.cuda.synchronize()
        outs.append((o_hi.clone(), o_lo.clone()))
    fact_lib.fact_set_flag(b"gemm_pair", 1)
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("kind", [L.EPI_SPLIT, L.EPI_BIAS_GELU_SPLIT, L.EPI_BIAS_RESID_F32, L.EPI_BIAS_F32,
                                  L.EPI_BIAS_GELU_SAVE, L.EPI_GELU_GRAD, "resid_inplace"])
@pytest.mark.parametrize("m,n,k", [(700, 3072, 800), (360, 2400, 800), (1000, 320, 136), (300, 800, 3072)])
def test_gemm_pair_tma_store_bitwise(fact_lib, cuda, m, n, k, kind, precise):
    """The pair kernel's two epilogues -- direct row-per-lane stores (flag 0) and staged bulk tensor stores (1; 2 =
    without the in-place bulk reduction) -- do the same arithmetic: identical bits, nothing written outside [m, n]."""
    if kind in (L.EPI_BIAS_GELU_SAVE, L.EPI_GELU_GRAD) and precise:
        pytest.skip("training epilogues take bf16 operands only")
    g = torch.Generator(device="cpu").manual_seed(3)
    a = torch.randn(m, k, generator=g).to(cuda)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(cuda)
    bias = (0.1 * torch.randn(n, generator=g)).to(cuda)
    resid = torch.randn(m, n, generator=g).to(cuda)
    z = torch.randn(m, n, generator=g).to(cuda).to(torch.bfloat16)
    a_hi, a_lo = split_ref(a)
    w_hi, w_lo = split_ref(w)
    pad = 64                                   # guard rows after the matrix: must stay NaN / sentinel
    results = []
    fact_lib.fact_set_flag(b"gemm_pair", 2)
    try:
        for flag in (0, 1, 2):
            fact_lib.fact_set_flag(b"gemm_tma_store", flag)
            e = L.GemmEpilogue()
            e.bias = bias.data_ptr()
            e.ldo = n
            if kind == "resid_inplace":
                out = torch.cat([resid.clone(), torch.full((pad, n), 7.0, device=cuda)])
                e.kind, e.out_f32, e.resid, e.ldr = L.EPI_BIAS_RESID_F32, out.data_ptr(), out.data_ptr(), n
                outs = (out,)
            elif kind in (L.EPI_BIAS_RESID_F32, L.EPI_BIAS_F32):
                out = torch.full((m + pad, n), 7.0, device=cuda)
                e.kind, e.out_f32, e.resid, e.ldr = kind, out.data_ptr(), resid.data_ptr(), n
                outs = (out,)
            else:
                o_hi = torch.full((m + pad, n), 7.0, dtype=torch.bfloat16, device=cuda)
                o_lo = torch.full((m + pad, n), 7.0, dtype=torch.bfloat16, device=cuda)
                e.kind, e.out_hi = kind, o_hi.data_ptr()
                if kind != L.EPI_GELU_GRAD:
                    e.out_lo = o_lo.data_ptr()
                e.scale, e.scale_cols = 0.37, n // 3
                e.aux, e.ldaux = z.data_ptr(), n
                outs = (o_hi, o_lo)
                if kind == L.EPI_GELU_GRAD:    # bias gradient folded into the epilogue (or a second pass: flag 0)
                    colsum = torch.zeros(n, device=cuda)
                    e.colsum = colsum.data_ptr()
            L.check(fact_lib.fact_gemm(a_hi.data_ptr(), a_lo.data_ptr() if precise else None, k, w_hi.data_ptr(),
                                       w_lo.data_ptr() if precise else None, k, m, n, k, C.byref(e), _st()))
            torch.cuda.synchronize()
            for o in outs:
                assert (o[m:].float() == 7.0).all(), f"flag {flag}: wrote past row {m}"
            if kind == L.EPI_GELU_GRAD:
                ref_cs = o_hi[:m].double().sum(0)
                assert (colsum.double() - ref_cs).abs().max() < 1e-4 * (1.0 + ref_cs.abs().max()), flag
            results.append([o.clone() for o in outs])
    finally:
        fact_lib.fact_set_flag(b"gemm_pair", 1)
        fact_lib.fact_set_flag(b"gemm_tma_store", 1)
    for r in results[1:]:
        for o_ref, o in zip(results[0], r):
            assert torch.equal(o_ref, o)
    if kind == "resid_inplace":                # and the value is right, not merely consistent
        a_eff = join(a_hi, a_lo).double() if precise else a_hi.double()
        w_eff = join(w_hi, w_lo).double() if precise else w_hi.double()
        ref = a_eff @ w_eff.t() + bias.double() + resid.double()
        assert rel_err(results[1][0][:m].double(), ref) < 3e-5


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("kind", [L.EPI_SPLIT, L.EPI_BIAS_GELU_SPLIT, L.EPI_BIAS_RESID_F32, L.EPI_BIAS_F32])
@pytest.mark.parametrize("m,n,k", [(360, 2400, 800), (360, 800, 3072), (360, 3072, 800), (120, 800, 800),
                                   (128, 226, 800), (77, 128, 200)])
def test_gemm_tc_splitk(fact_lib, cuda, m, n, k, kind, precise):
    """Small-M path: K blocks dealt over the SMs, per-split slabs reduced by the finish kernel (scratch pre-filled with
    NaN: every slab element that is read must have been written)."""
    _gemm_case(fact_lib, cuda, m, n, k, kind, precise, seed=11, splitk=True)


def test_gemm_tc_splitk_remap_and_pitch(fact_lib, cuda):
    _gemm_case(fact_lib, cuda, 2 * 120, 800, 3072, L.EPI_BIAS_RESID_F32, True, remap=(120, 360, 0), splitk=True)


def test_gemm_tc_remap(fact_lib, cuda):
    _gemm_case(fact_lib, cuda, 3 * 120, 800, 3072, L.EPI_BIAS_RESID_F32, True, remap=(120, 360, 0))
    _gemm_case(fact_lib, cuda, 3 * 240, 800, 3072, L.EPI_BIAS_RESID_F32, False, remap=(240, 360, 120))


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("m,seq,pair", [(2400, 120, 2), (2400, 120, 0), (480, 240, 0), (15360, 120, 1)])
def test_gemm_embedding_shape_broadcast_residual(fact_lib, cuda, m, seq, pair, precise):
    """The tensor-core LinearEmbedding: K = 225 inside a 232-element pitch (pad never read), position table as a
    [seq, d] residual broadcast over the clips (resid_rows), through the pair, 1-SM and split-K kernels."""
    n, k, ld = 800, 225, 232
    g = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(m, k, generator=g).to(cuda)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(cuda)
    bias = (0.1 * torch.randn(n, generator=g)).to(cuda)
    pos = torch.randn(seq, n, generator=g).to(cuda)

    def padded(t):
        hi, lo = split_ref(t)
        out_hi = torch.full((t.shape[0], ld), float("nan"), dtype=torch.bfloat16, device=cuda)
        out_lo = torch.full_like(out_hi, float("nan"))
        out_hi[:, :k], out_lo[:, :k] = hi, lo
        return out_hi, out_lo, (join(hi, lo) if precise else hi.float()).double()

    a_hi, a_lo, a_eff = padded(a)
    w_hi, w_lo, w_eff = padded(w)
    out = torch.full((m, n), float("nan"), device=cuda)
    scratch = torch.full((8 * m * n,), float("nan"), device=cuda) if m <= 1024 else None
    e = _epi(kind=L.EPI_BIAS_RESID_F32, out_f32=out, ldo=n, bias=bias, resid=pos, ldr=n)
    e.resid_rows = seq
    if scratch is not None:
        e.splitk_scratch, e.splitk_scratch_bytes = scratch.data_ptr(), scratch.numel() * 4
    fact_lib.fact_set_flag(b"gemm_pair", pair)
    try:
        L.check(fact_lib.fact_gemm(a_hi.data_ptr(), a_lo.data_ptr() if precise else None, ld, w_hi.data_ptr(),
                                   w_lo.data_ptr() if precise else None, ld, m, n, k, C.byref(e), _st()))
        torch.cuda.synchronize()
    finally:
        fact_lib.fact_set_flag(b"gemm_pair", 1)
    ref = a_eff @ w_eff.t() + bias.double() + pos.double().repeat(m // seq, 1)
    assert rel_err(out.double(), ref) < 3e-5


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("m,n,k,pitch_rows", [(360, 800, 800, 1), (360, 800, 3072, 1), (128, 800, 800, 360),
                                              (5000, 800, 800, 1)])
def test_gemm_residual_with_layernorm(fact_lib, cuda, m, n, k, pitch_rows, precise):
    """BIAS_RESID_F32 with the LayerNorm option: x += a.W^T + b in place, then LayerNorm(x) split to hi / lo.  Small m
    takes the split-K finish kernel with the normalisation fused (flag 1) -- bit-identical to the separate LayerNorm
    launch (flag 0) and to the large-m path, and right against torch.  pitch_rows > 1: operand / residual rows taken
    every pitch_rows-th row (the row-0 tail of the AR step)."""
    g = torch.Generator(device="cpu").manual_seed(9)
    a_full = torch.randn(m * pitch_rows, k, generator=g).to(cuda)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(cuda)
    bias = (0.1 * torch.randn(n, generator=g)).to(cuda)
    gamma = (1 + 0.1 * torch.randn(n, generator=g)).to(cuda)
    beta = (0.1 * torch.randn(n, generator=g)).to(cuda)
    x0 = torch.randn(m * pitch_rows, n, generator=g).to(cuda)
    a_hi, a_lo = split_ref(a_full)
    w_hi, w_lo = split_ref(w)
    scratch = torch.full((16 * min(m, 1024) * n,), float("nan"), device=cuda)
    results = []
    for flag in (1, 0):
        fact_lib.fact_set_flag(b"gemm_finish_ln", flag)
        inplace = pitch_rows == 1
        x = x0.clone()
        out = x if inplace else torch.full((m, n), float("nan"), device=cuda)
        ln_hi = torch.zeros(m, n, dtype=torch.bfloat16, device=cuda)
        ln_lo = torch.zeros_like(ln_hi)
        e = _epi(kind=L.EPI_BIAS_RESID_F32, out_f32=out, ldo=n, bias=bias, resid=x, ldr=n * pitch_rows,
                 ln_gamma=gamma, ln_beta=beta, ln_hi=ln_hi)
        if precise:
            e.ln_lo = ln_lo.data_ptr()
        e.splitk_scratch, e.splitk_scratch_bytes = scratch.data_ptr(), scratch.numel() * 4
        try:
            L.check(fact_lib.fact_gemm(a_hi.data_ptr(), a_lo.data_ptr() if precise else None, k * pitch_rows,
                                       w_hi.data_ptr(), w_lo.data_ptr() if precise else None, k, m, n, k, C.byref(e),
                                       _st()))
            torch.cuda.synchronize()
        finally:
            fact_lib.fact_set_flag(b"gemm_finish_ln", 1)
        results.append((out.clone(), ln_hi.clone(), ln_lo.clone()))
    for t0, t1 in zip(*results):
        assert torch.equal(t0, t1)
    out, ln_hi, ln_lo = results[0]
    rows = slice(None, None, pitch_rows)
    a_eff = (join(a_hi, a_lo) if precise else a_hi.float()).double()[rows]
    w_eff = (join(w_hi, w_lo) if precise else w_hi.float()).double()
    ref = a_eff @ w_eff.t() + bias.double() + x0.double()[rows]
    assert rel_err(out.double(), ref) < 3e-5
    ln_ref = torch.nn.functional.layer_norm(out.double(), (n,), gamma.double(), beta.double(), 1e-5)
    got = (join(ln_hi, ln_lo) if precise else ln_hi.float()).double()
    assert rel_err(got, ln_ref) < (3e-5 if precise else 4e-3)


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("inplace", [True, False])
@pytest.mark.parametrize("m,k", [(46080 // 4, 800), (20000, 3072), (19999, 800)])
def test_gemm_layernorm_inside_the_launch(fact_lib, cuda, m, k, inplace, precise):
    """CTA-pair GEMM whose LayerNorm warps normalise the finished rows inside the launch (ln_sync given, flag
    gemm_fuse_ln = 1): bit-identical to the same GEMM followed by the LayerNorm launch (flag 0), for the in-place
    residual (bulk reductions in the L2) and for out != resid (plain bulk stores); the counters end at zero; and a
    second call on the same counters gives the same bits (they are self-cleaning)."""
    n = 800
    g = torch.Generator(device="cpu").manual_seed(11)
    a = torch.randn(m, k, generator=g).to(cuda)
    w = (torch.randn(n, k, generator=g) / math.sqrt(k)).to(cuda)
    bias = (0.1 * torch.randn(n, generator=g)).to(cuda)
    gamma = (1 + 0.1 * torch.randn(n, generator=g)).to(cuda)
    beta = (0.1 * torch.randn(n, generator=g)).to(cuda)
    x0 = torch.randn(m, n, generator=g).to(cuda)
    a_hi, a_lo = split_ref(a)
    w_hi, w_lo = split_ref(w)
    sync = torch.zeros(2 * ((m + 31) // 32 + 1), dtype=torch.int32, device=cuda)
    results = []
    launches = []
    for flag in (1, 1, 0):
        fact_lib.fact_set_flag(b"gemm_fuse_ln", flag)
        x = x0.clone()
        out = x if inplace else torch.full((m, n), float("nan"), device=cuda)
        ln_hi = torch.zeros(m, n, dtype=torch.bfloat16, device=cuda)
        ln_lo = torch.zeros_like(ln_hi)
        e = _epi(kind=L.EPI_BIAS_RESID_F32, out_f32=out, ldo=n, bias=bias, resid=x, ldr=n, ln_gamma=gamma,
                 ln_beta=beta, ln_hi=ln_hi)
        e.ln_sync = sync.data_ptr()
        if precise:
            e.ln_lo = ln_lo.data_ptr()
        n0 = fact_lib.fact_launch_count()
        try:
            L.check(fact_lib.fact_gemm(a_hi.data_ptr(), a_lo.data_ptr() if precise else None, k, w_hi.data_ptr(),
                                       w_lo.data_ptr() if precise else None, k, m, n, k, C.byref(e), _st()))
            torch.cuda.synchronize()
        finally:
            fact_lib.fact_set_flag(b"gemm_fuse_ln", 1)
        launches.append(fact_lib.fact_launch_count() - n0)
        assert int(sync.abs().sum()) == 0, "arrival counters not left at zero"
        results.append((out.clone(), ln_hi.clone(), ln_lo.clone()))
    assert launches == [1, 1, 2], launches
    for other in results[1:]:
        for t0, t1 in zip(results[0], other):
            assert torch.equal(t0, t1)
    out, ln_hi, ln_lo = results[0]
    ln_ref = torch.nn.functional.layer_norm(out.double(), (n,), gamma.double(), beta.double(), 1e-5)
    got = (join(ln_hi, ln_lo) if precise else ln_hi.float()).double()
    assert rel_err(got, ln_ref) < (3e-5 if precise else 4e-3)


def test_gemm_tc_inplace_residual(fact_lib, cuda):
    """out-proj / FF2 write the residual stream in place (out == resid)."""
    m, n, k = 360, 800, 800
    a = torch.randn(m, k, device=cuda)
    w = torch.randn(k, n, device=cuda) / math.sqrt(k)
    x = torch.randn(m, n, device=cuda)
    x0 = x.clone()
    bias = torch.randn(n, device=cuda)
    a_hi, a_lo = split_ref(a)
    w_hi, w_lo = split_ref(w.t().contiguous())
    e = _epi(kind=L.EPI_BIAS_RESID_F32, out_f32=x, ldo=n, bias=bias, resid=x, ldr=n)
    L.check(fact_lib.fact_gemm(a_hi.data_ptr(), a_lo.data_ptr(), k, w_hi.data_ptr(), w_lo.data_ptr(), k, m, n, k,
                               C.byref(e), _st()))
    ref = join(a_hi, a_lo).double() @ join(w_hi, w_lo).double().t() + bias.double() + x0.double()
    assert rel_err(x, ref) < 3e-5


def test_gemm_big_tiles(fact_lib, cuda):
    """Many CTAs / several waves, K pipeline wraps many times."""
    _gemm_case(fact_lib, cuda, 4096 + 40, 3072, 800, L.EPI_BIAS_GELU_SPLIT, True, seed=3)
    _gemm_case(fact_lib, cuda, 2048, 800, 3072, L.EPI_BIAS_RESID_F32, False, seed=4)


@pytest.mark.parametrize("m", [130, 1100])
@pytest.mark.parametrize("kind", [L.EPI_SPLIT, L.EPI_BIAS_GELU_SPLIT, L.EPI_BIAS_RESID_F32, L.EPI_BIAS_F32])
def test_gemm_f32(fact_lib, cuda, kind, m):
    n, k = 200, 225
    a = torch.randn(m, k, device=cuda)
    w = torch.randn(k, n, device=cuda) / math.sqrt(k)
    bias = torch.randn(n, device=cuda)
    resid = torch.randn(m, n, device=cuda)
    o_hi = torch.zeros(m, n, dtype=torch.bfloat16, device=cuda)
    o_lo = torch.zeros_like(o_hi)
    out = torch.zeros(m, n, device=cuda)
    e = _epi(kind=kind, out_f32=out, out_hi=o_hi, out_lo=o_lo, ldo=n, bias=bias, resid=resid, ldr=n, scale=0.5,
             scale_cols=100)
    L.check(fact_lib.fact_gemm_f32(a.data_ptr(), k, w.data_ptr(), m, n, k, C.byref(e), _st()))
    acc = a.double() @ w.double()
    if kind == L.EPI_SPLIT:
        ref = acc.clone(); ref[:, :100] *= 0.5; got = join(o_hi, o_lo)
    elif kind == L.EPI_BIAS_GELU_SPLIT:
        ref = torch.nn.functional.gelu(acc + bias.double(), approximate="tanh"); got = join(o_hi, o_lo)
    elif kind == L.EPI_BIAS_RESID_F32:
        ref = acc + bias.double() + resid.double(); got = out
    else:
        ref = acc + bias.double(); got = out
    assert rel_err(got, ref) < 2e-5


@pytest.mark.parametrize("legacy", [0, 1])
@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("batch,n,heads,dh", [(2, 120, 10, 80), (1, 240, 10, 80), (2, 360, 10, 80), (3, 37, 2, 16),
                                              (1, 128, 3, 64), (1, 65, 4, 32), (5, 384, 2, 80), (40, 360, 10, 80),
                                              (2, 7, 1, 80)])
def test_sdpa(fact_lib, cuda, batch, n, heads, dh, precise, legacy):
    """legacy=0: tcgen05 kernel where the shape allows (dh 80, n <= 384), else mma.sync; legacy=1: always mma.sync."""
    if legacy and dh != 80:
        pytest.skip("already the mma.sync path")
    fact_lib.fact_set_flag(b"sdpa_legacy", legacy)
    try:
        _sdpa_case(fact_lib, cuda, batch, n, heads, dh, precise)
    finally:
        fact_lib.fact_set_flag(b"sdpa_legacy", 0)


@pytest.mark.parametrize("precise", [True, False])
@pytest.mark.parametrize("batch,n", [(2, 120), (1, 240), (3, 360), (2, 384), (2, 7), (40, 360), (70, 360)])
def test_sdpa_operand_layout_variants(fact_lib, cuda, batch, n, precise):
    """sdpa_wide = 1 (Q / K as a 64-column SWIZZLE_128B box + a 16-column box, default) vs 0 (five 16-column boxes):
    right against torch and bit-identical -- the same products accumulated in the same order.  (70, 360) gives every
    CTA 14 or 15 tiles."""
    outs = []
    for wide in (0, 1):
        fact_lib.fact_set_flag(b"sdpa_wide", wide)
        try:
            outs.append(_sdpa_case(fact_lib, cuda, batch, n, 10, 80, precise))
        finally:
            fact_lib.fact_set_flag(b"sdpa_wide", 1)
    assert torch.equal(outs[1], outs[0])


def _sdpa_case(fact_lib, cuda, batch, n, heads, dh, precise):
    d = heads * dh
    g = torch.Generator(device="cpu").manual_seed(5)
    qkv = torch.randn(batch * n, 3 * d, generator=g).to(cuda)
    qkv[:, :d] *= 1.5   # stands for q * scale * log2(e)
    hi, lo = split_ref(qkv)
    o_hi = torch.zeros(batch * n, d, dtype=torch.bfloat16, device=cuda)
    o_lo = torch.zeros_like(o_hi)
    L.check(fact_lib.fact_sdpa(hi.data_ptr(), lo.data_ptr() if precise else None, o_hi.data_ptr(),
                               o_lo.data_ptr() if precise else None, batch, n, heads, dh, _st()))
    eff = (join(hi, lo) if precise else hi.float()).double().view(batch, n, 3, heads, dh)
    q, k, v = (eff[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    s = q @ k.transpose(-1, -2)                       # log2 domain
    p = torch.softmax(s * math.log(2.0), dim=-1)
    ref = (p @ v).permute(0, 2, 1, 3).reshape(batch * n, d)
    torch.cuda.synchronize()
    got = join(o_hi, o_lo if precise else None).double()
    tol = 5e-5 if precise else 2e-2   # bf16 mode rounds p and the output to 8 bits
    assert (got - ref).abs().max() < tol * max(1.0, float(ref.abs().max())), float((got - ref).abs().max())
    return got


@pytest.mark.parametrize("batch,n_tok,d", [(3, 12, 64), (10, 120, 800)])
def test_embed_and_offsets(fact_lib, cuda, batch, n_tok, d):
    x_len, f = n_tok + 38, 35
    x = torch.randn(batch, x_len, f, device=cuda)
    w = torch.randn(f, d, device=cuda)
    b = torch.randn(d, device=cuda)
    pos = torch.randn(n_tok, d, device=cuda)
    y = torch.empty(batch * n_tok, d, device=cuda)
    step = torch.tensor([7], dtype=torch.int32, device=cuda)
    for sp, start in ((None, 0), (step, 7)):
        L.check(fact_lib.fact_embed(x.data_ptr(), x_len * f, sp.data_ptr() if sp is not None else None, w.data_ptr(),
                                    b.data_ptr(), pos.data_ptr(), y.data_ptr(), batch, n_tok, f, d, _st()))
        ref = x[:, start:start + n_tok].double() @ w.double() + b.double() + pos.double()
        assert rel_err(y.view(batch, n_tok, d), ref) < 1e-5


def test_head_rows(fact_lib, cuda):
    batch, seq, d, od, frames = 4, 9, 800, 225, 5
    x = torch.randn(batch * seq, d, device=cuda)
    w = torch.randn(d, od, device=cuda) * 0.02
    b = torch.randn(od, device=cuda)
    out = torch.zeros(batch, frames, od, device=cuda)
    step = torch.tensor([3], dtype=torch.int32, device=cuda)
    L.check(fact_lib.fact_head_rows(x.data_ptr(), seq, w.data_ptr(), b.data_ptr(), out.data_ptr(), frames * od,
                                    step.data_ptr(), batch, d, od, _st()))
    ref = x.view(batch, seq, d)[:, 0].double() @ w.double() + b.double()
    assert rel_err(out[:, 3], ref) < 1e-5
    assert float(out[:, :3].abs().max()) == 0 and float(out[:, 4:].abs().max()) == 0


def test_mse(fact_lib, cuda):
    b, n, t, od = 5, 36, 20, 225
    pred = torch.randn(b, n, od, device=cuda)
    target = torch.randn(b, t, od, device=cuda)
    loss = torch.zeros((), device=cuda)
    dpred = torch.full((b, n, od), 7.0, device=cuda)
    partial = torch.zeros(1024, device=cuda)
    L.check(fact_lib.fact_mse(target.data_ptr(), pred.data_ptr(), loss.data_ptr(), dpred.data_ptr(),
                              partial.data_ptr(), b, t, n, od, 0.25, _st()))
    p = pred.double().requires_grad_(True)
    ref = ((target.double() - p[:, :t]) ** 2).mean()
    (ref * 0.25).backward()
    assert abs(float(loss) - float(ref)) < 1e-6 * float(ref)
    assert rel_err(dpred, p.grad) < 1e-5


def test_errors_are_reported(fact_lib, cuda):
    e = L.GemmEpilogue()
    rc = fact_lib.fact_gemm(None, None, 8, None, None, 8, 1, 1, 8, C.byref(e), None)
    assert rc == -1 and b"null" in fact_lib.fact_last_error()
    x = torch.zeros(4, 6, device=cuda)
    rc = fact_lib.fact_layernorm_split(x.data_ptr(), None, None, x.data_ptr(), None, 4, 6, None)
    assert rc == -1
