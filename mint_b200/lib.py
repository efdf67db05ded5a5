"""ctypes binding of libfact_sm100.so (the C ABI declared in include/fact_sm100.h).

There is no CPU fallback: if the library is missing, importing callers get a RuntimeError telling them to build it
(`python -m mint_b200.build`, or `__graft_entry__.build()`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfact_sm100.so")

MODE_PRECISE, MODE_BF16, MODE_FP32_SIMT = 0, 1, 2
MODES = {"precise": MODE_PRECISE, "bf16": MODE_BF16, "fp32_simt": MODE_FP32_SIMT}
EPI_SPLIT, EPI_BIAS_GELU_SPLIT, EPI_BIAS_RESID_F32, EPI_BIAS_F32, EPI_BIAS_GELU_SAVE, EPI_GELU_GRAD = 0, 1, 2, 3, 4, 5

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float


class LayerWeights(C.Structure):
    _fields_ = [(n, _vp) for n in (
        "ln1_gamma", "ln1_beta", "wqkv_hi", "wqkv_lo", "wo_hi", "wo_lo", "bo", "ln2_gamma", "ln2_beta",
        "w1_hi", "w1_lo", "b1", "w2_hi", "w2_lo", "b2", "wqkv_f32", "wo_f32", "w1_f32", "w2_f32",
        "wqkv_kl", "wo_kl", "w1_kl", "w2_kl")]


class LayerGrads(C.Structure):
    _fields_ = [(n, _vp) for n in (
        "ln1_gamma", "ln1_beta", "wqkv", "wo", "bo", "ln2_gamma", "ln2_beta", "w1", "b1", "w2", "b2")]


class Dims(C.Structure):
    _fields_ = [(n, _i) for n in (
        "d_model", "n_heads", "d_ff", "motion_layers", "audio_layers", "cross_layers", "motion_seq", "audio_seq",
        "motion_dim", "audio_dim", "out_dim")]


class Weights(C.Structure):
    _fields_ = [("motion_layers", C.POINTER(LayerWeights)), ("audio_layers", C.POINTER(LayerWeights)),
                ("cross_layers", C.POINTER(LayerWeights))] + [(n, _vp) for n in (
        "motion_embed_w", "motion_embed_b", "motion_pos", "audio_embed_w", "audio_embed_b", "audio_pos",
        "out_w", "out_b", "out_w_hi", "out_w_lo", "out_w_kl")]


class Grads(C.Structure):
    _fields_ = [("motion_layers", C.POINTER(LayerGrads)), ("audio_layers", C.POINTER(LayerGrads)),
                ("cross_layers", C.POINTER(LayerGrads))] + [(n, _vp) for n in (
        "motion_embed_w", "motion_embed_b", "motion_pos", "audio_embed_w", "audio_embed_b", "audio_pos",
        "out_w", "out_b")]


ABI_VERSION = 4  # FACT_ABI_VERSION of include/fact_sm100.h these ctypes declarations mirror


class GemmEpilogue(C.Structure):
    _fields_ = [("kind", _i), ("out_f32", _vp), ("out_hi", _vp), ("out_lo", _vp), ("ldo", _i), ("bias", _vp),
                ("resid", _vp), ("ldr", _i), ("scale", _f), ("scale_cols", _i), ("seq_in", _i), ("seq_out", _i),
                ("seq_off", _i), ("aux", _vp), ("ldaux", _i), ("splitk_scratch", _vp), ("splitk_scratch_bytes", C.c_size_t),
                ("colsum", _vp), ("resid_rows", _i),
                ("ln_gamma", _vp), ("ln_beta", _vp), ("ln_hi", _vp), ("ln_lo", _vp), ("ln_sync", _vp)]


# symbol -> (restype, argtypes); this table is also what tests/test_abi.py checks against include/fact_sm100.h
SIGNATURES = {
    "fact_abi_version": (_i, []),
    "fact_last_error": (C.c_char_p, []),
    "fact_launch_count": (C.c_longlong, []),
    "fact_pack_weight": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "fact_pack_weights": (_i, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _vp]),
    "fact_layernorm_split": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
    "fact_gemm": (_i, [_vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, C.POINTER(GemmEpilogue), _vp]),
    "fact_gemm_f32": (_i, [_vp, _i, _vp, _i, _i, _i, C.POINTER(GemmEpilogue), _vp]),
    "fact_sdpa": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fact_set_flag": (_i, [C.c_char_p, _i]),
    "fact_embed": (_i, [_vp, _ll, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "fact_head_rows": (_i, [_vp, _ll, _vp, _vp, _vp, _ll, _vp, _i, _i, _i, _vp]),
    "fact_mse": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "fact_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), _i, _i]),
    "fact_forward": (_i, [C.POINTER(Dims), C.POINTER(Weights), _vp, _vp, _vp, _i, _vp, C.c_size_t, _i, _vp]),
    "fact_infer_auto_regressive": (_i, [C.POINTER(Dims), C.POINTER(Weights), _vp, _i, _vp, _i, _i, _i, _i, _vp,
                                        _vp, C.c_size_t, _i, _i, _vp, _vp]),
    "fact_ar_session_create": (_vp, []),
    "fact_ar_session_destroy": (_i, [_vp]),
    "fact_ar_session_graphs": (_i, [_vp]),
}

_lib = None


class FactError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the FACT hot path has no CPU fallback. Build it with "
            "`python -m mint_b200.build` (needs nvcc; cross-compiles for sm_100a without a GPU).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.fact_abi_version() != ABI_VERSION:
        raise FactError(f"{LIB_PATH} has ABI version {lib.fact_abi_version()}, this package needs {ABI_VERSION}: "
                        "rebuild with `python -m mint_b200.build --force`")
    from . import lib_bwd  # optional training symbols (same .so)
    lib_bwd.bind(lib)
    # developer aid: FACT_FLAGS="gemm_tma_store=0,gemm_pair=1" -> fact_set_flag for A/B timing of kernel variants
    for item in filter(None, os.environ.get("FACT_FLAGS", "").split(",")):
        name, _, value = item.partition("=")
        if lib.fact_set_flag(name.strip().encode(), int(value)) != 0:
            raise FactError(f"FACT_FLAGS: unknown flag {name!r}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().fact_last_error()
        raise FactError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()
