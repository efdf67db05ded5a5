"""The C-ABI library loads and exports exactly what include/fact_sm100.h declares (no compute calls: CPU-only)."""
import ctypes
import os
import re

from mint_b200 import lib as L

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "fact_sm100.h")


def _declared():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"FACT_API\s+[\w\s\*]+?\b(fact_\w+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    names = _declared()
    for must in ("fact_forward", "fact_infer_auto_regressive", "fact_gemm", "fact_sdpa", "fact_layernorm_split",
                 "fact_embed", "fact_head_rows", "fact_mse", "fact_pack_weight", "fact_workspace_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol(fact_lib):
    raw = ctypes.CDLL(L.LIB_PATH)
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in fact_sm100.h but not exported"


def test_ctypes_table_covers_the_header(fact_lib):
    from mint_b200 import lib_bwd
    bound = set(L.SIGNATURES) | set(lib_bwd.SIGNATURES)
    assert set(_declared()) == bound


def test_abi_version_and_error_string(fact_lib):
    assert fact_lib.fact_abi_version() == L.ABI_VERSION == 4
    assert fact_lib.fact_set_flag(b"no_such_flag", 1) == -5
    assert b"unknown flag" in fact_lib.fact_last_error()
    d = L.Dims(800, 10, 3072, 2, 2, 12, 120, 240, 225, 35, 225)
    need = fact_lib.fact_workspace_bytes(ctypes.byref(d), 128, L.MODE_PRECISE)
    assert 1.0e9 < need < 2.5e9          # ~1.6 GB of activations at batch 128
    assert fact_lib.fact_workspace_bytes(ctypes.byref(d), 128, L.MODE_BF16) < need


def test_ctypes_structs_match_the_c_layout(tmp_path):
    """The header is plain C (gcc, no CUDA headers) and every struct the ctypes layer mirrors has the same size and
    field offsets as the C compiler gives it -- a silent mismatch would corrupt pointers across the boundary."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    assert gcc, "gcc is part of the image"
    structs = {"fact_gemm_epilogue": L.GemmEpilogue, "fact_dims": L.Dims, "fact_layer_weights": L.LayerWeights,
               "fact_weights": L.Weights, "fact_layer_grads": L.LayerGrads, "fact_grads": L.Grads}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, what, value = line.split()
        ct = structs[cname]
        expect = ctypes.sizeof(ct) if what == "size" else getattr(ct, what).offset
        assert int(value) == expect, f"{cname}.{what}: C {value} != ctypes {expect}"
        seen += 1
    assert seen == sum(len(ct._fields_) + 1 for ct in structs.values())


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        L.load()
    except RuntimeError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must raise when the CUDA library is absent")
