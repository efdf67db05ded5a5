"""ctypes signatures of the training entry points of libfact_sm100.so (bound when present)."""
from __future__ import annotations

import ctypes as C

SIGNATURES: dict = {}


def bind(lib) -> None:
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
