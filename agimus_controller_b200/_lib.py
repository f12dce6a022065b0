"""Loader of the product CUDA library (``libagx.so``, C ABI of ``include/agx.h``).

There is no CPU fallback: a missing library or a missing CUDA device is an error.
"""
from __future__ import annotations

import ctypes as C

from . import _abi
from .build import LIB

_LIB = None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        import os
        import pathlib

        path = pathlib.Path(os.environ.get("AGX_LIBRARY", LIB))  # tuning hook: another build of the same library
        if not path.exists():
            raise RuntimeError(
                f"{path} is missing: build it with `python -m agimus_controller_b200.build` "
                "(the solve path has no CPU fallback)")
        _LIB = _abi.bind(C.CDLL(str(path)))
    return _LIB
