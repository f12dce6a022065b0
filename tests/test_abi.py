"""The C-ABI library: it loads and exports every symbol include/agx.h declares (no compute without a GPU)."""
import ctypes
import pathlib
import re

import pytest

from agimus_controller_b200 import _abi, build

ROOT = pathlib.Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "agx.h").read_text()
    return sorted(set(re.findall(r"\b(agx_[a-z0-9_]+)\s*\(", text)))


def test_header_and_python_mirror_agree():
    assert set(declared_symbols()) == set(_abi.EXPORTED_SYMBOLS)


def test_library_builds_and_exports_every_symbol():
    lib_path = build.build()
    lib = ctypes.CDLL(str(lib_path))
    for name in declared_symbols():
        assert hasattr(lib, name), name
    _abi.bind(lib)
    assert lib.agx_ref_size(7) == _abi.ref_size(7) == 62
    o = _abi.AgxFddpOpts()
    lib.agx_fddp_opts_default(ctypes.byref(o))
    d = _abi.default_fddp_opts()
    for f, _ in _abi.AgxFddpOpts._fields_:
        a, b = getattr(o, f), getattr(d, f)
        assert a == b or (a != a and b != b), f


def test_struct_sizes_match_the_header():
    # agx_model: 2 + 16 + 16 ints, then doubles (8-byte aligned)
    n_d = 16 * 3 + 16 * 9 + 16 * 3 + 16 + 16 * 3 + 16 * 6 + 16 + 3 + 9 + 3 + 4 * 3 + 4 * 3 + 4 + 1
    assert ctypes.sizeof(_abi.AgxModel) == 34 * 4 + n_d * 8 + 12 * 4
    assert ctypes.sizeof(_abi.AgxFddpOpts) == 11 * 8 + 4 * 4 + 8
    assert ctypes.sizeof(_abi.AgxSqpOpts) == 4 * 8 + 2 * 4 + 8


def test_product_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from agimus_controller_b200 import panda_table
    from agimus_controller_b200.solver import BatchedShootingProblem

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        BatchedShootingProblem(panda_table(), [0.01] * 3, 1)
