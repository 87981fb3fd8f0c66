"""The C-ABI library loads and exports every symbol include/e3b200.h declares (no compute calls:
this runs without a GPU)."""
import ctypes
import os
import re

from e3b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "e3b200.h")).read()
    return sorted(set(re.findall(r"E3B_API[^;(]*?\b(e3b_\w+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 20
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in e3b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in e3b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.e3b_abi_version() == 1
    assert isinstance(lib.e3b_last_error(), bytes)
    # argument validation happens on the host before any launch
    assert lib.e3b_tp_plan_create(None, None) == -1
    assert b"null" in lib.e3b_last_error()


def test_struct_sizes_match_header():
    # TpDesc: 1 + 1 + 16*2 + 1 + 16*2 + 1 + 96*4 + 1 int32
    assert ctypes.sizeof(_lib.TpDesc) == 4 * (2 + 32 + 1 + 32 + 1 + 384 + 1)
    assert ctypes.sizeof(_lib.GateDesc) % 8 == 0
    lib = _lib.load()
    for which, st in enumerate((_lib.TpDesc, _lib.GateDesc, _lib.GemmProblem, _lib.GemmPackDesc)):
        assert lib.e3b_struct_size(which) == ctypes.sizeof(st), st.__name__
    assert lib.e3b_struct_size(99) == -1


def test_product_refuses_cpu_tensors():
    import pytest
    import torch

    from e3b200 import ops

    with pytest.raises(RuntimeError, match="CUDA"):
        ops.spherical_harmonics(torch.randn(4, 3), 2)
