"""CPU: the C-ABI library loads and exports every symbol include/onepeace_b200.h declares, and the ctypes
signature table covers exactly that set (no compute calls without a GPU)."""
import ctypes
import os
import re

from one_peace_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "onepeace_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(opb_[a-z0-9_]+)\s*\(", src))


def test_library_loads_and_exports_all_declared_symbols():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.opb_abi_version() >= 1
    assert lib.opb_status_string(0) == b"ok"


def test_ctypes_table_matches_header():
    assert set(_lib.SIGNATURES) == declared_symbols()


def test_invalid_arguments_are_rejected_without_a_gpu():
    lib = _lib.load()
    # null pointers / bad shapes are refused before any CUDA call
    assert lib.opb_layernorm(None, 0, 8, None, 1, 8, None, None, 4, 8, ctypes.c_float(1e-5), 0, 0, 0, 0, 0, 0, 0, 0, 0, None) == 1
    assert lib.opb_attention_fwd(None, None, None, None, None, None, 1, 1, 1, 0, 0, None) == 1
