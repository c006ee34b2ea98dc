"""CPU: the C-ABI library loads and exports every symbol include/nbk.h declares; without a GPU
the product path fails loudly instead of falling back to anything."""
import os
import re

import pytest

import nbk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nbk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nbk_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(nbk.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = nbk.load_library()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.nbk_version()


def test_host_mirror_library_exports_its_header():
    """libnbh (C++ mirror of the OCaml integrators) exports every function of host/nbh.h."""
    import nbh
    src = open(os.path.join(ROOT, "direct-nbody_b200", "host", "nbh.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\b(nbh_[a-z0-9_]+)\s*\(", src)))
    lib = nbh.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert set(nbh.SYMBOLS) <= set(declared)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the failure path is only reachable on a CPU box")
    with pytest.raises(nbk.NbkError) as e:
        nbk.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_oracle():
    """No file under direct-nbody_b200/ may mention the oracle (tier rule ③)."""
    pkg = os.path.join(ROOT, "direct-nbody_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".cc")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), os.path.join(dirpath, f)
