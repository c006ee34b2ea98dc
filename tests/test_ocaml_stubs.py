"""CPU: ocaml/nbk_stubs.c, COMPILED and CALLED.

The image has no OCaml, so the C side of the OCaml binding had never seen a compiler.
tests/caml_shim/ holds a stand-in for <caml/*.h> written from the OCaml manual's description of the
value representation (tagged ints, boxed doubles, tuples, custom blocks, Bigarray); against it the
stubs compile UNCHANGED, with include/nbk.h in scope, so every call they make into the C ABI is
checked by the compiler (argument count, pointer and scalar types), and the resulting library can
be driven from Python.  Without a GPU: symbols, arities, the constant-returning primitives and the
exception path (Failure carrying nbk_last_error).  With one: tests/test_ocaml_stubs_gpu.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "caml_shim"))
import harness  # noqa: E402


@pytest.fixture(scope="module")
def stubs():
    return harness.Stubs()


def test_stubs_compile_unchanged_against_the_abi_with_strict_warnings():
    subprocess.check_call(["make", "-s", "-B", "-C", os.path.join(ROOT, "tests", "caml_shim")])
    assert os.path.exists(harness.LIB)


def test_every_external_has_its_stub_with_the_declared_arity(stubs):
    import re
    text = open(os.path.join(ROOT, "ocaml", "nbk_stubs.c")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    params = {m.group(1): m.group(2) for m in
              re.finditer(r"CAMLprim\s+value\s+(\w+)\s*\(([^)]*)\)", text)}
    exported = subprocess.check_output(["nm", "-D", "--defined-only", harness.LIB], text=True)
    assert len(stubs.ext) >= 56
    for name, (arity, native, bc) in stubs.ext.items():
        assert f" T {native}\n" in exported, native
        n_native = len([p for p in params[native].split(",") if p.strip()])
        if arity > 5:
            assert bc is not None, f"{name}: {arity} arguments need a bytecode entry"
            assert f" T {bc}\n" in exported, bc
            assert re.sub(r"\s+", " ", params[bc].strip()) in ("value *a, int n", "value *argv, int argn"), bc
            # the native entry takes all the arguments, or IS the (argv, argn) function's delegate
            assert n_native == arity, (name, n_native, arity)
        else:
            assert bc is None and n_native == max(arity, 1), (name, n_native, arity)


def test_constant_primitives_and_value_representation(stubs):
    sys.path.insert(0, os.path.join(ROOT, "direct-nbody_b200"))
    import nbk
    lib = nbk.load_library()
    assert stubs.call("lf_subsystem_max", None) == lib.nbk_lf_subsystem_max() > 0
    assert stubs.call("adv_range_max", None) == lib.nbk_adv_range_max() > 0
    assert stubs.call("hermite_resident_max", None) == lib.nbk_hermite_resident_max() >= 12288
    # round trips of the representation itself
    assert stubs.py(stubs.val(-7)) == -7 and stubs.py(stubs.val(2.5)) == 2.5
    assert stubs.py(stubs.val((1, 2.0, (3, 4.5)))) == (1, 2.0, (3, 4.5))


def test_errors_surface_as_ocaml_failure(stubs):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine WITHOUT a GPU (nbk_create must fail)")
    with pytest.raises(harness.OCamlFailure) as ei:
        stubs.call("create", 0)
    assert "Failure:" in str(ei.value) and len(str(ei.value)) > len("Failure: ")
