"""tests/caml_shim/harness.py - TEST INFRASTRUCTURE: call the C stubs of ocaml/nbk_stubs.c from Python.

The stubs are compiled unchanged against this directory's stand-in for the OCaml runtime headers
(caml/*.h, shim.c) into _build/libnbk_ml.so.  `Stubs` builds OCaml values (tagged ints, boxed
doubles, tuples, Bigarray.Array1 over numpy memory), calls a primitive by the name ocaml/nbk.ml
declares - through the native calling convention, or the bytecode one (argv, argn) that exists
for externals of more than five arguments - and decodes the result.  `bind_miniml` plugs the same
calls into an oracle/miniml interpreter, so that ocaml/nbk.ml runs over the real stubs."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "_build", "libnbk_ml.so")
value = C.c_ssize_t


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])
    return LIB


def externals():
    """{ocaml name: (arity, native primitive, bytecode primitive or None)} from ocaml/nbk.ml."""
    import re
    src = open(os.path.join(ROOT, "ocaml", "nbk.ml")).read()
    out = {}
    # arity = the arrows of the declared type outside parentheses
    code = re.sub(r"\(\*.*?\*\)", " ", src, flags=re.S)
    for m in re.finditer(r"external\s+(\w+)\s*:\s*(.*?)=\s*((?:\"\w+\"\s*)+)", code, flags=re.S):
        name, ty, prims = m.group(1), m.group(2), re.findall(r'"(\w+)"', m.group(3))
        depth, arrows = 0, 0
        for k, ch in enumerate(ty):
            depth += ch == "("
            depth -= ch == ")"
            if ch == "-" and ty[k:k + 2] == "->" and depth == 0:
                arrows += 1
        out[name] = (arrows, prims[-1], prims[0] if len(prims) == 2 else None)
    return out


class OCamlFailure(RuntimeError):
    pass


class Stubs:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.shim_val_int.restype, L.shim_val_int.argtypes = value, [C.c_long]
        L.shim_int_val.restype, L.shim_int_val.argtypes = C.c_long, [value]
        L.shim_val_double.restype, L.shim_val_double.argtypes = value, [C.c_double]
        L.shim_double_val.restype, L.shim_double_val.argtypes = C.c_double, [value]
        L.shim_tuple.restype, L.shim_tuple.argtypes = value, [C.c_int, C.POINTER(value)]
        L.shim_field.restype, L.shim_field.argtypes = value, [value, C.c_int]
        L.shim_is_block.restype, L.shim_is_block.argtypes = C.c_int, [value]
        L.shim_wosize.restype, L.shim_wosize.argtypes = C.c_long, [value]
        L.shim_tag.restype, L.shim_tag.argtypes = C.c_int, [value]
        L.shim_bigarray1.restype, L.shim_bigarray1.argtypes = value, [C.c_int, C.c_void_p, C.c_long]
        L.shim_finalize.restype, L.shim_finalize.argtypes = None, [value]
        L.shim_last_error.restype = C.c_char_p
        L.shim_invoke.restype = C.c_int
        L.shim_invoke.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(value), C.POINTER(value)]
        self.ext = externals()
        self.keep = []

    # -- Python -> OCaml value
    def val(self, x):
        L = self.lib
        if x is None or (isinstance(x, tuple) and len(x) == 0):
            return L.shim_val_int(0)
        if isinstance(x, Handle):
            return x.v
        if isinstance(x, (bool, np.bool_)):
            return L.shim_val_int(int(x))
        if isinstance(x, (int, np.integer)):
            return L.shim_val_int(int(x))
        if isinstance(x, (float, np.floating)):
            return L.shim_val_double(float(x))
        if isinstance(x, np.ndarray):
            assert x.flags["C_CONTIGUOUS"] and x.dtype in (np.float64, np.int32), x.dtype
            self.keep.append(x)
            kind = 1 if x.dtype == np.float64 else 6
            return L.shim_bigarray1(kind, x.ctypes.data, x.size)
        if isinstance(x, tuple):
            arr = (value * len(x))(*[self.val(e) for e in x])
            return L.shim_tuple(len(x), arr)
        raise TypeError(type(x))

    # -- OCaml value -> Python (ints stay ints: unit comes back as 0)
    def py(self, v):
        L = self.lib
        if not L.shim_is_block(v):
            return L.shim_int_val(v)
        tag = L.shim_tag(v)
        if tag == 253:
            return L.shim_double_val(v)
        if tag == 255:
            return Handle(v)
        if tag == 0:
            return tuple(self.py(L.shim_field(v, i)) for i in range(L.shim_wosize(v)))
        raise TypeError("block with tag %d" % tag)

    def prim(self, prim_name, args, bytecode=False):
        fn = getattr(self.lib, prim_name)
        a = (value * max(len(args), 1))(*[self.val(x) for x in args])
        r = value(0)
        rc = self.lib.shim_invoke(C.cast(fn, C.c_void_p), int(bytecode), len(args), a, C.byref(r))
        self.keep.clear()
        if rc:
            raise OCamlFailure(self.lib.shim_last_error().decode())
        return self.py(r.value)

    def call(self, name, *args, bytecode=False):
        """Call the external `name` of ocaml/nbk.ml with Python arguments."""
        arity, native, bc = self.ext[name]
        assert len(args) == arity, (name, arity, len(args))
        if bytecode:
            assert bc is not None, name + " has no bytecode entry (<= 5 arguments)"
            return self.prim(bc, args, bytecode=True)
        return self.prim(native, args)

    def finalize(self, h):
        self.lib.shim_finalize(h.v)


class Handle:
    """An opaque OCaml custom block (the ctx)."""

    def __init__(self, v):
        self.v = v


def bind_miniml(it, stubs, bytecode=False):
    """Register every primitive of ocaml/nbk.ml with the interpreter `it`, routed to the real C
    stubs.  miniml's Bigarrays are Python lists: they are copied into numpy memory for the call
    and copied back afterwards (the stubs write their outputs in place)."""
    from oracle.miniml import Builtin

    def conv_in(x, mirrors):
        if isinstance(x, list):
            dt = np.int32 if x and all(isinstance(e, int) and not isinstance(e, bool) for e in x) else np.float64
            arr = np.array(x, dtype=dt)
            mirrors.append((x, arr))
            return arr
        if isinstance(x, tuple) and len(x) > 0:
            return tuple(conv_in(e, mirrors) for e in x)
        return x

    def make(name):
        arity, native, bc = stubs.ext[name]

        def fn(*args):
            mirrors = []
            res = stubs.call(name, *[conv_in(a, mirrors) for a in args], bytecode=bytecode and bc is not None)
            for lst, arr in mirrors:
                lst[:] = [int(e) for e in arr] if arr.dtype == np.int32 else [float(e) for e in arr]
            return res
        return Builtin(fn, arity, name), native, bc
    for name in stubs.ext:
        b, native, bc = make(name)
        it.externals[native] = b
        if bc:
            it.externals[bc] = b
