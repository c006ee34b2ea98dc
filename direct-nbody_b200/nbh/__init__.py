"""nbh — Python (ctypes) binding of libnbh, the C++ host mirror of the reference's OCaml
modules (direct-nbody_b200/host/nbh.h).  Test / bench harness only; the integrator logic and
every pair loop live in the native libraries (libnbh -> libnbk -> CUDA)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

import nbk

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG), "lib", "libnbh.so")

_d, _i = C.c_double, C.c_int
_pd = C.POINTER(C.c_double)
_vp = C.c_void_p
_u64 = C.c_uint64

SYMBOLS = {
    "nbh_last_error": (C.c_char_p, []),
    "nbh_make_plummer": (_i, [_i, _u64, _pd, _pd, _pd]),
    "nbh_make_hot_spherical": (_i, [_i, _u64, _pd, _pd, _pd]),
    "nbh_make_cold_spherical": (_i, [_i, _u64, _pd, _pd, _pd]),
    "nbh_make_uniform_box": (_i, [_i, _u64, _pd, _pd, _pd]),
    "nbh_adjust_frame": (_i, [_i, _pd, _pd, _pd]),
    "nbh_add_mass_disc": (_i, [_i, _u64, _d, _d, _pd, _pd]),
    "nbh_rescale_to_standard_units": (_i, [_vp, _i, _pd, _pd, _pd, _d]),
}


class NbhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libnbh error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is None:
        nbk.load_library()  # libnbh links libnbk
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} not found: run __graft_entry__.build()")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ck(rc):
    if rc != 0:
        raise NbhError(rc, load_library().nbh_last_error().decode())


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_pd)


def _make(kind, n, seed):
    m, q, p = np.empty(n), np.empty((n, 3)), np.empty((n, 3))
    _ck(getattr(load_library(), "nbh_make_" + kind)(n, seed, _p(m), _p(q), _p(p)))
    return m, q, p


def make_plummer(n, seed):
    """Ics.make_plummer (ics.ml:211-227), seeded."""
    return _make("plummer", n, seed)


def make_hot_spherical(n, seed):
    return _make("hot_spherical", n, seed)


def make_cold_spherical(n, seed):
    return _make("cold_spherical", n, seed)


def make_uniform_box(n, seed):
    return _make("uniform_box", n, seed)


def adjust_frame(m, q, p):
    m, q, p = (np.array(a, dtype=np.float64, order="C") for a in (m, q, p))
    _ck(load_library().nbh_adjust_frame(len(m), _p(m), _p(q), _p(p)))
    return m, q, p


def add_mass_disc(m, q, p, seed, heavyfrac=0.01, heavyfac=10.0):
    m, q, p = (np.array(a, dtype=np.float64, order="C") for a in (m, q, p))
    _ck(load_library().nbh_add_mass_disc(len(m), seed, heavyfrac, heavyfac, _p(m), _p(p)))
    return m, q, p


def rescale_to_standard_units(ctx: nbk.Context, m, q, p, eps2=0.0):
    m, q, p = (np.array(a, dtype=np.float64, order="C") for a in (m, q, p))
    _ck(load_library().nbh_rescale_to_standard_units(ctx._h, len(m), _p(m), _p(q), _p(p), eps2))
    return m, q, p
