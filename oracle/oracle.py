"""oracle/oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loader for oracle/_build/liboracle.so, the CPU restatement of the reference's
pairwise-gravity path (arithmetic and citations: oracle/base.h; parity status: reproduces bit for
bit what the reference's unmodified sources compute when interpreted by oracle/miniml — see base.h
and oracle/ref_fixtures/README.md; no ocamlopt run exists).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  Nothing under direct-nbody_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
ADV_STRIDE = 35

_d = C.c_double
_i = C.c_int
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int)
_pl = C.POINTER(C.c_long)
EXTPOT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, _pd, _pd, _pd)


def build(force: bool = False) -> str:
    """Compile the restatement with oracle/Makefile (gcc/g++, -ffp-contract=off)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


def _declare(L):
    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("oracle_num_threads", _i)
    sig("oracle_set_num_threads", None, _i)
    sig("oracle_set_step_cap", None, C.c_long)
    sig("oracle_start_bs_sweep", None, _i, _pd, _pd, _pd, _d, _pd, _pd)
    sig("oracle_start_bs_sweep_omp", None, _i, _pd, _pd, _pd, _d, _pd, _pd)
    sig("oracle_grad_v_sum", None, _i, _pd, _pd, _d, _i, _i, _i, _i, _pd)
    sig("oracle_grad_v_dgrad_v_sum", None, _i, _pd, _pd, _pd, _d, _i, _i, _i, _i, _pd, _pd)
    sig("oracle_total_potential_energy", _d, _i, _pd, _pd, _d)
    sig("oracle_total_kinetic_energy", _d, _i, _pd, _pd)
    sig("oracle_energy", _d, _i, _pd, _pd, _pd, _d)
    sig("oracle_v_sum", None, _i, _pd, _pd, _d, _i, _i, _i, _i, _pd)
    sig("oracle_bodies_to_el_phase_space", None, _i, _pd, _pd, _pd, _d, _pd)
    sig("oracle_kick_all_pairs", None, _i, _pd, _pd, _pd, _pd, _d, _d)
    sig("oracle_kick_block", None, _i, _pd, _pd, _pd, _pd, _d, _d, _i, _i)
    sig("oracle_kick_sum", None, _i, _pd, _pd, _pd, _d, _d, _i, _i, _i, _i, _pd, _pd)
    sig("oracle_binaries", _i, _i, _pd, _pd, _pd, _d, _d, _i, _pi, _pi, _pd)
    sig("oracle_grad_v_dgrad_v_sum_quad", None, _i, _pd, _pd, _pd, _d, _i, _i, _i, _i, _pd, _pd)
    sig("oracle_v_sum_quad", None, _i, _pd, _pd, _d, _i, _i, _i, _i, _pd)
    sig("oracle_total_potential_energy_quad", _d, _i, _pd, _pd, _d)
    sig("oracle_adjust_frame", None, _i, _pd, _pd, _pd)
    for nm in ("plummer", "hot_spherical", "cold_spherical", "uniform_box"):
        sig("oracle_make_" + nm, None, _i, C.c_uint64, _pd, _pd, _pd)
    sig("oracle_add_mass_disc", None, _i, C.c_uint64, _d, _d, _pd, _pd)
    sig("oracle_rescale_to_standard_units", _i, _i, _pd, _pd, _pd, _d)
    sig("oracle_hermite_advance", None, _i, _pi, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _d, _d, _d, _pl)
    sig("oracle_hermite_advance_body", None, _i, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _i, _d, _d, _pd)
    sig("oracle_dfloat_hermite_advance_body", None, _i, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _pd,
        _pd, _pd, _pd, _pd, _i, _d, _d, _pd, _pd)
    sig("oracle_grad_v_dgrad_v_sum_tangent", None, _i, _pd, _pd, _pd, _pd, _pd, _d, _i, _i, _i, _i,
        _pd, _pd, _pd, _pd)
    sig("oracle_grad_v_sum_tangent", None, _i, _pd, _pd, _pd, _d, _i, _i, _i, _i, _pd, _pd)
    sig("oracle_hermite_jacobian_advance", None, _i, _pd, _pd, _pd, _d, _d, _d, _pd, _pd)
    sig("oracle_omega_pushforward", _d, _i, _pd)
    sig("oracle_leapfrog_advance", None, _i, _pi, _pd, _pd, _pd, _pd, _pd, _d, _d, _pl)
    sig("oracle_leapfrog_advance_const_dt", None, _i, _pi, _pd, _pd, _pd, _pd, _pd, _d, _i)
    sig("oracle_omelyan_advance", None, _i, _pd, _pd, _pd, _pd, _d, _d)
    sig("oracle_advancer_advance", None, _i, _pi, _pd, _d, _d, _d, C.c_void_p, C.c_void_p, _pl)
    sig("oracle_extpot_plummer_test_energy", _d, _i, _pd, _pd)


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_pd)


def _ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_pi)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def set_step_cap(cap: int) -> None:
    lib().oracle_set_step_cap(cap)


def num_threads() -> int:
    return lib().oracle_num_threads()


def set_num_threads(n: int) -> None:
    """Thread count of the OpenMP sweeps from here on (ignores OMP_NUM_THREADS)."""
    lib().oracle_set_num_threads(int(n))


def host_cores() -> int:
    """Cores this process may run on (the GPU box's host cores)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ------------------------------------------------------------------ initial conditions
def _make(kind, n, seed):
    m, q, p = np.empty(n), np.empty((n, 3)), np.empty((n, 3))
    getattr(lib(), "oracle_make_" + kind)(n, seed, _p(m), _p(q), _p(p))
    return m, q, p


def make_plummer(n, seed):
    return _make("plummer", n, seed)


def make_hot_spherical(n, seed):
    return _make("hot_spherical", n, seed)


def make_cold_spherical(n, seed):
    return _make("cold_spherical", n, seed)


def make_uniform_box(n, seed):
    return _make("uniform_box", n, seed)


def adjust_frame(m, q, p):
    m, q, p = _f(m).copy(), _f(q).copy(), _f(p).copy()
    lib().oracle_adjust_frame(len(m), _p(m), _p(q), _p(p))
    return m, q, p


def add_mass_disc(m, q, p, seed, heavyfrac=0.01, heavyfac=10.0):
    m, q, p = _f(m).copy(), _f(q).copy(), _f(p).copy()
    lib().oracle_add_mass_disc(len(m), seed, heavyfrac, heavyfac, _p(m), _p(p))
    return m, q, p


def rescale_to_standard_units(m, q, p, eps2=0.0):
    m, q, p = _f(m).copy(), _f(q).copy(), _f(p).copy()
    rc = lib().oracle_rescale_to_standard_units(len(m), _p(m), _p(q), _p(p), eps2)
    if rc != 0:
        raise AssertionError("rescale_to_standard_units: e < 0.0 failed (ics.ml:318)")
    return m, q, p


# ------------------------------------------------------------------ sweeps
def start_bs_sweep(m, q, p, eps2=0.0, omp=False):
    m, q, p = _f(m), _f(q), _f(p)
    n = len(m)
    f, df = np.empty((n, 3)), np.empty((n, 3))
    fn = lib().oracle_start_bs_sweep_omp if omp else lib().oracle_start_bs_sweep
    fn(n, _p(m), _p(q), _p(p), eps2, _p(f), _p(df))
    return f, df


def _ranges(n, targets, sources):
    t0, t1 = targets if targets is not None else (0, n)
    s0, s1 = sources if sources is not None else (0, n)
    return t0, t1, s0, s1


def grad_v_sum(m, q, eps2=0.0, targets=None, sources=None):
    m, q = _f(m), _f(q)
    n = len(m)
    t0, t1, s0, s1 = _ranges(n, targets, sources)
    g = np.empty((t1 - t0, 3))
    lib().oracle_grad_v_sum(n, _p(m), _p(q), eps2, t0, t1, s0, s1, _p(g))
    return g


def grad_v_dgrad_v_sum(m, q, p, eps2=0.0, targets=None, sources=None, quad=False):
    m, q, p = _f(m), _f(q), _f(p)
    n = len(m)
    t0, t1, s0, s1 = _ranges(n, targets, sources)
    g, dg = np.empty((t1 - t0, 3)), np.empty((t1 - t0, 3))
    fn = lib().oracle_grad_v_dgrad_v_sum_quad if quad else lib().oracle_grad_v_dgrad_v_sum
    fn(n, _p(m), _p(q), _p(p), eps2, t0, t1, s0, s1, _p(g), _p(dg))
    return g, dg


def v_sum(m, q, eps2=0.0, targets=None, sources=None, quad=False):
    m, q = _f(m), _f(q)
    n = len(m)
    t0, t1, s0, s1 = _ranges(n, targets, sources)
    phi = np.empty(t1 - t0)
    fn = lib().oracle_v_sum_quad if quad else lib().oracle_v_sum
    fn(n, _p(m), _p(q), eps2, t0, t1, s0, s1, _p(phi))
    return phi


def total_potential_energy(m, q, eps2=0.0, quad=False):
    m, q = _f(m), _f(q)
    fn = (lib().oracle_total_potential_energy_quad if quad
          else lib().oracle_total_potential_energy)
    return fn(len(m), _p(m), _p(q), eps2)


def total_kinetic_energy(m, p):
    m, p = _f(m), _f(p)
    return lib().oracle_total_kinetic_energy(len(m), _p(m), _p(p))


def energy(m, q, p, eps2=0.0):
    m, q, p = _f(m), _f(q), _f(p)
    return lib().oracle_energy(len(m), _p(m), _p(q), _p(p), eps2)


def bodies_to_el_phase_space(m, q, p, eps2=0.0):
    m, q, p = _f(m), _f(q), _f(p)
    el = np.empty((len(m), 4))
    lib().oracle_bodies_to_el_phase_space(len(m), _p(m), _p(q), _p(p), eps2, _p(el))
    return el


def kick_all_pairs(m, q, v, dt_max, eta, dt):
    m, q = _f(m), _f(q)
    v, dt_max = _f(v).copy(), _f(dt_max).copy()
    lib().oracle_kick_all_pairs(len(m), _p(m), _p(q), _p(v), _p(dt_max), eta, dt)
    return v, dt_max


def kick_block(m, q, v, dt_max, eta, dt, ibegin, iend):
    m, q = _f(m), _f(q)
    v, dt_max = _f(v).copy(), _f(dt_max).copy()
    lib().oracle_kick_block(len(m), _p(m), _p(q), _p(v), _p(dt_max), eta, dt, ibegin, iend)
    return v, dt_max


def kick_sum(m, q, v, eta, dt, targets=None, sources=None):
    m, q, v = _f(m), _f(q), _f(v)
    n = len(m)
    t0, t1, s0, s1 = _ranges(n, targets, sources)
    dv, tmin = np.empty((t1 - t0, 3)), np.empty(t1 - t0)
    lib().oracle_kick_sum(n, _p(m), _p(q), _p(v), eta, dt, t0, t1, s0, s1, _p(dv), _p(tmin))
    return dv, tmin


def binaries(m, q, p, eps2=0.0, threshold=0.0, max_pairs=1 << 20):
    """Analysis.binaries / binaries_threshold (analysis.ml:836-861): (i, j, e) in loop order."""
    m, q, p = _f(m), _f(q), _f(p)
    pi, pj = np.empty(max_pairs, dtype=np.int32), np.empty(max_pairs, dtype=np.int32)
    e = np.empty(max_pairs)
    k = lib().oracle_binaries(len(m), _p(m), _p(q), _p(p), eps2, threshold, max_pairs, _ip(pi),
                              _ip(pj), _p(e))
    return pi[:k].copy(), pj[:k].copy(), e[:k].copy()


def grad_v_dgrad_v_sum_tangent(m, q, p, dq, dp, eps2=0.0, targets=None, sources=None):
    m, q, p, dq, dp = _f(m), _f(q), _f(p), _f(dq), _f(dp)
    n = len(m)
    t0, t1, s0, s1 = _ranges(n, targets, sources)
    outs = [np.empty((t1 - t0, 3)) for _ in range(4)]
    lib().oracle_grad_v_dgrad_v_sum_tangent(n, _p(m), _p(q), _p(p), _p(dq), _p(dp), eps2,
                                            t0, t1, s0, s1, *[_p(o) for o in outs])
    return tuple(outs)


def grad_v_sum_tangent(m, q, dq, eps2=0.0, targets=None, sources=None):
    m, q, dq = _f(m), _f(q), _f(dq)
    n = len(m)
    t0, t1, s0, s1 = _ranges(n, targets, sources)
    g, gt = np.empty((t1 - t0, 3)), np.empty((t1 - t0, 3))
    lib().oracle_grad_v_sum_tangent(n, _p(m), _p(q), _p(dq), eps2, t0, t1, s0, s1, _p(g), _p(gt))
    return g, gt


# ------------------------------------------------------------------ integrators
class HermiteState:
    """Flat image of a `Hermite.b array` (hermite.ml:20-29); fresh bodies have h = -1."""

    def __init__(self, m, q, p, t=0.0):
        n = len(m)
        self.id = np.arange(1, n + 1, dtype=np.int32)
        self.t = np.full(n, float(t))
        self.m, self.q, self.p = _f(m).copy(), _f(q).copy(), _f(p).copy()
        self.h = np.full(n, -1.0)
        self.f, self.df = np.zeros((n, 3)), np.zeros((n, 3))
        self.nsteps = 0


def hermite_advance(s: HermiteState, dt, sf, eps2=0.0) -> HermiteState:
    import copy
    o = copy.deepcopy(s)
    ns = C.c_long(o.nsteps)
    lib().oracle_hermite_advance(len(o.m), _ip(o.id), _p(o.t), _p(o.m), _p(o.q), _p(o.p), _p(o.h),
                                 _p(o.f), _p(o.df), dt, sf, eps2, C.byref(ns))
    o.nsteps = ns.value
    return o


def hermite_advance_body(t, m, q, p, h, f, df, ib, eta, eps2=0.0):
    arrs = [_f(a) for a in (t, m, q, p, h, f, df)]
    out = np.empty(14)
    lib().oracle_hermite_advance_body(len(arrs[1]), *[_p(a) for a in arrs], ib, eta, eps2, _p(out))
    return dict(t=out[0], h=out[1], q=out[2:5], p=out[5:8], f=out[8:11], df=out[11:14])


def dfloat_hermite_advance_body(t, m, q, p, h, f, df, t_d, q_d, p_d, h_d, f_d, df_d, ib, eta,
                                eps2=0.0):
    """DFloat_hermite.advance_body (dFloat_hermite.ml:115-150), one tangent direction:
    returns (values, tangents) dicts of the new body."""
    vals = [_f(a) for a in (t, m, q, p, h, f, df)]
    tans = [_f(a) for a in (t_d, q_d, p_d, h_d, f_d, df_d)]
    out, out_d = np.empty(14), np.empty(14)
    lib().oracle_dfloat_hermite_advance_body(len(vals[1]), *[_p(a) for a in vals],
                                             *[_p(a) for a in tans], ib, eta, eps2, _p(out),
                                             _p(out_d))
    unpack = lambda o: dict(t=o[0], h=o[1], q=o[2:5], p=o[5:8], f=o[8:11], df=o[11:14])
    return unpack(out), unpack(out_d)


def hermite_jacobian_advance(m, q, p, dt, sf, eps2=0.0):
    m, q, p = _f(m), _f(q), _f(p)
    n = len(m)
    jac, out = np.empty((6 * n, 6 * n)), np.empty(6 * n)
    lib().oracle_hermite_jacobian_advance(n, _p(m), _p(q), _p(p), dt, sf, eps2, _p(jac), _p(out))
    return jac, out


def omega_pushforward(jac):
    jac = _f(jac)
    return lib().oracle_omega_pushforward(jac.shape[0] // 6, _p(jac))


class LeapfrogState:
    """Flat image of a `Leapfrog.b array` (leapfrog.ml:5-12): Leapfrog.make sets t = 0,
    dt_max = 0, v = p / m (leapfrog.ml:28-34)."""

    def __init__(self, m, q, p):
        n = len(m)
        self.id = np.arange(1, n + 1, dtype=np.int32)
        self.t, self.dt_max = np.zeros(n), np.zeros(n)
        self.m, self.q = _f(m).copy(), _f(q).copy()
        self.v = _f(p) / self.m[:, None]
        self.nkicks = 0

    @property
    def p(self):  # leapfrog.ml:20
        return self.v * self.m[:, None]


def leapfrog_advance(s: LeapfrogState, dt, eta) -> LeapfrogState:
    import copy
    o = copy.deepcopy(s)
    nk = C.c_long(0)
    lib().oracle_leapfrog_advance(len(o.m), _ip(o.id), _p(o.t), _p(o.dt_max), _p(o.m), _p(o.q),
                                  _p(o.v), dt, eta, C.byref(nk))
    o.nkicks = nk.value
    return o


def leapfrog_advance_const_dt(s: LeapfrogState, dt, nsteps) -> LeapfrogState:
    import copy
    o = copy.deepcopy(s)
    lib().oracle_leapfrog_advance_const_dt(len(o.m), _ip(o.id), _p(o.t), _p(o.dt_max), _p(o.m),
                                           _p(o.q), _p(o.v), dt, nsteps)
    return o


def omelyan_advance(m, q, p, t, h, eps2=0.0):
    m = _f(m)
    q, p = _f(q).copy(), _f(p).copy()
    tt = C.c_double(t)
    lib().oracle_omelyan_advance(len(m), _p(m), _p(q), _p(p), C.byref(tt), h, eps2)
    return q, p, tt.value


class AdvancerState:
    """Flat image of an `Advancer.A.body array` (advancer.ml:27-45); A.make leaves every
    field after p as NaN (advancer.ml:98-116)."""

    def __init__(self, m, q, p, t=0.0):
        n = len(m)
        self.id = np.arange(n, dtype=np.int32)
        self.state = np.full((n, ADV_STRIDE), np.nan)
        self.state[:, 0] = t
        self.state[:, 1] = m
        self.state[:, 2:5] = q
        self.state[:, 5:8] = p
        self.stats = (0, 0)

    t = property(lambda s: s.state[:, 0])
    m = property(lambda s: s.state[:, 1])
    q = property(lambda s: np.ascontiguousarray(s.state[:, 2:5]))
    p = property(lambda s: np.ascontiguousarray(s.state[:, 5:8]))
    h = property(lambda s: s.state[:, 8])
    hmax = property(lambda s: s.state[:, 9])


def advancer_advance(s: AdvancerState, dt, sf, eps2=0.0, extpot=None) -> AdvancerState:
    """extpot: None, the string "plummer_test" (test/advance_test.ml:39-46 field), or a Python
    callable (m, q[3], p[3]) -> gv[3]."""
    import copy
    o = copy.deepcopy(s)
    o.state = np.ascontiguousarray(o.state)
    stats = (C.c_long * 2)(0, 0)
    keep = None
    if extpot is None:
        fn = None
    elif extpot == "plummer_test":
        fn = C.cast(lib().oracle_extpot_plummer_test, C.c_void_p)
    else:
        def tramp(_arg, m, q, p, gv):
            r = extpot(m, np.array([q[0], q[1], q[2]]), np.array([p[0], p[1], p[2]]))
            gv[0], gv[1], gv[2] = r[0], r[1], r[2]
        keep = EXTPOT_FN(tramp)
        fn = C.cast(keep, C.c_void_p)
    lib().oracle_advancer_advance(len(o.id), _ip(o.id), _p(o.state), dt, sf, eps2, fn, None, stats)
    o.stats = (stats[0], stats[1])
    del keep
    return o


def extpot_plummer_test_energy(m, q):
    m, q = _f(m), _f(q)
    return lib().oracle_extpot_plummer_test_energy(len(m), _p(m), _p(q))


# ---- text snapshots: A.print / A.read (advancer.ml:88-96, 118-124), Leapfrog.print / read
# (leapfrog.ml:44-62).  Pure-Python restatement: OCaml's Printf "%g" / "%d" are C's, and so is
# Python's % operator; Scanf's " key: %g " (blanks match any white space) is a regular expression.
def _particle_text(i, t, q, v, m):
    return ("--- !Particle\n" + "id: %d\n" % i + "t: %g\n" % t + "r:\n" +
            "  - %g\n  - %g\n  - %g\n" % (q[0], q[1], q[2]) + "v:\n" +
            "  - %g\n  - %g\n  - %g\n" % (v[0], v[1], v[2]) + "m: %g\n" % m)


def advancer_print_text(ids, t, m, q, p):
    """A.print for every body (advancer.ml:88-96): v = p / m."""
    return "".join(_particle_text(int(ids[i]), t[i], q[i], (p[i][0] / m[i], p[i][1] / m[i],
                                                             p[i][2] / m[i]), m[i])
                   for i in range(len(m)))


def leapfrog_print_text(ids, t, dt_max, m, q, v):
    """Leapfrog.print for every body (leapfrog.ml:44-53): t_max = t + dt_max."""
    return "".join(_particle_text(int(ids[i]), t[i], q[i], v[i], m[i]) +
                   "t_max: %g\n" % (t[i] + dt_max[i]) for i in range(len(m)))


_FLT = r"([-+]?(?:nan|inf(?:inity)?|(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?))"
_REC = (r"\s*---\s*!Particle\s*id:\s*([-+]?\d+)\s*t:\s*" + _FLT + r"\s*r:\s*-\s*" + _FLT +
        r"\s*-\s*" + _FLT + r"\s*-\s*" + _FLT + r"\s*v:\s*-\s*" + _FLT + r"\s*-\s*" + _FLT +
        r"\s*-\s*" + _FLT + r"\s*m:\s*" + _FLT + r"\s*")


def advancer_read_text(text):
    """A.read repeated to end of input (advancer.ml:118-124): p = v * m."""
    import re
    ids, t, m, q, p = [], [], [], [], []
    pos = 0
    pat = re.compile(_REC, re.I)
    while text[pos:].strip():
        mt = pat.match(text, pos)
        if not mt:
            raise ValueError(f"Scan_failure at byte {pos}")
        g = mt.groups()
        ids.append(int(g[0])); t.append(float(g[1])); m.append(float(g[8]))
        q.append([float(x) for x in g[2:5]])
        p.append([float(x) * float(g[8]) for x in g[5:8]])
        pos = mt.end()
    return (np.array(ids, dtype=np.int32), np.array(t), np.array(m), np.array(q).reshape(-1, 3),
            np.array(p).reshape(-1, 3))


def leapfrog_read_text(text):
    """Leapfrog.read repeated to end of input (leapfrog.ml:55-62): dt_max = t_max - t."""
    import re
    ids, t, dtm, m, q, v = [], [], [], [], [], []
    pos = 0
    pat = re.compile(_REC + r"t_max:\s*" + _FLT + r"\s*", re.I)
    while text[pos:].strip():
        mt = pat.match(text, pos)
        if not mt:
            raise ValueError(f"Scan_failure at byte {pos}")
        g = mt.groups()
        ids.append(int(g[0])); t.append(float(g[1])); m.append(float(g[8]))
        q.append([float(x) for x in g[2:5]])
        v.append([float(x) for x in g[5:8]])
        dtm.append(float(g[9]) - float(g[1]))
        pos = mt.end()
    return (np.array(ids, dtype=np.int32), np.array(t), np.array(dtm), np.array(m),
            np.array(q).reshape(-1, 3), np.array(v).reshape(-1, 3))
