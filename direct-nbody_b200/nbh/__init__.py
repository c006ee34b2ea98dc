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
_pi = C.POINTER(C.c_int)
_pl = C.POINTER(C.c_long)

SYMBOLS = {
    "nbh_last_error": (C.c_char_p, []),
    "nbh_make_plummer": (_i, [_i, _u64, _pd, _pd, _pd]),
    "nbh_make_hot_spherical": (_i, [_i, _u64, _pd, _pd, _pd]),
    "nbh_make_cold_spherical": (_i, [_i, _u64, _pd, _pd, _pd]),
    "nbh_make_uniform_box": (_i, [_i, _u64, _pd, _pd, _pd]),
    "nbh_adjust_frame": (_i, [_i, _pd, _pd, _pd]),
    "nbh_add_mass_disc": (_i, [_i, _u64, _d, _d, _pd, _pd]),
    "nbh_rescale_to_standard_units": (_i, [_vp, _i, _pd, _pd, _pd, _d]),
    "nbh_energy": (_i, [_vp, _i, _pd, _pd, _pd, _d, _pd, _pd]),
    "nbh_bodies_to_el_phase_space": (_i, [_vp, _i, _pd, _pd, _pd, _d, _pd]),
    "nbh_omelyan_advance": (_i, [_vp, _i, _pd, _pd, _pd, _pd, _d, _d]),
    "nbh_omelyan_advance_steps": (_i, [_vp, _i, _pd, _pd, _pd, _pd, _d, _d, _i, _pd, _pd]),
    "nbh_leapfrog_advance": (_i, [_vp, _i, _pi, _pd, _pd, _pd, _pd, _pd, _d, _d, _pl]),
    "nbh_leapfrog_advance_const_dt": (_i, [_vp, _i, _pi, _pd, _pd, _pd, _pd, _pd, _d, _i]),
    "nbh_hermite_advance": (_i, [_vp, _i, _pi, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _d, _d, _d, _pl]),
    "nbh_dfloat_hermite_advance": (_i, [_vp, _i, _i, _pi, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _pd, _pd,
                                        _pd, _pd, _pd, _pd, _d, _d, _d, _pl]),
    "nbh_dfloat_jacobian_advance": (_i, [_vp, _i, _pd, _pd, _pd, _d, _d, _d, _pd, _pd, _pl]),
    "nbh_omega_pushforward": (_d, [_i, _pd]),
    "nbh_advancer_advance": (_i, [_vp, _i, _pi, _pd, _d, _d, _d, _vp, _vp, _pl]),
    "nbh_extpot_plummer_test_energy": (_d, [_i, _pd, _pd]),
    "nbh_constant_sf_integrate": (_i, [_vp, _i, _pi, _pd, _d, _d, _d, _d, _d, _pd, _i, _pi]),
    "nbh_advancer_print": (_i, [C.c_char_p, _i, _i, _pi, _pd, _pd, _pd, _pd]),
    "nbh_advancer_read": (_i, [C.c_char_p, _i, _pi, _pi, _pd, _pd, _pd, _pd]),
    "nbh_leapfrog_print": (_i, [C.c_char_p, _i, _i, _pi, _pd, _pd, _pd, _pd, _pd]),
    "nbh_marshal_count_states": (_i, [C.c_char_p, _pi]),
    "nbh_marshal_read_state": (_i, [C.c_char_p, _i, _i, _pi, _pi, _pi, _pd]),
    "nbh_marshal_write_state": (_i, [C.c_char_p, _i, _i, _pi, _pd]),
    "nbh_leapfrog_read": (_i, [C.c_char_p, _i, _pi, _pi, _pd, _pd, _pd, _pd, _pd]),
}
ADV_STRIDE = 35


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
    if a is None:
        return None
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


def _ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_pi)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ------------------------------------------------------------------ Energy / Analysis
def energy(ctx, m, q, p, eps2=0.0):
    """Energy.Make(B).energy split as (ke, pe) (energy.ml:57-78)."""
    m, q, p = _f(m), _f(q), _f(p)
    ke, pe = _d(), _d()
    _ck(load_library().nbh_energy(ctx._h, len(m), _p(m), _p(q), _p(p), eps2, C.byref(ke),
                                  C.byref(pe)))
    return ke.value, pe.value


def bodies_to_el_phase_space(ctx, m, q, p, eps2=0.0):
    m, q, p = _f(m), _f(q), _f(p)
    el = np.empty((len(m), 4))
    _ck(load_library().nbh_bodies_to_el_phase_space(ctx._h, len(m), _p(m), _p(q), _p(p), eps2,
                                                    _p(el)))
    return el


# ------------------------------------------------------------------ Omelyan_advancer.OA
def omelyan_advance(ctx, m, q, p, t, h, eps2=0.0):
    """OA.advance h bs (omelyan_advancer.ml:96-104): functional, returns (q, p, t)."""
    m = _f(m)
    q, p = _f(q).copy(), _f(p).copy()
    tt = _d(t)
    _ck(load_library().nbh_omelyan_advance(ctx._h, len(m), _p(m), _p(q), _p(p), C.byref(tt), h,
                                           eps2))
    return q, p, tt.value


def omelyan_advance_steps(ctx, m, q, p, t, h, nsteps, eps2=0.0, want_energies=True, want_el=False):
    """nsteps of OA.advance h with the bodies resident on the device; returns
    (q, p, t, energies[nsteps] | None, el[n,4] | None)."""
    m = _f(m)
    q, p = _f(q).copy(), _f(p).copy()
    tt = _d(t)
    en = np.zeros(nsteps) if want_energies else None
    el = np.zeros((len(m), 4)) if want_el else None
    _ck(load_library().nbh_omelyan_advance_steps(ctx._h, len(m), _p(m), _p(q), _p(p), C.byref(tt),
                                                 h, eps2, nsteps, _p(en), _p(el)))
    return q, p, tt.value, en, el


# ------------------------------------------------------------------ Leapfrog
class LeapfrogState:
    """Flat image of a Leapfrog.b array; Leapfrog.make sets t = 0, dt_max = 0, v = p / m
    (leapfrog.ml:28-34)."""

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

    def copy(self):
        import copy
        return copy.deepcopy(self)


def leapfrog_advance(ctx, s, dt, eta, resident=2, tree_max=0):
    """Leapfrog.advance bs dt eta; resident: 2 (default) bodies on the device and every subtree
    of advance_subsystem that fits in one launch (nbk_lf_advance_subsystem; tree_max bodies,
    0 = automatic); 1/True bodies on the device (nbk_lf_*), block recursion on the host; 0/False
    one nbk_kick_block_sum (upload + read-back) per kick block (1 and 0 are bit-identical)."""
    load_library().nbh_set_leapfrog_resident(int(resident))
    load_library().nbh_set_leapfrog_tree_max(int(tree_max))
    o = s.copy()
    nk = C.c_long(0)
    _ck(load_library().nbh_leapfrog_advance(ctx._h, len(o.m), _ip(o.id), _p(o.t), _p(o.dt_max),
                                            _p(o.m), _p(o.q), _p(o.v), dt, eta, C.byref(nk)))
    o.nkicks = nk.value
    return o


def leapfrog_advance_const_dt(ctx, s, dt, nsteps, resident=True):
    load_library().nbh_set_leapfrog_resident(int(resident))
    o = s.copy()
    _ck(load_library().nbh_leapfrog_advance_const_dt(ctx._h, len(o.m), _ip(o.id), _p(o.t),
                                                     _p(o.dt_max), _p(o.m), _p(o.q), _p(o.v), dt,
                                                     nsteps))
    return o


# ------------------------------------------------------------------ Hermite
class HermiteState:
    """Flat image of a Hermite.b array (hermite.ml:20-29); Hermite.make sets h = -1."""

    def __init__(self, m, q, p, t=0.0):
        n = len(m)
        self.id = np.arange(1, n + 1, dtype=np.int32)
        self.t = np.full(n, float(t))
        self.m, self.q, self.p = _f(m).copy(), _f(q).copy(), _f(p).copy()
        self.h = np.full(n, -1.0)
        self.f, self.df = np.zeros((n, 3)), np.zeros((n, 3))
        self.nsteps = 0

    def copy(self):
        import copy
        return copy.deepcopy(self)


def hermite_advance(ctx, s, dt, sf, eps2=0.0, mode=0):
    """Hermite.advance bs dt sf.  mode 0: automatic; 1: host scheduler, one launch per step;
    2: the whole stepping loop on the device (n <= nbk_hermite_resident_max())."""
    load_library().nbh_set_hermite_mode(mode)
    o = s.copy()
    ns = C.c_long(o.nsteps)
    _ck(load_library().nbh_hermite_advance(ctx._h, len(o.m), _ip(o.id), _p(o.t), _p(o.m), _p(o.q),
                                           _p(o.p), _p(o.h), _p(o.f), _p(o.df), dt, sf, eps2,
                                           C.byref(ns)))
    o.nsteps = ns.value
    return o


# ------------------------------------------------------------------ Advancer.A / Integrator
class AdvancerState:
    """Flat image of an Advancer.A.body array (advancer.ml:27-45); A.make leaves everything
    after p as NaN (advancer.ml:98-116)."""

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
    hmax = property(lambda s: s.state[:, 9])

    def copy(self):
        import copy
        o = copy.deepcopy(self)
        o.state = np.ascontiguousarray(o.state)
        return o


def advancer_advance(ctx, s, dt, sf, eps2=0.0, extpot=None, resident=2, tree_max=0):
    """Advancer.A.advance ?extpot bs dt sf; extpot is None or "plummer_test"
    (the field of test/advance_test.ml:39-46).  resident: 2 (default) records on the device and
    the block recursion of every subtree that fits on the device too (nbk_adv_advance_range;
    tree_max rows, 0 = automatic); 1/True records on the device, recursion on the host; 0/False
    ship the bodies per interact_bodies (1 and 0 are bit-identical)."""
    load_library().nbh_set_advancer_resident(int(resident))
    load_library().nbh_set_advancer_tree_max(int(tree_max))
    o = s.copy()
    stats = (C.c_long * 2)(0, 0)
    fn = None
    if extpot == "plummer_test":
        fn = C.cast(load_library().nbh_extpot_plummer_test, _vp)
    elif extpot is not None:
        raise ValueError("extpot: None or 'plummer_test'")
    _ck(load_library().nbh_advancer_advance(ctx._h, len(o.id), _ip(o.id), _p(o.state), dt, sf,
                                            eps2, fn, None, stats))
    o.stats = (stats[0], stats[1])
    return o


def extpot_plummer_test_energy(m, q):
    m, q = _f(m), _f(q)
    return load_library().nbh_extpot_plummer_test_energy(len(m), _p(m), _p(q))


def dfloat_jacobian_advance(ctx, m, q, p, dt, sf, eps2=0.0):
    """DFloat_pushforward.Make(DFloat_hermite).jacobian_advance dt sf bs
    (dFloat_pushforward.ml:61-76) -> (jac [6n][6n], advanced [q; p] state [6n], advance_body calls)."""
    m, q, p = (np.ascontiguousarray(a, dtype=np.float64) for a in (m, q, p))
    n = len(m)
    jac, out = np.zeros((6 * n, 6 * n)), np.zeros(6 * n)
    ns = C.c_long(0)
    _ck(load_library().nbh_dfloat_jacobian_advance(ctx._h, n, _p(m), _p(q), _p(p), dt, sf, eps2,
                                                   _p(jac), _p(out), C.byref(ns)))
    return jac, out, ns.value


def omega_pushforward(jac):
    jac = np.ascontiguousarray(jac, dtype=np.float64)
    return load_library().nbh_omega_pushforward(jac.shape[0] // 6, _p(jac))


def dfloat_hermite_advance(ctx, m, q, p, dq, dp, dt, sf, eps2=0.0):
    """DFloat_hermite.advance on fresh bodies (t = 0, h = -1) carrying the K tangent directions
    (dq, dp) [K][n][3] -> dict of advanced values and tangents (id order)."""
    m, q, p = (np.array(a, dtype=np.float64, order="C") for a in (m, q, p))
    n = len(m)
    q_tan = np.array(dq, dtype=np.float64, order="C").reshape(-1, n, 3)
    p_tan = np.array(dp, dtype=np.float64, order="C").reshape(-1, n, 3)
    K = q_tan.shape[0]
    ids = np.arange(1, n + 1, dtype=np.int32)
    t, h = np.zeros(n), np.full(n, -1.0)
    f, df = np.zeros((n, 3)), np.zeros((n, 3))
    t_tan, h_tan = np.zeros((K, n)), np.zeros((K, n))
    f_tan, df_tan = np.zeros((K, n, 3)), np.zeros((K, n, 3))
    ns = C.c_long(0)
    _ck(load_library().nbh_dfloat_hermite_advance(
        ctx._h, n, K, _ip(ids), _p(t), _p(m), _p(q), _p(p), _p(h), _p(f), _p(df), _p(t_tan),
        _p(q_tan), _p(p_tan), _p(h_tan), _p(f_tan), _p(df_tan), dt, sf, eps2, C.byref(ns)))
    return dict(t=t, q=q, p=p, h=h, f=f, df=df, t_tan=t_tan, q_tan=q_tan, p_tan=p_tan,
                h_tan=h_tan, f_tan=f_tan, df_tan=df_tan, nsteps=ns.value)


class EgErr(RuntimeError):
    """Integrator.Eg_err (integrator.ml:33): carries the last good state."""

    def __init__(self, state, errors):
        super().__init__("energy error bound exceeded")
        self.state, self.errors = state, errors


def constant_sf_integrate(ctx, s, dt, ci=1.0, sf=2e-3, max_eg_err=1e-2, eps2=0.0, resident=2):
    """Integrator.Make(B)(Advancer.A).constant_sf_integrate (integrator.ml:58-84).  Returns
    (final state, energy errors after each step); raises EgErr like the reference.
    resident: as in advancer_advance; with 1 or 2 the checkpoint energy is taken from the
    resident records."""
    load_library().nbh_set_advancer_resident(int(resident))
    load_library().nbh_set_advancer_tree_max(0)
    o = s.copy()
    cap = 4096
    errs = np.zeros(cap)
    nerr = C.c_int(0)
    rc = load_library().nbh_constant_sf_integrate(ctx._h, len(o.id), _ip(o.id), _p(o.state), dt,
                                                  ci, sf, max_eg_err, eps2, _p(errs), cap,
                                                  C.byref(nerr))
    e = errs[:min(nerr.value, cap)].copy()
    if rc == -11:
        raise EgErr(o, e)
    _ck(rc)
    return o, e


# ---- text snapshots (A.print / A.read, Leapfrog.print / read) ----------------------------------
def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def advancer_print(path, ids, t, m, q, p, append=False):
    ids, t, m, q, p = _i32(ids), _f(t), _f(m), _f(q), _f(p)
    _ck(load_library().nbh_advancer_print(os.fsencode(path), int(append), len(m), _ip(ids), _p(t),
                                          _p(m), _p(q), _p(p)))


def advancer_read(path, cap=1 << 20):
    n = C.c_int(0)
    ids = np.zeros(cap, dtype=np.int32)
    t, m = np.zeros(cap), np.zeros(cap)
    q, p = np.zeros((cap, 3)), np.zeros((cap, 3))
    _ck(load_library().nbh_advancer_read(os.fsencode(path), cap, C.byref(n), _ip(ids), _p(t), _p(m),
                                         _p(q), _p(p)))
    k = n.value
    return ids[:k], t[:k], m[:k], q[:k], p[:k]


def leapfrog_print(path, ids, t, dt_max, m, q, v, append=False):
    ids, t, dt_max, m, q, v = _i32(ids), _f(t), _f(dt_max), _f(m), _f(q), _f(v)
    _ck(load_library().nbh_leapfrog_print(os.fsencode(path), int(append), len(m), _ip(ids), _p(t),
                                          _p(dt_max), _p(m), _p(q), _p(v)))


def leapfrog_read(path, cap=1 << 20):
    n = C.c_int(0)
    ids = np.zeros(cap, dtype=np.int32)
    t, dtm, m = np.zeros(cap), np.zeros(cap), np.zeros(cap)
    q, v = np.zeros((cap, 3)), np.zeros((cap, 3))
    _ck(load_library().nbh_leapfrog_read(os.fsencode(path), cap, C.byref(n), _ip(ids), _p(t),
                                         _p(dtm), _p(m), _p(q), _p(v)))
    k = n.value
    return ids[:k], t[:k], dtm[:k], m[:k], q[:k], v[:k]


# ---- binary state streams (Marshal.to_channel / Analysis.load_states, analysis.ml:556-575) -------
def marshal_count_states(path):
    n = C.c_int(0)
    _ck(load_library().nbh_marshal_count_states(os.fsencode(path), C.byref(n)))
    return n.value


def marshal_read_state(path, index, cap=1 << 20):
    """Analysis.load_state index file -> (kind, ids, state[n][35]); kind 17 = Advancer.A.body,
    6 = Leapfrog.b ({t, m, q, v, dt_max} in the first 9 columns)."""
    n, kind = C.c_int(0), C.c_int(0)
    ids = np.zeros(cap, dtype=np.int32)
    state = np.zeros((cap, ADV_STRIDE))
    _ck(load_library().nbh_marshal_read_state(os.fsencode(path), index, cap, C.byref(n),
                                              C.byref(kind), _ip(ids), _p(state)))
    return kind.value, ids[:n.value].copy(), state[:n.value].copy()


def marshal_load_states(path, cap=1 << 20):
    """Analysis.load_states file: every state of the stream, in order."""
    return [marshal_read_state(path, k, cap) for k in range(marshal_count_states(path))]


def marshal_write_state(path, ids, state, append=True):
    """Marshal.to_channel out (bs : Advancer.A.body array) [] (bin/el_mass_segregation.ml:104-108)."""
    ids = _i32(ids)
    state = np.ascontiguousarray(state, dtype=np.float64)
    _ck(load_library().nbh_marshal_write_state(os.fsencode(path), int(append), len(ids), _ip(ids),
                                               _p(state)))
