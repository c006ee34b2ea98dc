"""nbk — Python (ctypes) binding of libnbk's C ABI (include/nbk.h).

This is the harness-side mirror of what the OCaml `external` stubs of INTEGRATION.md bind: the
same entry points, numpy arrays standing in for Bigarray.Array1 (float64, c_layout).  It is
plumbing only — every sweep runs in the CUDA library; there is NO CPU fallback: if
libnbk.so is missing or no B200 is visible, loading / `Context()` raises.

Error convention follows the reference's (OCaml exceptions): a non-zero
return code raises `NbkError` (the stub's `Failure (nbk_last_error ctx)`).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# NBK_LIBRARY: another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("NBK_LIBRARY") or os.path.join(os.path.dirname(_PKG), "lib", "libnbk.so")
PACKED_DOUBLES = 8

_d, _i, _f = C.c_double, C.c_int, C.c_float
_pd = C.POINTER(C.c_double)
_vp = C.c_void_p

MAX_PEERS = 8
IPC_HANDLE_BYTES = 64


class Shards(C.Structure):
    """nbk_shards (include/nbk.h): where each rank's packed rows live, as seen from this GPU."""
    _fields_ = [("nranks", _i), ("self", _i), ("shard_rows", _i),
                ("packed", _vp * MAX_PEERS), ("aux", _vp * MAX_PEERS),
                ("ready", _vp), ("epoch", C.c_uint64)]


_psh = C.POINTER(Shards)

# name -> (restype, argtypes): every symbol include/nbk.h declares.
SYMBOLS = {
    "nbk_create": (_i, [C.POINTER(_vp), _i]),
    "nbk_create_multi": (_i, [C.POINTER(_vp), _i, C.POINTER(_i)]),
    "nbk_device_count": (_i, [_vp]),
    "nbk_destroy": (None, [_vp]),
    "nbk_last_error": (C.c_char_p, [_vp]),
    "nbk_version": (C.c_char_p, []),
    "nbk_sm_count": (_i, [_vp]),
    "nbk_launch_count": (C.c_int64, [_vp]),
    "nbk_last_kernel_ms": (_i, [_vp, C.POINTER(_f)]),
    "nbk_grad_v_sum": (_i, [_vp, _i, _pd, _pd, _d, _i, _i, _i, _i, _pd]),
    "nbk_grad_v_dgrad_v_sum": (_i, [_vp, _i, _pd, _pd, _pd, _d, _i, _i, _i, _i, _pd, _pd]),
    "nbk_v_sum": (_i, [_vp, _i, _pd, _pd, _d, _i, _i, _i, _i, _pd, _pd]),
    "nbk_energy": (_i, [_vp, _i, _pd, _pd, _pd, _d, _pd, _pd]),
    "nbk_kick_sum": (_i, [_vp, _i, _pd, _pd, _pd, _d, _d, _i, _i, _i, _i, _pd, _pd]),
    "nbk_binary_scan": (_i, [_vp, _i, _pd, _pd, _pd, _d, _d, _i, _i, _i, _i, _pd, C.POINTER(_i),
                             C.POINTER(_i)]),
    "nbk_binary_list": (_i, [_vp, _i, _pd, _pd, _pd, _d, _d, _i, C.POINTER(_i), C.POINTER(_i), _pd,
                             C.POINTER(_i)]),
    "nbk_interact_sum": (_i, [_vp, _i, _i, _pd, _pd, _d, _pd]),
    "nbk_kick_block_sum": (_i, [_vp, _i, _i, _pd, _pd, _pd, _d, _d, _pd, _pd]),
    "nbk_grad_v_dgrad_v_sum_predicted": (_i, [_vp, _i, _pd, _pd, _pd, _pd, _pd, _pd, _d, _d,
                                              _i, _i, _i, _i, _pd, _pd]),
    "nbk_grad_v_dgrad_v_sum_tangent": (_i, [_vp, _i, _pd, _pd, _pd, _i, _pd, _pd, _d,
                                            _i, _i, _i, _i, _pd, _pd, _pd, _pd]),
    "nbk_grad_v_dgrad_v_sum_predicted_tangent": (_i, [_vp, _i, _pd, _pd, _pd, _pd, _pd, _pd, _i, _pd,
                                                      _pd, _pd, _pd, _pd, _d, _pd, _d, _i, _i, _i,
                                                      _i, _pd, _pd, _pd, _pd]),
    "nbk_grad_v_sum_tangent": (_i, [_vp, _i, _pd, _pd, _i, _pd, _d, _i, _i, _i, _i, _pd, _pd]),
    "nbk_bodies_upload": (_i, [_vp, _i, _pd, _pd, _pd, _i]),
    "nbk_bodies_update": (_i, [_vp, _i, _i, _pd, _pd, _i]),
    "nbk_resident_count": (_i, [_vp]),
    "nbk_resident_grad_v_sum": (_i, [_vp, _d, _i, _i, _i, _i, _pd]),
    "nbk_resident_grad_v_dgrad_v_sum": (_i, [_vp, _d, _i, _i, _i, _i, _pd, _pd]),
    "nbk_resident_v_sum": (_i, [_vp, _d, _i, _i, _i, _i, _pd, _pd]),
    "nbk_resident_kick_sum": (_i, [_vp, _d, _d, _i, _i, _i, _i, _pd, _pd]),
    "nbk_resident_leapfrog_steps": (_i, [_vp, _d, _i, _pd, _pd]),
    "nbk_adv_upload": (_i, [_vp, _i, _pd]),
    "nbk_adv_download": (_i, [_vp, _i, _i, _pd]),
    "nbk_adv_permute": (_i, [_vp, _i, C.POINTER(_i)]),
    "nbk_adv_begin_step": (_i, [_vp, _i, _i, _d]),
    "nbk_adv_interact": (_i, [_vp, _i, _i, _d, _d, _d]),
    "nbk_adv_finish": (_i, [_vp, _i, _i, _d, _pd]),
    "nbk_adv_get_t": (_i, [_vp, _i, _pd]),
    "nbk_adv_init_first": (_i, [_vp, _d]),
    "nbk_adv_energy": (_i, [_vp, _d, _pd, _pd]),
    "nbk_adv_advance_range": (_i, [_vp, _i, _d, _d, _d, _pd, C.POINTER(_i), C.POINTER(C.c_longlong)]),
    "nbk_adv_range_max": (_i, []),
    "nbk_adv_range_profile": (_i, [_vp, C.POINTER(C.c_longlong), _i]),
    "nbk_lf_advance_subsystem": (_i, [_vp, _i, _d, _d, _pd, _pd, C.POINTER(_i), C.POINTER(C.c_longlong)]),
    "nbk_lf_subsystem_max": (_i, []),
    "nbk_lf_upload": (_i, [_vp, _i, _pd, _pd, _pd, _pd]),
    "nbk_lf_download": (_i, [_vp, _pd, _pd, _pd]),
    "nbk_lf_permute": (_i, [_vp, _i, C.POINTER(_i)]),
    "nbk_lf_sort_sub": (_i, [_vp, _i, C.POINTER(_i)]),
    "nbk_lf_begin": (_i, [_vp, _i, _i]),
    "nbk_lf_drift": (_i, [_vp, _i, _i, _d]),
    "nbk_lf_kick_block": (_i, [_vp, _i, _i, _d, _d, _i, _i, _d, _pd]),
    "nbk_oa_upload": (_i, [_vp, _i, _pd, _pd, _pd]),
    "nbk_oa_steps": (_i, [_vp, _d, _d, _i]),
    "nbk_oa_download": (_i, [_vp, _pd, _pd]),
    "nbk_oa_diagnostics": (_i, [_vp, _d, _pd, _pd, _pd]),
    "nbk_hermite_upload": (_i, [_vp, _i, _pd, _pd, _pd, _pd, _pd, _pd, _pd]),
    "nbk_hermite_resident_max": (_i, []),
    "nbk_hermite_advance_resident": (_i, [_vp, _d, _d, _d, C.c_long, C.POINTER(C.c_long)]),
    "nbk_hermite_download": (_i, [_vp, _pd, _pd, _pd, _pd, _pd, _pd]),
    "nbk_hermite_body_sum": (_i, [_vp, _i, _d, _d, _i, _pd, _pd, _pd]),
    "nbk_dev_pack": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "nbk_dev_grad_v_dgrad_v_sum": (_i, [_vp, _vp, _i, _d, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "nbk_dev_grad_v_sum": (_i, [_vp, _vp, _i, _d, _i, _i, _i, _i, _vp, _i, _vp]),
    "nbk_dev_v_sum": (_i, [_vp, _vp, _i, _d, _i, _i, _i, _i, _vp, _i, _vp]),
    "nbk_dev_kick_sum": (_i, [_vp, _vp, _i, _d, _d, _i, _i, _i, _i, _vp, _vp, _vp]),
    "nbk_dev_kick_sum_peer": (_i, [_vp, _psh, _i, _d, _d, _i, _i, _i, _i, _vp, _vp, _vp]),
    "nbk_dev_pack_aux": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "nbk_dev_predict_pack": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _vp]),
    "nbk_dev_predict_pack_tangent": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _vp, _vp,
                                          _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "nbk_dev_grad_v_dgrad_v_sum_tangent": (_i, [_vp, _vp, _vp, _i, _i, _d, _i, _i, _i, _i, _vp, _vp,
                                                _vp]),
    "nbk_peer_alloc": (_i, [_vp, C.c_size_t, C.POINTER(_vp), C.c_char_p]),
    "nbk_peer_open": (_i, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "nbk_peer_close": (_i, [_vp, _vp]),
    "nbk_peer_free": (_i, [_vp, _vp]),
    "nbk_dev_publish": (_i, [_vp, _i, _i, C.POINTER(_vp), C.c_uint64, _vp]),
    "nbk_dev_grad_v_dgrad_v_sum_peer": (_i, [_vp, _psh, _i, _d, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "nbk_dev_grad_v_sum_peer": (_i, [_vp, _psh, _i, _d, _i, _i, _i, _i, _vp, _i, _vp]),
    "nbk_dev_v_sum_peer": (_i, [_vp, _psh, _i, _d, _i, _i, _i, _i, _vp, _i, _vp]),
    "nbk_dev_grad_v_dgrad_v_sum_tangent_peer": (_i, [_vp, _psh, _i, _i, _d, _i, _i, _i, _i, _vp,
                                                     _vp, _vp]),
    "nbk_set_peer_timeout": (_i, [_vp, _d]),
    "nbk_tune": (_i, [_vp, _i]),
    "nbk_set_sym": (_i, [_vp, _i]),
    "nbk_sym_items": (_i, [_i, C.POINTER(C.c_longlong)]),
    "nbk_dev_grad_v_dgrad_v_sym": (_i, [_vp, _vp, _i, _d, C.c_longlong, C.c_longlong, _vp, _vp]),
    "nbk_dev_grad_v_dgrad_v_sym_peer": (_i, [_vp, _psh, _i, _d, C.c_longlong, C.c_longlong, _vp,
                                             _vp]),
    "nbk_dev_split6": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "nbk_dev_sym_gather_peer": (_i, [_vp, _i, _i, C.POINTER(_vp), _vp, C.c_uint64, _i, _i, _vp, _vp,
                                     _vp]),
    "nbk_dev_pack_push": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, C.POINTER(_vp), _vp]),
    "nbk_fp64_peak": (_i, [_vp, _i, _pd, C.POINTER(_f)]),
}


class NbkError(RuntimeError):
    """Raised on any non-zero libnbk return code (the OCaml stubs raise Failure)."""

    def __init__(self, code, msg):
        super().__init__(f"libnbk error {code}: {msg}")
        self.code = code


_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libnbk.so and declare every symbol of include/nbk.h.  Needs no GPU."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise FileNotFoundError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _arr(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        a = a.reshape(shape)
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(_pd)


class Context:
    """One libnbk context = one GPU (nbk_create / nbk_destroy)."""

    def __init__(self, device: int = 0, devices=None):
        """device: one GPU.  devices=[d0, d1, ...]: a single-process multi-GPU context
        (nbk_create_multi) whose host-buffer sweeps shard their targets over those GPUs."""
        self._lib = load_library()
        h = _vp()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            rc = self._lib.nbk_create_multi(C.byref(h), len(devices), arr)
            device = devices[0]
        else:
            rc = self._lib.nbk_create(C.byref(h), device)
        if rc != 0:
            raise NbkError(rc, self._lib.nbk_last_error(None).decode())
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nbk_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise NbkError(rc, self._lib.nbk_last_error(self._h).decode())

    # -------------------------------------------------------------- info
    @property
    def device_count(self):
        return self._lib.nbk_device_count(self._h)

    @property
    def sm_count(self):
        return self._lib.nbk_sm_count(self._h)

    @property
    def launch_count(self):
        return self._lib.nbk_launch_count(self._h)

    def last_kernel_ms(self):
        ms = _f()
        self._ck(self._lib.nbk_last_kernel_ms(self._h, C.byref(ms)))
        return ms.value

    def set_peer_timeout(self, seconds):
        """How long a *_peer sweep waits for a peer's ready flag before trapping; 0 = forever."""
        self._ck(self._lib.nbk_set_peer_timeout(self._h, float(seconds)))

    def tune(self, variant):
        self._ck(self._lib.nbk_tune(self._h, variant))

    def set_sym(self, mode):
        """Pair-once acc+jerk sweep for full squares: 0 never, 1 automatic, 2 whenever possible."""
        self._ck(self._lib.nbk_set_sym(self._h, int(mode)))

    def sym_items(self, n_total):
        """Block pairs (work items) of the pair-once sweep over n_total bodies."""
        k = C.c_longlong()
        self._ck(self._lib.nbk_sym_items(int(n_total), C.byref(k)))
        return k.value

    def dev_grad_v_dgrad_v_sym(self, d_packed, n_total, eps2, items, d_sum6, stream=0):
        self._ck(self._lib.nbk_dev_grad_v_dgrad_v_sym(self._h, d_packed, n_total, eps2,
                                                      int(items[0]), int(items[1]), d_sum6,
                                                      stream or None))

    def dev_grad_v_dgrad_v_sym_peer(self, shards, n_total, eps2, items, d_sum6, stream=0):
        self._ck(self._lib.nbk_dev_grad_v_dgrad_v_sym_peer(
            self._h, C.byref(shards), n_total, eps2, int(items[0]), int(items[1]), d_sum6,
            stream or None))

    def dev_pack_push(self, n, d_m, d_q, d_vel, is_velocity, dst, stream=0):
        """dev_pack with every row stored into len(dst) destinations (peer-mapped addresses)."""
        arr = (_vp * len(dst))(*dst)
        self._ck(self._lib.nbk_dev_pack_push(self._h, n, d_m, d_q, d_vel or None, int(is_velocity),
                                             len(dst), arr, stream or None))

    def dev_sym_gather_peer(self, self_rank, sum6_of_rank, d_ready, epoch, bodies, d_gsum, d_dgsum,
                            stream=0):
        arr = (_vp * len(sum6_of_rank))(*sum6_of_rank)
        self._ck(self._lib.nbk_dev_sym_gather_peer(
            self._h, len(sum6_of_rank), self_rank, arr, d_ready or None, epoch, int(bodies[0]),
            int(bodies[1]), d_gsum, d_dgsum, stream or None))

    def dev_split6(self, d_sum6, count, d_gsum, d_dgsum, stream=0):
        self._ck(self._lib.nbk_dev_split6(self._h, d_sum6, int(count), d_gsum, d_dgsum,
                                          stream or None))

    def fp64_peak(self, iters=20000):
        tf, ms = _d(), _f()
        self._ck(self._lib.nbk_fp64_peak(self._h, iters, C.byref(tf), C.byref(ms)))
        return tf.value, ms.value

    # -------------------------------------------------------------- host-buffer sweeps
    @staticmethod
    def _rng(n, targets, sources):
        t0, t1 = targets if targets is not None else (0, n)
        s0, s1 = sources if sources is not None else (0, n)
        return int(t0), int(t1), int(s0), int(s1)

    def grad_v_sum(self, m, q, eps2=0.0, targets=None, sources=None):
        m, q = _arr(m), _arr(q)
        n = len(m)
        t0, t1, s0, s1 = self._rng(n, targets, sources)
        g = np.empty((max(t1 - t0, 0), 3))
        self._ck(self._lib.nbk_grad_v_sum(self._h, n, _p(m), _p(q), eps2, t0, t1, s0, s1, _p(g)))
        return g

    def grad_v_dgrad_v_sum(self, m, q, p, eps2=0.0, targets=None, sources=None):
        m, q, p = _arr(m), _arr(q), _arr(p)
        n = len(m)
        t0, t1, s0, s1 = self._rng(n, targets, sources)
        g, dg = np.empty((max(t1 - t0, 0), 3)), np.empty((max(t1 - t0, 0), 3))
        self._ck(self._lib.nbk_grad_v_dgrad_v_sum(self._h, n, _p(m), _p(q), _p(p), eps2, t0, t1,
                                                  s0, s1, _p(g), _p(dg)))
        return g, dg

    def v_sum(self, m, q, eps2=0.0, targets=None, sources=None, want_phi=True):
        m, q = _arr(m), _arr(q)
        n = len(m)
        t0, t1, s0, s1 = self._rng(n, targets, sources)
        phi = np.empty(max(t1 - t0, 0)) if want_phi else None
        tot = _d()
        self._ck(self._lib.nbk_v_sum(self._h, n, _p(m), _p(q), eps2, t0, t1, s0, s1, _p(phi),
                                     C.byref(tot)))
        return phi, tot.value

    def energy(self, m, q, p, eps2=0.0):
        """(ke, pe) = Energy.total_kinetic_energy, total_potential_energy (energy.ml:57-73)."""
        m, q, p = _arr(m), _arr(q), _arr(p)
        ke, pe = _d(), _d()
        self._ck(self._lib.nbk_energy(self._h, len(m), _p(m), _p(q), _p(p), eps2, C.byref(ke),
                                      C.byref(pe)))
        return ke.value, pe.value

    def kick_sum(self, m, q, v, eta, dt, targets=None, sources=None, want_tmin=True):
        m, q, v = _arr(m), _arr(q), _arr(v)
        n = len(m)
        t0, t1, s0, s1 = self._rng(n, targets, sources)
        dv = np.empty((max(t1 - t0, 0), 3))
        tmin = np.empty(max(t1 - t0, 0)) if want_tmin else None
        self._ck(self._lib.nbk_kick_sum(self._h, n, _p(m), _p(q), _p(v), eta, dt, t0, t1, s0, s1,
                                        _p(dv), _p(tmin)))
        return dv, tmin

    def binary_scan(self, m, q, p, eps2=0.0, threshold=0.0, targets=None, sources=None):
        m, q, p = _arr(m), _arr(q), _arr(p)
        n = len(m)
        t0, t1, s0, s1 = self._rng(n, targets, sources)
        nt = max(t1 - t0, 0)
        emin = np.empty(nt)
        partner, count = np.empty(nt, dtype=np.int32), np.empty(nt, dtype=np.int32)
        ip = C.POINTER(_i)
        self._ck(self._lib.nbk_binary_scan(self._h, n, _p(m), _p(q), _p(p), eps2, threshold, t0, t1,
                                           s0, s1, _p(emin), partner.ctypes.data_as(ip),
                                           count.ctypes.data_as(ip)))
        return emin, partner, count

    def binary_list(self, m, q, p, eps2=0.0, threshold=0.0, max_pairs=1 << 20):
        m, q, p = _arr(m), _arr(q), _arr(p)
        pi, pj = np.empty(max_pairs, dtype=np.int32), np.empty(max_pairs, dtype=np.int32)
        e = np.empty(max_pairs)
        npairs = _i(0)
        ip = C.POINTER(_i)
        self._ck(self._lib.nbk_binary_list(self._h, len(m), _p(m), _p(q), _p(p), eps2, threshold,
                                           max_pairs, pi.ctypes.data_as(ip), pj.ctypes.data_as(ip),
                                           _p(e), C.byref(npairs)))
        k = npairs.value
        return pi[:k].copy(), pj[:k].copy(), e[:k].copy()

    def interact_sum(self, m, q, na, eps2=0.0):
        m, q = _arr(m), _arr(q)
        S = np.empty((len(m), 3))
        self._ck(self._lib.nbk_interact_sum(self._h, len(m), na, _p(m), _p(q), eps2, _p(S)))
        return S

    def kick_block_sum(self, m, q, v, eta, dt, ibegin, want_tmin=True):
        m, q, v = _arr(m), _arr(q), _arr(v)
        n = len(m)
        dv = np.empty((n, 3))
        tmin = np.empty(n) if want_tmin else None
        self._ck(self._lib.nbk_kick_block_sum(self._h, n, ibegin, _p(m), _p(q), _p(v), eta, dt,
                                              _p(dv), _p(tmin)))
        return dv, tmin

    def grad_v_dgrad_v_sum_predicted(self, t, m, q, p, f, df, nt, eps2=0.0, targets=None,
                                     sources=None):
        t, m, q, p, f, df = (_arr(a) for a in (t, m, q, p, f, df))
        n = len(m)
        t0, t1, s0, s1 = self._rng(n, targets, sources)
        g, dg = np.empty((max(t1 - t0, 0), 3)), np.empty((max(t1 - t0, 0), 3))
        self._ck(self._lib.nbk_grad_v_dgrad_v_sum_predicted(
            self._h, n, _p(t), _p(m), _p(q), _p(p), _p(f), _p(df), nt, eps2, t0, t1, s0, s1,
            _p(g), _p(dg)))
        return g, dg

    def grad_v_dgrad_v_sum_tangent(self, m, q, p, dq, dp, eps2=0.0, targets=None, sources=None):
        m, q, p = _arr(m), _arr(q), _arr(p)
        n = len(m)
        dq, dp = _arr(dq).reshape(-1, n, 3), _arr(dp).reshape(-1, n, 3)
        K = dq.shape[0]
        t0, t1, s0, s1 = self._rng(n, targets, sources)
        nt = max(t1 - t0, 0)
        g, dg = np.empty((nt, 3)), np.empty((nt, 3))
        gt, dgt = np.empty((K, nt, 3)), np.empty((K, nt, 3))
        self._ck(self._lib.nbk_grad_v_dgrad_v_sum_tangent(
            self._h, n, _p(m), _p(q), _p(p), K, _p(dq), _p(dp), eps2, t0, t1, s0, s1, _p(g),
            _p(dg), _p(gt), _p(dgt)))
        return g, dg, gt, dgt

    def grad_v_dgrad_v_sum_predicted_tangent(self, t, m, q, p, f, df, t_tan, q_tan, p_tan, f_tan,
                                             df_tan, nt, nt_tan=None, eps2=0.0, targets=None,
                                             sources=None):
        """DFloat_hermite.advance_body's loop: *_tan are [K][n][3] ([K][n] for t_tan, [K] for
        nt_tan; either of those may be None = zeros)."""
        t, m, q, p, f, df = (_arr(a) for a in (t, m, q, p, f, df))
        n = len(m)
        q_tan, p_tan, f_tan, df_tan = (_arr(a).reshape(-1, n, 3) for a in (q_tan, p_tan, f_tan, df_tan))
        K = q_tan.shape[0]
        t_tan = None if t_tan is None else _arr(t_tan).reshape(K, n)
        nt_tan = None if nt_tan is None else _arr(nt_tan).reshape(K)
        t0, t1, s0, s1 = self._rng(n, targets, sources)
        ntg = max(t1 - t0, 0)
        g, dg = np.empty((ntg, 3)), np.empty((ntg, 3))
        gt, dgt = np.empty((K, ntg, 3)), np.empty((K, ntg, 3))
        self._ck(self._lib.nbk_grad_v_dgrad_v_sum_predicted_tangent(
            self._h, n, _p(t), _p(m), _p(q), _p(p), _p(f), _p(df), K, _p(t_tan), _p(q_tan),
            _p(p_tan), _p(f_tan), _p(df_tan), float(nt), _p(nt_tan), eps2, t0, t1, s0, s1, _p(g),
            _p(dg), _p(gt), _p(dgt)))
        return g, dg, gt, dgt

    def grad_v_sum_tangent(self, m, q, dq, eps2=0.0, targets=None, sources=None):
        m, q = _arr(m), _arr(q)
        n = len(m)
        dq = _arr(dq).reshape(-1, n, 3)
        K = dq.shape[0]
        t0, t1, s0, s1 = self._rng(n, targets, sources)
        nt = max(t1 - t0, 0)
        g, gt = np.empty((nt, 3)), np.empty((K, nt, 3))
        self._ck(self._lib.nbk_grad_v_sum_tangent(self._h, n, _p(m), _p(q), K, _p(dq), eps2, t0,
                                                  t1, s0, s1, _p(g), _p(gt)))
        return g, gt

    # -------------------------------------------------------------- resident bodies
    def bodies_upload(self, m, q, vel, is_velocity=False):
        m, q = _arr(m), _arr(q)
        vel = None if vel is None else _arr(vel)
        self._ck(self._lib.nbk_bodies_upload(self._h, len(m), _p(m), _p(q), _p(vel),
                                             int(is_velocity)))

    def bodies_update(self, i0, i1, q, vel, is_velocity=False):
        q = _arr(q)
        vel = None if vel is None else _arr(vel)
        self._ck(self._lib.nbk_bodies_update(self._h, i0, i1, _p(q), _p(vel), int(is_velocity)))

    @property
    def resident_count(self):
        return self._lib.nbk_resident_count(self._h)

    def resident_grad_v_sum(self, eps2=0.0, targets=None, sources=None):
        t0, t1, s0, s1 = self._rng(self.resident_count, targets, sources)
        g = np.empty((max(t1 - t0, 0), 3))
        self._ck(self._lib.nbk_resident_grad_v_sum(self._h, eps2, t0, t1, s0, s1, _p(g)))
        return g

    def resident_grad_v_dgrad_v_sum(self, eps2=0.0, targets=None, sources=None):
        t0, t1, s0, s1 = self._rng(self.resident_count, targets, sources)
        g, dg = np.empty((max(t1 - t0, 0), 3)), np.empty((max(t1 - t0, 0), 3))
        self._ck(self._lib.nbk_resident_grad_v_dgrad_v_sum(self._h, eps2, t0, t1, s0, s1, _p(g),
                                                           _p(dg)))
        return g, dg

    def resident_v_sum(self, eps2=0.0, targets=None, sources=None, want_phi=True):
        t0, t1, s0, s1 = self._rng(self.resident_count, targets, sources)
        phi = np.empty(max(t1 - t0, 0)) if want_phi else None
        tot = _d()
        self._ck(self._lib.nbk_resident_v_sum(self._h, eps2, t0, t1, s0, s1, _p(phi),
                                              C.byref(tot)))
        return phi, tot.value

    def resident_kick_sum(self, eta, dt, targets=None, sources=None, want_tmin=True):
        t0, t1, s0, s1 = self._rng(self.resident_count, targets, sources)
        dv = np.empty((max(t1 - t0, 0), 3))
        tmin = np.empty(max(t1 - t0, 0)) if want_tmin else None
        self._ck(self._lib.nbk_resident_kick_sum(self._h, eta, dt, t0, t1, s0, s1, _p(dv),
                                                 _p(tmin)))
        return dv, tmin

    def resident_leapfrog_steps(self, dt, nsteps):
        n = self.resident_count
        q, v = np.empty((n, 3)), np.empty((n, 3))
        self._ck(self._lib.nbk_resident_leapfrog_steps(self._h, dt, nsteps, _p(q), _p(v)))
        return q, v

    # -------------------------------------------------------------- resident Hermite state
    # -------------------------------------------------------------- resident Leapfrog.b array
    def lf_upload(self, dt_max, m, q, v):
        dt_max, m, q, v = _arr(dt_max), _arr(m), _arr(q), _arr(v)
        self._ck(self._lib.nbk_lf_upload(self._h, len(m), _p(dt_max), _p(m), _p(q), _p(v)))

    def lf_download(self, n):
        dt_max, q, v = np.empty(n), np.empty((n, 3)), np.empty((n, 3))
        self._ck(self._lib.nbk_lf_download(self._h, _p(dt_max), _p(q), _p(v)))
        return dt_max, q, v

    def lf_sort_sub(self, iend):
        """sort_sub bs iend (leapfrog.ml:106-111) on the device -> perm (new row s = old row perm[s])."""
        perm = np.empty(max(iend, 1), dtype=np.int32)
        self._ck(self._lib.nbk_lf_sort_sub(self._h, iend, perm.ctypes.data_as(C.POINTER(_i))))
        return perm[:iend]

    def hermite_upload(self, t, m, q, p, f, df, h=None):
        t, m, q, p, f, df = (_arr(a) for a in (t, m, q, p, f, df))
        h = None if h is None else _arr(h)
        self._ck(self._lib.nbk_hermite_upload(self._h, len(m), _p(t), _p(m), _p(q), _p(p), _p(h),
                                              _p(f), _p(df)))

    def hermite_advance_resident(self, stop_time, eta, eps2=0.0, max_steps=1 << 40):
        ns = C.c_long(0)
        self._ck(self._lib.nbk_hermite_advance_resident(self._h, stop_time, eta, eps2, max_steps,
                                                        C.byref(ns)))
        return ns.value

    def hermite_download(self, n):
        t, h = np.empty(n), np.empty(n)
        q, p, f, df = (np.empty((n, 3)) for _ in range(4))
        self._ck(self._lib.nbk_hermite_download(self._h, _p(t), _p(q), _p(p), _p(h), _p(f), _p(df)))
        return t, q, p, h, f, df

    def hermite_body_sum(self, i, nt, eps2=0.0, upd=-1, upd_state=None):
        g, dg = np.empty(3), np.empty(3)
        us = None if upd_state is None else _arr(upd_state)
        self._ck(self._lib.nbk_hermite_body_sum(self._h, i, nt, eps2, upd, _p(us), _p(g), _p(dg)))
        return g, dg

    # -------------------------------------------------------------- raw device pointers
    # Arguments are integer device addresses (e.g. torch.Tensor.data_ptr()) and an integer
    # cudaStream_t (torch.cuda.current_stream().cuda_stream); 0 = the legacy default stream.
    def dev_pack(self, n, d_m, d_q, d_vel, is_velocity, d_packed, stream=0):
        self._ck(self._lib.nbk_dev_pack(self._h, n, d_m, d_q, d_vel or None, int(is_velocity),
                                        d_packed, stream or None))

    def dev_grad_v_dgrad_v_sum(self, d_packed, n_total, eps2, targets, sources, d_gsum, d_dgsum,
                               accumulate=False, stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_grad_v_dgrad_v_sum(self._h, d_packed, n_total, eps2, t0, t1,
                                                      s0, s1, d_gsum, d_dgsum, int(accumulate),
                                                      stream or None))

    def dev_pack_aux(self, n, d_m, d_dq, d_dp, d_aux, stream=0):
        self._ck(self._lib.nbk_dev_pack_aux(self._h, n, d_m, d_dq, d_dp or None, d_aux,
                                            stream or None))

    def dev_predict_pack(self, n, d_t, d_m, d_q, d_p, d_f, d_df, nt, d_packed, stream=0):
        self._ck(self._lib.nbk_dev_predict_pack(self._h, n, d_t, d_m, d_q, d_p, d_f, d_df,
                                                float(nt), d_packed, stream or None))

    def dev_predict_pack_tangent(self, n, K, d_t, d_m, d_q, d_p, d_f, d_df, nt, d_t_tan, d_q_tan,
                                 d_p_tan, d_f_tan, d_df_tan, d_nt_tan, d_aux, aux_stride, stream=0):
        self._ck(self._lib.nbk_dev_predict_pack_tangent(
            self._h, n, K, d_t, d_m, d_q, d_p, d_f, d_df, float(nt), d_t_tan or None, d_q_tan,
            d_p_tan, d_f_tan, d_df_tan, d_nt_tan or None, d_aux, aux_stride, stream or None))

    def dev_grad_v_dgrad_v_sum_tangent(self, d_packed, d_aux, n_total, K, eps2, targets, sources,
                                       d_gsum_tan, d_dgsum_tan, stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_grad_v_dgrad_v_sum_tangent(
            self._h, d_packed, d_aux, n_total, K, eps2, t0, t1, s0, s1, d_gsum_tan, d_dgsum_tan,
            stream or None))

    def dev_grad_v_sum(self, d_packed, n_total, eps2, targets, sources, d_gsum, accumulate=False,
                       stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_grad_v_sum(self._h, d_packed, n_total, eps2, t0, t1, s0, s1,
                                              d_gsum, int(accumulate), stream or None))

    def dev_v_sum(self, d_packed, n_total, eps2, targets, sources, d_phi, accumulate=False,
                  stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_v_sum(self._h, d_packed, n_total, eps2, t0, t1, s0, s1, d_phi,
                                         int(accumulate), stream or None))

    # -------------------------------------------------------------- resident Omelyan OA
    def oa_upload(self, m, q, p):
        m, q, p = _arr(m), _arr(q), _arr(p)
        self._oa_n = len(m)
        self._ck(self._lib.nbk_oa_upload(self._h, len(m), _p(m), _p(q), _p(p)))

    def oa_steps(self, h, eps2=0.0, nsteps=1):
        self._ck(self._lib.nbk_oa_steps(self._h, h, eps2, nsteps))

    def oa_download(self):
        q, p = np.empty((self._oa_n, 3)), np.empty((self._oa_n, 3))
        self._ck(self._lib.nbk_oa_download(self._h, _p(q), _p(p)))
        return q, p

    def oa_diagnostics(self, eps2=0.0, want_el=True):
        ke, pe = _d(), _d()
        el = np.empty((self._oa_n, 4)) if want_el else None
        self._ck(self._lib.nbk_oa_diagnostics(self._h, eps2, C.byref(ke), C.byref(pe), _p(el)))
        return ke.value, pe.value, el

    # -------------------------------------------------------------- peer-memory exchange
    def peer_alloc(self, nbytes, want_handle=True):
        """-> (device address, 64-byte IPC handle or None) of a zeroed region."""
        ptr = _vp()
        buf = C.create_string_buffer(IPC_HANDLE_BYTES) if want_handle else None
        self._ck(self._lib.nbk_peer_alloc(self._h, nbytes, C.byref(ptr), buf))
        return ptr.value, (buf.raw if want_handle else None)

    def peer_open(self, handle: bytes):
        ptr = _vp()
        self._ck(self._lib.nbk_peer_open(self._h, handle, C.byref(ptr)))
        return ptr.value

    def peer_close(self, ptr):
        self._ck(self._lib.nbk_peer_close(self._h, ptr))

    def peer_free(self, ptr):
        self._ck(self._lib.nbk_peer_free(self._h, ptr))

    def dev_publish(self, self_rank, ready_of_rank, epoch, stream=0):
        arr = (_vp * len(ready_of_rank))(*ready_of_rank)
        self._ck(self._lib.nbk_dev_publish(self._h, len(ready_of_rank), self_rank, arr, epoch,
                                           stream or None))

    @staticmethod
    def make_shards(self_rank, shard_rows, packed, ready=0, epoch=0, aux=None):
        sh = Shards()
        sh.nranks, sh.self, sh.shard_rows = len(packed), self_rank, shard_rows
        for r, ptr in enumerate(packed):
            sh.packed[r] = ptr
            sh.aux[r] = aux[r] if aux is not None else None
        sh.ready = ready or None
        sh.epoch = epoch
        return sh

    def dev_grad_v_dgrad_v_sum_peer(self, shards, n_total, eps2, targets, sources, d_gsum, d_dgsum,
                                    accumulate=False, stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_grad_v_dgrad_v_sum_peer(
            self._h, C.byref(shards), n_total, eps2, t0, t1, s0, s1, d_gsum, d_dgsum,
            int(accumulate), stream or None))

    def dev_grad_v_sum_peer(self, shards, n_total, eps2, targets, sources, d_gsum,
                            accumulate=False, stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_grad_v_sum_peer(self._h, C.byref(shards), n_total, eps2, t0, t1,
                                                   s0, s1, d_gsum, int(accumulate), stream or None))

    def dev_v_sum_peer(self, shards, n_total, eps2, targets, sources, d_phi, accumulate=False,
                       stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_v_sum_peer(self._h, C.byref(shards), n_total, eps2, t0, t1, s0,
                                              s1, d_phi, int(accumulate), stream or None))

    def dev_grad_v_dgrad_v_sum_tangent_peer(self, shards, n_total, K, eps2, targets, sources,
                                            d_gsum_tan, d_dgsum_tan, stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_grad_v_dgrad_v_sum_tangent_peer(
            self._h, C.byref(shards), n_total, K, eps2, t0, t1, s0, s1, d_gsum_tan, d_dgsum_tan,
            stream or None))

    def dev_kick_sum(self, d_packed, n_total, eta, dt, targets, sources, d_dv, d_tmin=0, stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_kick_sum(self._h, d_packed, n_total, eta, dt, t0, t1, s0, s1,
                                            d_dv, d_tmin or None, stream or None))

    def dev_kick_sum_peer(self, shards, n_total, eta, dt, targets, sources, d_dv, d_tmin=0,
                          stream=0):
        (t0, t1), (s0, s1) = targets, sources
        self._ck(self._lib.nbk_dev_kick_sum_peer(self._h, C.byref(shards), n_total, eta, dt, t0, t1,
                                                 s0, s1, d_dv, d_tmin or None, stream or None))
