#!/usr/bin/env python
"""Writes oracle/ref_fixtures/inputs/*.txt (seeded bodies as IEEE-754 bit patterns, read by
gen.ml) and expected/*.hex: what the ORACLE returns for exactly the calls gen.ml makes to the
reference, in gen.ml's output format.  `diff -r expected ref` after running gen.ml on a machine
with OCaml is the pin; tests/test_ref_fixtures.py does the same comparison.
Test infrastructure - regenerate with:  python oracle/ref_fixtures/make_inputs.py"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def bits(x):
    return "%016x" % struct.unpack("<Q", struct.pack("<d", float(x)))[0]


def write_bodies(path, m, q, p):
    with open(path, "w") as f:
        f.write(f"n {len(m)}\n")
        for i in range(len(m)):
            f.write(" ".join(bits(x) for x in (m[i], *q[i], *p[i])) + "\n")


def read_bodies(path):
    rows = [ln.split() for ln in open(path).read().splitlines()[1:]]
    a = np.array([[struct.unpack("<d", struct.pack("<Q", int(w, 16)))[0] for w in r] for r in rows])
    return a[:, 0].copy(), a[:, 1:4].copy(), a[:, 4:7].copy()


class Out:
    def __init__(self):
        self.lines = []

    def emit(self, key, i, j, k, x):
        self.lines.append(f"{key} {i} {j} {k} {bits(x)}")

    def emit3(self, key, i, j, v):
        for k in range(3):
            self.emit(key, i, j, k, v[k])

    def text(self):
        return "\n".join(self.lines) + "\n"


def pairs(m, q, p):
    o, npair = Out(), 12
    for e, eps2 in enumerate((0.0, 0.0016)):
        for i in range(npair):
            for j in range(i + 1, npair):
                tg, sr = (i, i + 1), (j, j + 1)
                o.emit(f"v{e}", i, j, 0, O.v_sum(m, q, eps2, targets=tg, sources=sr)[0])
                o.emit3(f"grad_v{e}", i, j, O.grad_v_sum(m, q, eps2, targets=tg, sources=sr)[0])
                g, dg = O.grad_v_dgrad_v_sum(m, q, p, eps2, targets=tg, sources=sr)
                o.emit3(f"gvdgv_g{e}", i, j, g[0])
                o.emit3(f"gvdgv_dg{e}", i, j, dg[0])
    v = p / m[:, None]
    for i in range(npair):
        for j in range(i + 1, npair):
            idx = [i, j]
            nv, ndt = O.kick_all_pairs(m[idx], q[idx], v[idx], np.full(2, np.inf), 1e-3, 1e-4)
            o.emit3("kick_v1", i, j, nv[0])
            o.emit3("kick_v2", i, j, nv[1])
            o.emit("kick_dtmax1", i, j, 0, ndt[0])
            o.emit("kick_dtmax2", i, j, 0, ndt[1])
    return o.text()


def sums(m, q, p):
    o = Out()
    f, df = O.start_bs_sweep(m, q, p, 0.0)
    for i in range(len(m)):
        o.emit3("start_f", i, 0, f[i])
        o.emit3("start_df", i, 0, df[i])
        # start_timescale, hermite.ml:145-146: 6 eta |f| / (m |df|) with Base.norm
        nf = np.sqrt(f[i][0] * f[i][0] + f[i][1] * f[i][1] + f[i][2] * f[i][2])
        ndf = np.sqrt(df[i][0] * df[i][0] + df[i][1] * df[i][1] + df[i][2] * df[i][2])
        o.emit("start_h", i, 0, 0, 6.0 * 1e-2 * nf / (m[i] * ndf))
    ke, pe = O.total_kinetic_energy(m, p), O.total_potential_energy(m, q, 0.0)
    o.emit("ke", 0, 0, 0, ke)
    o.emit("pe", 0, 0, 0, pe)
    o.emit("energy", 0, 0, 0, ke + pe)
    return o.text()


def traj(m, q, p):
    o = Out()
    h = O.hermite_advance(O.HermiteState(m, q, p), 0.125, 1e-4)
    for i in range(len(m)):
        o.emit("hermite_t", i, 0, 0, h.t[i])
        o.emit("hermite_h", i, 0, 0, h.h[i])
        for key, arr in (("hermite_q", h.q), ("hermite_p", h.p), ("hermite_f", h.f), ("hermite_df", h.df)):
            o.emit3(key, i, 0, arr[i])
    o.emit("hermite_nsteps", 0, 0, 0, float(h.nsteps))
    for key, s in (("lfc", O.leapfrog_advance_const_dt(O.LeapfrogState(m, q, p), 1.0 / 256, 8)),
                   ("lfa", O.leapfrog_advance(O.LeapfrogState(m, q, p), 1.0 / 64, 1e-2))):
        for r in range(len(m)):
            o.emit(f"{key}_id", r, 0, 0, float(s.id[r] - 1))
            o.emit(f"{key}_t", r, 0, 0, s.t[r])
            o.emit(f"{key}_dtmax", r, 0, 0, s.dt_max[r])
            o.emit3(f"{key}_q", r, 0, s.q[r])
            o.emit3(f"{key}_v", r, 0, s.v[r])
    oq, op, ot = O.omelyan_advance(m, q, p, 0.0, 1.0 / 64)
    for i in range(len(m)):
        o.emit("oa_t", i, 0, 0, ot)
        o.emit3("oa_q", i, 0, oq[i])
        o.emit3("oa_p", i, 0, op[i])
    a = O.advancer_advance(O.AdvancerState(m, q, p), 0.125, 1e-3, 0.0)
    for r in range(len(m)):
        o.emit("adv_id", r, 0, 0, float(a.id[r]))
        o.emit("adv_t", r, 0, 0, a.t[r])
        o.emit("adv_h", r, 0, 0, a.h[r])
        o.emit("adv_hmax", r, 0, 0, a.hmax[r])
        o.emit3("adv_q", r, 0, a.q[r])
        o.emit3("adv_p", r, 0, a.p[r])
    return o.text()


def _norm3(v):
    return np.sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2])


def _emit_advancer(o, key, a):
    for r in range(len(a.id)):
        o.emit(f"{key}_id", r, 0, 0, float(a.id[r]))
        o.emit(f"{key}_t", r, 0, 0, a.t[r])
        o.emit(f"{key}_h", r, 0, 0, a.h[r])
        o.emit(f"{key}_hmax", r, 0, 0, a.hmax[r])
        o.emit3(f"{key}_q", r, 0, a.q[r])
        o.emit3(f"{key}_p", r, 0, a.p[r])


def wide(big, small):
    """The oracle's answers to gen_wide.ml's calls, in its order."""
    o = Out()
    m, q, p = big
    sm, sq, sp = small
    eps2 = 0.2 * 0.2  # Base.set_eps 0.2 = square 0.2 (base.ml:30-31)
    # -- softened()
    f, df = O.start_bs_sweep(m, q, p, eps2)
    for i in range(len(m)):
        o.emit3("eps_start_f", i, 0, f[i])
        o.emit3("eps_start_df", i, 0, df[i])
        o.emit("eps_start_h", i, 0, 0, 6.0 * 1e-2 * _norm3(f[i]) / (m[i] * _norm3(df[i])))
    ke, pe = O.total_kinetic_energy(m, p), O.total_potential_energy(m, q, eps2)
    o.emit("eps_ke", 0, 0, 0, ke)
    o.emit("eps_pe", 0, 0, 0, pe)
    o.emit("eps_energy", 0, 0, 0, ke + pe)
    # Leapfrog bodies keep v = p / m and B.p hands back v * m (leapfrog.ml:31-43)
    plf = (p / m[:, None]) * m[:, None]
    kelf = O.total_kinetic_energy(m, plf)
    o.emit("eps_lf_ke", 0, 0, 0, kelf)
    o.emit("eps_lf_energy", 0, 0, 0, kelf + pe)
    h = O.hermite_advance(O.HermiteState(sm, sq, sp), 0.0625, 1e-3, eps2)
    h = O.hermite_advance(h, 0.0625, 1e-3, eps2)
    for i in range(len(sm)):
        o.emit("h2_t", i, 0, 0, h.t[i])
        o.emit("h2_h", i, 0, 0, h.h[i])
        for key, arr in (("h2_q", h.q), ("h2_p", h.p), ("h2_f", h.f), ("h2_df", h.df)):
            o.emit3(key, i, 0, arr[i])
    o.emit("h2_nsteps", 0, 0, 0, float(h.nsteps))
    a = O.advancer_advance(O.AdvancerState(sm, sq, sp), 0.0625, 1e-3, eps2)
    a = O.advancer_advance(a, 0.0625, 1e-3, eps2)
    _emit_advancer(o, "a2", a)
    o.emit("a2_energy", 0, 0, 0, O.energy(a.m, a.q, a.p, eps2))
    oq, op, ot = sq, sp, 0.0
    for _ in range(4):
        oq, op, ot = O.omelyan_advance(sm, oq, op, ot, 1.0 / 128.0, eps2)
    for i in range(len(sm)):
        o.emit("oa4_t", i, 0, 0, ot)
        o.emit3("oa4_q", i, 0, oq[i])
        o.emit3("oa4_p", i, 0, op[i])
    # -- unsoftened()
    s = O.leapfrog_advance(O.LeapfrogState(sm, sq, sp), 1.0 / 32.0, 1e-3)
    for r in range(len(sm)):
        o.emit("lf3_id", r, 0, 0, float(s.id[r] - 1))
        o.emit("lf3_t", r, 0, 0, s.t[r])
        o.emit("lf3_dtmax", r, 0, 0, s.dt_max[r])
        o.emit3("lf3_q", r, 0, s.q[r])
        o.emit3("lf3_v", r, 0, s.v[r])
    o.emit("lf3_energy", 0, 0, 0, O.energy(s.m, s.q, s.v * s.m[:, None], 0.0))
    a = O.advancer_advance(O.AdvancerState(sm, sq, sp), 0.0625, 1e-2, 0.0, extpot="plummer_test")
    _emit_advancer(o, "ax", a)
    bi, bj, be = O.binaries(m, q, p, 0.0, 0.0)
    # Analysis.binaries conses as it scans: the list is the scan reversed (analysis.ml:836-847)
    for k, (i, j, e) in enumerate(zip(bi[::-1], bj[::-1], be[::-1])):
        o.emit("bin_i", k, 0, 0, float(i))
        o.emit("bin_j", k, 0, 0, float(j))
        o.emit("bin_e", k, 0, 0, e)
    ti, tj, te = O.binaries(m, q, p, 0.0, -abs(1e-4))
    for k, (i, j, e) in enumerate(zip(ti[::-1], tj[::-1], te[::-1])):
        o.emit("thr_i", k, 0, 0, float(i))
        o.emit("thr_j", k, 0, 0, float(j))
        o.emit("thr_e", k, 0, 0, e)
    # tightest_binary folds from the head and replaces on a strictly larger |e| (analysis.ml:863-876)
    best = 0
    rev = list(zip(bi[::-1], bj[::-1], be[::-1]))
    for k in range(1, len(rev)):
        if abs(rev[k][2]) > abs(rev[best][2]):
            best = k
    o.emit("tight_i", 0, 0, 0, float(rev[best][0]))
    o.emit("tight_j", 0, 0, 0, float(rev[best][1]))
    el = O.bodies_to_el_phase_space(m, q, p, 0.0)
    for i in range(len(m)):
        for k in range(4):
            o.emit("el", i, 0, k, el[i, k])
    a = O.advancer_advance(O.AdvancerState(sm, sq, sp), 0.03125, 1e-3, 0.0)
    a = O.advancer_advance(a, 0.03125, 1e-3, 0.0)
    _emit_advancer(o, "int", a)
    return o.text()


def ics(big):
    o = Out()
    fm, fq, fp = O.adjust_frame(*big)
    sm, sq, sp = O.rescale_to_standard_units(fm, fq, fp)
    for key, (m, q, p) in (("frame", (fm, fq, fp)), ("std", (sm, sq, sp))):
        for i in range(len(m)):
            o.emit(f"{key}_m", i, 0, 0, m[i])
            o.emit3(f"{key}_q", i, 0, q[i])
            o.emit3(f"{key}_p", i, 0, p[i])
    return o.text()


def dfloat(small):
    o = Out()
    m, q, p = small
    for key, n, dt, sf in (("n3", 3, 0.03125, 1e-2), ("n5", 5, 0.015625, 1e-2)):
        jac, _ = O.hermite_jacobian_advance(m[:n], q[:n], p[:n], dt, sf)
        for i in range(6 * n):
            for j in range(6 * n):
                o.emit(f"{key}_jac", i, j, 0, jac[i, j])
        o.emit(f"{key}_omega", 0, 0, 0, O.omega_pushforward(jac))
    return o.text()


def unit(bodies):
    """The oracle's answers to gen_unit.ml: four quarter-time-unit advances of Hermite and
    Leapfrog, energy and state after each."""
    o = Out()
    m, q, p = bodies
    n = len(m)
    h = O.HermiteState(m, q, p)
    o.emit("he_energy", 0, 0, 0, O.energy(h.m, h.q, h.p, 0.0))
    for k in range(1, 5):
        h = O.hermite_advance(h, 0.25, 1e-4)
        o.emit("he_energy", k, 0, 0, O.energy(h.m, h.q, h.p, 0.0))
        o.emit("he_nsteps", k, 0, 0, float(h.nsteps))
        for i in range(n):
            o.emit("he_t", k, i, 0, h.t[i])
            o.emit3("he_q", k, i, h.q[i])
            o.emit3("he_p", k, i, h.p[i])
    s = O.LeapfrogState(m, q, p)
    o.emit("lf_energy", 0, 0, 0, O.energy(s.m, s.q, s.v * s.m[:, None], 0.0))
    for k in range(1, 5):
        s = O.leapfrog_advance(s, 0.25, 2e-3)
        o.emit("lf_energy", k, 0, 0, O.energy(s.m, s.q, s.v * s.m[:, None], 0.0))
        for r in range(n):
            o.emit("lf_id", k, r, 0, float(s.id[r] - 1))
            o.emit("lf_t", k, r, 0, s.t[r])
            o.emit3("lf_q", k, r, s.q[r])
            o.emit3("lf_v", k, r, s.v[r])
    return o.text()


def expected(inputs_dir=None):
    """{file name: text} of the oracle's outputs for the committed inputs."""
    d = inputs_dir or os.path.join(HERE, "inputs")
    big = read_bodies(os.path.join(d, "plummer_n64_seed20240.txt"))
    small = read_bodies(os.path.join(d, "plummer_n16_std_seed20241.txt"))
    mid = read_bodies(os.path.join(d, "plummer_n24_std_seed20242.txt"))
    return {"pairs.hex": pairs(*big), "sums.hex": sums(*big), "traj.hex": traj(*small),
            "wide.hex": wide(big, small), "ics.hex": ics(big), "dfloat.hex": dfloat(small),
            "unit.hex": unit(mid)}


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "expected"), exist_ok=True)
    write_bodies(os.path.join(HERE, "inputs", "plummer_n64_seed20240.txt"), *O.make_plummer(64, 20240))
    write_bodies(os.path.join(HERE, "inputs", "plummer_n16_std_seed20241.txt"),
                 *O.rescale_to_standard_units(*O.make_plummer(16, 20241)))
    write_bodies(os.path.join(HERE, "inputs", "plummer_n24_std_seed20242.txt"),
                 *O.rescale_to_standard_units(*O.make_plummer(24, 20242)))
    for name, text in expected().items():
        open(os.path.join(HERE, "expected", name), "w").write(text)
    print("wrote inputs/ and expected/")
