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


def expected(inputs_dir=None):
    """{file name: text} of the oracle's outputs for the committed inputs."""
    d = inputs_dir or os.path.join(HERE, "inputs")
    big = read_bodies(os.path.join(d, "plummer_n64_seed20240.txt"))
    small = read_bodies(os.path.join(d, "plummer_n16_std_seed20241.txt"))
    return {"pairs.hex": pairs(*big), "sums.hex": sums(*big), "traj.hex": traj(*small)}


if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "inputs"), exist_ok=True)
    os.makedirs(os.path.join(HERE, "expected"), exist_ok=True)
    write_bodies(os.path.join(HERE, "inputs", "plummer_n64_seed20240.txt"), *O.make_plummer(64, 20240))
    write_bodies(os.path.join(HERE, "inputs", "plummer_n16_std_seed20241.txt"),
                 *O.rescale_to_standard_units(*O.make_plummer(16, 20241)))
    for name, text in expected().items():
        open(os.path.join(HERE, "expected", name), "w").write(text)
    print("wrote inputs/ and expected/")
