#!/usr/bin/env python
"""Generates tests/golden/*.npz: fixed-seed inputs and the ORACLE's outputs for them.

The reference (OCaml) cannot be run in the build image and ships no golden vectors
(SURVEY.md §8c), so these fixtures pin the oracle restatement itself: tests/test_oracle.py
asserts that the oracle, however and wherever it is compiled (e.g. on the GPU box), reproduces
them BIT FOR BIT (guards -ffp-contract / libm drift), and the GPU parity tests read the same
files.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def sweeps(n, seed, eps):
    m, q, p = O.make_plummer(n, seed)
    eps2 = eps * eps
    g, dg = O.grad_v_dgrad_v_sum(m, q, p, eps2)
    f, df = O.start_bs_sweep(m, q, p, eps2)
    gg = O.grad_v_sum(m, q, eps2)
    phi = O.v_sum(m, q, eps2)
    v = p / m[:, None]
    dv, tmin = O.kick_sum(m, q, v, 1e-3, 1e-4)
    el = O.bodies_to_el_phase_space(m, q, p, eps2)
    return dict(m=m, q=q, p=p, eps2=eps2, gsum=g, dgsum=dg, f=f, df=df, grad=gg, phi=phi,
                pe=O.total_potential_energy(m, q, eps2), ke=O.total_kinetic_energy(m, p),
                kick_dv=dv, kick_tmin=tmin, el=el)


def trajectories():
    out = {}
    m, q, p = O.rescale_to_standard_units(*O.make_plummer(16, 20241))
    out.update(m=m, q=q, p=p)
    h = O.hermite_advance(O.HermiteState(m, q, p), 0.125, 1e-4)
    out.update(hermite_q=h.q, hermite_p=h.p, hermite_h=h.h, hermite_nsteps=h.nsteps)
    lf = O.leapfrog_advance_const_dt(O.LeapfrogState(m, q, p), 1.0 / 256, 8)
    out.update(leapfrog_q=lf.q, leapfrog_v=lf.v)
    la = O.leapfrog_advance(O.LeapfrogState(m, q, p), 1.0 / 64, 1e-2)
    out.update(leapfrog_adapt_id=la.id, leapfrog_adapt_q=la.q, leapfrog_adapt_v=la.v,
               leapfrog_adapt_dtmax=la.dt_max)
    oq, op, ot = O.omelyan_advance(m, q, p, 0.0, 1.0 / 64)
    out.update(omelyan_q=oq, omelyan_p=op, omelyan_t=ot)
    a = O.advancer_advance(O.AdvancerState(m, q, p), 0.125, 1e-3, 0.0)
    out.update(advancer_id=a.id, advancer_state=a.state, advancer_stats=np.array(a.stats))
    return out


def snapshot():
    """particles_advancer.yaml: five bodies in the reference's `--- !Particle` text format
    (A.print, advancer.ml:88-96), written by the oracle's restatement of it."""
    m, q, p = O.rescale_to_standard_units(*O.make_plummer(5, 20246))
    ids = np.arange(5, dtype=np.int32)
    t = np.array([0.0, 0.25, 0.25, 1.0, 1e-3])
    with open(os.path.join(HERE, "particles_advancer.yaml"), "w") as f:
        f.write(O.advancer_print_text(ids, t, m, q, p))


if __name__ == "__main__":
    what = sys.argv[1:] or ["npz", "yaml"]
    if "npz" in what:
        np.savez(os.path.join(HERE, "sweeps_n64_eps0.npz"), **sweeps(64, 20240, 0.0))
        np.savez(os.path.join(HERE, "sweeps_n257_eps004.npz"), **sweeps(257, 20241, 0.04))
        np.savez(os.path.join(HERE, "trajectories_n16.npz"), **trajectories())
    if "yaml" in what:
        snapshot()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith((".npz", ".yaml"))))
