#!/usr/bin/env python
"""Times Hermite.advance (BASELINE configs[0]: Plummer N=1024, 1 N-body time unit) on the
device-resident stepping loop; prints steps, us/step and the energy error."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import nbh  # noqa: E402
import nbk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eta = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
span = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
with nbk.Context(0) as ctx:
    m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(n, 20240))
    e0 = sum(nbh.energy(ctx, m, q, p))
    nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), 0.01, eta)
    t0 = time.perf_counter()
    a = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), span, eta)
    dt = time.perf_counter() - t0
    de = abs((sum(nbh.energy(ctx, a.m, a.q, a.p)) - e0) / e0)
    print(f"N={n} eta={eta}: {a.nsteps} steps in {dt:.3f} s = {1e6 * dt / a.nsteps:.2f} us/step, dE/E = {de:.4e}")
