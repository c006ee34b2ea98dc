#!/usr/bin/env python
"""BASELINE configs[1] with block time steps: Leapfrog.advance on the Plummer N=16384 system,
device recursion of the small subtrees (resident=2) against the host recursion (resident=1).
usage: tools/time_leapfrog_cfg1.py [N] [span] [eta]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import nbh  # noqa: E402
import nbk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
span = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0 / 256
eta = float(sys.argv[3]) if len(sys.argv) > 3 else 2e-2
with nbk.Context(0) as ctx:
    m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(n, 20241))
    nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span / 64, eta)
    for mode in (2, 1):
        t0 = time.perf_counter()
        a = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span, eta, resident=mode)
        dt = time.perf_counter() - t0
        ke, pe = nbh.energy(ctx, a.m, a.q, a.p)
        print(f"Leapfrog.advance N={n} span {span:g} eta {eta:g} resident={mode}: {a.nkicks} kick blocks in "
              f"{dt:.3f} s = {1e6 * dt / max(a.nkicks, 1):.1f} us/block, dE/E = {abs((ke + pe + 0.25) / 0.25):.3e}",
              flush=True)
