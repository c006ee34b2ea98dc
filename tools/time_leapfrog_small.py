#!/usr/bin/env python
"""Leapfrog.advance with deep block recursions (hot spheres, small eta): device-side
advance_subsystem (resident=2) against the host recursion over resident bodies (resident=1)."""
import sys, time
sys.path[:0] = ['.', 'direct-nbody_b200']
import numpy as np, nbh, nbk
with nbk.Context(0) as ctx:
    for n, span, eta in ((100, 1/32, 1e-3), (512, 1/64, 1e-3), (2000, 1/64, 1e-3)):
        m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_hot_spherical(n, 5))
        nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span / 64, eta)
        for mode in (2, 1):
            t0 = time.perf_counter()
            a = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span, eta, resident=mode)
            dt = time.perf_counter() - t0
            print(f"Leapfrog.advance N={n} span {span:g} eta {eta:g} resident={mode}: {a.nkicks} kick blocks in {dt:.3f} s = {1e6*dt/max(a.nkicks,1):.1f} us/block", flush=True)
