#!/usr/bin/env python
"""Times Advancer.A.advance (the integrator bin/nbody.ml and bin/el_mass_segregation.ml run) with
the records resident on the device against the host-buffer path.
usage: [ADV_MODES=2,1,0] [ADV_TREE_MAX=rows] tools/time_advancer.py N span [eps] [two_mass]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
span = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0 / 16
eps = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
two_mass = len(sys.argv) > 4
with nbk.Context(0) as ctx:
    m, q, p = nbh.make_plummer(n, 20244)
    if two_mass:
        m, q, p = nbh.add_mass_disc(m, q, p, 20244)
        m, q, p = nbh.adjust_frame(m, q, p)
    m, q, p = nbh.rescale_to_standard_units(ctx, m, q, p)
    e0 = sum(nbh.energy(ctx, m, q, p, eps * eps))
    s0 = nbh.AdvancerState(m, q, p)
    nbh.advancer_advance(ctx, s0, span / 64, 1e-4, eps * eps)  # warm-up
    modes = [int(x) for x in os.environ.get("ADV_MODES", "2,1,0").split(",")]
    tree_max = int(os.environ.get("ADV_TREE_MAX", "0"))
    for resident in modes:  # 2: block recursion on the device, 1: resident records, 0: host buffers
        t0 = time.perf_counter()
        a = nbh.advancer_advance(ctx, s0, span, 1e-4, eps * eps, resident=resident, tree_max=tree_max)
        dt = time.perf_counter() - t0
        de = abs((sum(nbh.energy(ctx, a.m, a.q, a.p, eps * eps)) - e0) / e0)
        if resident == 2:
            import ctypes
            prof = (ctypes.c_longlong * 11)()
            nbk.load_library().nbk_adv_range_profile(ctx._h, prof, 1)
            names = ("control", "fold+actives", "predict", "slot x active", "active x tile", "barrier",
                     "finish", "barrier", "sort")
            tot = sum(prof[1:10]) or 1
            print(f"  device recursion: {prof[10]} launches, {prof[0]} block steps, CTA 0 cycles "
                  f"{tot / 1.965e3:.0f} us: " +
                  ", ".join(f"{nm} {100.0 * c / tot:.1f}%" for nm, c in zip(names, prof[1:10])))
        print(f"N={n} span={span} resident={resident}: {a.stats[0]} interact_bodies calls, "
              f"{a.stats[1]:.3e} pair interactions in {dt:.3f} s = {1e6 * dt / a.stats[0]:.1f} us/call, "
              f"dE/E = {de:.3e}")
