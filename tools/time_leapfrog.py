#!/usr/bin/env python
"""Times const-dt leapfrog steps (BASELINE configs[1]: Plummer N=16384) on the device-resident
path and through the host recursion (one kick sweep per step)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nst = int(sys.argv[2]) if len(sys.argv) > 2 else 64
with nbk.Context(0) as ctx:
    m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(n, 20241))
    v = p / m[:, None]
    for rep in range(2):
        ctx.bodies_upload(m, q, v, is_velocity=True)
        t0 = time.perf_counter()
        ctx.resident_leapfrog_steps(1.0 / 1024, nst)
        dt = time.perf_counter() - t0
        print(f"resident ABI call: N={n}, {nst} steps: {1e3 * dt / nst:.3f} ms/step (last kick kernel {ctx.last_kernel_ms():.3f} ms)")
    s = nbh.LeapfrogState(m, q, p)
    s.dt_max[:] = np.inf
    for resident in (True, False):
        t0 = time.perf_counter()
        a = nbh.leapfrog_advance_const_dt(ctx, s, 1.0 / 1024, nst, resident=resident)
        dt = time.perf_counter() - t0
        ke, pe = nbh.energy(ctx, a.m, a.q, a.p)
        print(f"host mirror resident={resident}: {1e3 * dt / nst:.3f} ms/step, dE/E = {abs((ke + pe + 0.25) / 0.25):.3e}")
    # adaptive block steps (Leapfrog.advance): bodies resident on the device vs one
    # upload + read-back per kick block
    na = min(n, 4096)
    m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(na, 20241))
    span, eta = 1.0 / 256, 5e-3
    for resident in (2, 1, 0):  # 2: block recursion of the small subtrees on the device too
        t0 = time.perf_counter()
        a = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span, eta, resident=resident)
        dt = time.perf_counter() - t0
        print(f"Leapfrog.advance N={na}, span {span:g}, eta {eta:g}, resident={resident}: {a.nkicks} kick "
              f"blocks in {dt:.3f} s = {1e6 * dt / max(a.nkicks, 1):.1f} us/block")
