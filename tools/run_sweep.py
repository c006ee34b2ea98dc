#!/usr/bin/env python
"""Runs a few sweeps of one kernel through the C ABI (profiling target for ncu; no oracle).
usage: tools/run_sweep.py [op] [n] [reps]     op in accjerk|acc|pot|kick|kickt|tangent (K = 6)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import nbh  # noqa: E402
import nbk  # noqa: E402

op = sys.argv[1] if len(sys.argv) > 1 else "accjerk"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
shape = int(sys.argv[4]) if len(sys.argv) > 4 else 0  # nbk_tune: 0 auto, 1 large, 2 medium, 3 small
m, q, p = nbh.make_plummer(n, 20243)
v = p / m[:, None]
if op == "tangent":
    import numpy as np
    rng = np.random.default_rng(20245)
    dq = rng.standard_normal((6, n, 3))
    dp = rng.standard_normal((6, n, 3)) * m[None, :, None]
with nbk.Context(0) as ctx:
    ctx.tune(shape)
    for _ in range(reps):
        if op == "accjerk":
            ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
        elif op == "acc":
            ctx.grad_v_sum(m, q, 0.0)
        elif op == "pot":
            ctx.v_sum(m, q, 0.0)
        elif op == "kick":
            ctx.kick_sum(m, q, v, float("inf"), 1e-3, want_tmin=False)
        elif op == "kickt":
            ctx.kick_sum(m, q, v, 1e-3, 1e-3)
        elif op == "tangent":
            ctx.grad_v_dgrad_v_sum_tangent(m, q, p, dq, dp, 0.0)
        ms = ctx.last_kernel_ms()
        k = 6 if op == "tangent" else 1
        print(f"{op} n={n}: kernel {ms:.3f} ms, {k * n * (n - 1.0) / ms / 1e6:.1f} G interactions/s")
