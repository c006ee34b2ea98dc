#!/usr/bin/env python
"""Small-size tour of every kernel through the C ABI, for compute-sanitizer (one tool per run):
  compute-sanitizer --tool memcheck  python tools/sanitize_driver.py
  compute-sanitizer --tool racecheck python tools/sanitize_driver.py
No oracle, no timing: it only has to touch every code path (all tile shapes, ragged and
self-overlapping ranges, few-target kernels, tangent tiles, resident Hermite/leapfrog loops,
binaries scan/list)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
assert n >= 1500, "the tour's fixed ranges need n >= 1500"
m, q, p = nbh.make_plummer(n, 5)
v = p / m[:, None]
rng = np.random.default_rng(1)
with nbk.Context(0) as ctx:
    for shape in (0, 1, 2, 3):
        ctx.tune(shape)
        ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
        ctx.grad_v_dgrad_v_sum(m, q, p, 1e-4, targets=(100, 1301), sources=(7, 1493))
        ctx.grad_v_sum(m, q, 0.0)
        ctx.v_sum(m, q, 0.0)
        ctx.kick_sum(m, q, v, 1e-3, 1e-4)
        ctx.kick_sum(m, q, v, 1e-3, 1e-4, want_tmin=False)
    ctx.tune(0)
    for nt in (1, 3, 32):
        ctx.grad_v_dgrad_v_sum(m, q, p, 0.0, targets=(500, 500 + nt))
        ctx.kick_sum(m, q, v, 1e-3, 1e-4, targets=(500, 500 + nt))
        ctx.v_sum(m, q, 0.0, targets=(500, 500 + nt))
    ctx.energy(m, q, p, 0.0)
    ctx.interact_sum(m, q, 5, 1e-4)
    ctx.kick_block_sum(m, q, v, 1e-3, 1e-4, n - 7)
    for K in (1, 3, 4):
        dq = rng.standard_normal((K, n, 3))
        dp = rng.standard_normal((K, n, 3)) * m[None, :, None]
        ctx.grad_v_dgrad_v_sum_tangent(m, q, p, dq, dp, 1e-4)
        ctx.grad_v_sum_tangent(m, q, dq, 1e-4)
    g, dg = ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
    t = rng.uniform(0, 1e-3, n)
    ctx.grad_v_dgrad_v_sum_predicted(t, m, q, p, -g, -dg, 2e-3, 0.0, targets=(3, 4))
    ctx.bodies_upload(m, q, v, is_velocity=True)
    ctx.bodies_update(10, 20, q[10:20], v[10:20], is_velocity=True)
    ctx.resident_grad_v_dgrad_v_sum(0.0, targets=(0, 40))
    ctx.resident_leapfrog_steps(1e-3, 2)
    ctx.binary_scan(m, q, p, 0.0, 0.0)
    ctx.binary_list(m, q, p, 0.0, 0.0)
    # resident Hermite: launch-per-step path and the persistent loop
    s = nbh.HermiteState(m[:300] * (n / 300), q[:300], p[:300] * (n / 300))
    nbh.load_library().nbh_set_step_cap(2_000_000)
    nbh.hermite_advance(ctx, s, 1e-3, 1e-4, mode=1)
    nbh.hermite_advance(ctx, s, 1e-3, 1e-4, mode=2)
    # resident integrators: Advancer.A (per-phase kernels, then the device-side block recursion
    # in its default and in forced grid shapes), Leapfrog block steps, Omelyan steps
    mm, qq, pp = nbh.rescale_to_standard_units(ctx, m[:300] * (n / 300), q[:300], p[:300] * (n / 300))
    a0 = nbh.AdvancerState(mm, qq, pp)
    nbh.advancer_advance(ctx, a0, 2.0 ** -9, 1e-4, 1e-4, resident=1)
    a1 = nbh.advancer_advance(ctx, a0, 2.0 ** -9, 1e-4, 1e-4, resident=2)
    nbh.advancer_advance(ctx, a1, 2.0 ** -9, 1e-4, 1e-4, resident=2, tree_max=40)
    for shape in ("128,2", "256,1", "512,1"):
        os.environ["NBK_ADT_SHAPE"] = shape
        nbh.advancer_advance(ctx, a1, 2.0 ** -10, 1e-4, 1e-4, resident=2)
    del os.environ["NBK_ADT_SHAPE"]
    nbh.leapfrog_advance(ctx, nbh.LeapfrogState(mm, qq, pp), 2.0 ** -9, 5e-3)
    nbh.omelyan_advance_steps(ctx, mm, qq, pp, 0.0, 2.0 ** -9, 2, 1e-4, want_el=True)
    print("sanitize tour done:", ctx.launch_count, "launches")
