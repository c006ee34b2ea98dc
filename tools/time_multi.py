#!/usr/bin/env python
"""Single-process multi-GPU context (nbk_create_multi): the metric sweep through the HOST-buffer
call at N = 131072 on 1..all GPUs of the box, wall-clock per call (H2D + broadcast + sweeps +
D2H)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
what = sys.argv[2] if len(sys.argv) > 2 else "accjerk"   # accjerk | tangent (K = 6) | kick
m, q, p = nbh.make_plummer(n, 20243)
if what != "accjerk":
    # BASELINE configs[4] from ONE process (tangent acc+jerk sums, K = 6), or the kick with its
    # bit-exact time-scale, through the same multi context
    K = 6
    rng = np.random.default_rng(20245)
    dq = rng.standard_normal((K, n, 3))
    dp = rng.standard_normal((K, n, 3)) * m[None, :, None]
    v = p / m[:, None]
    ref = None
    g = 1
    while g <= torch.cuda.device_count():
        with nbk.Context(devices=list(range(g))) as ctx:
            call = (lambda: ctx.grad_v_dgrad_v_sum_tangent(m, q, p, dq, dp, 0.0)) if what == "tangent" \
                else (lambda: ctx.kick_sum(m, q, v, 1e-3, 1e-4))
            call()
            t0 = time.perf_counter()
            for _ in range(5):
                out = call()
            dt = (time.perf_counter() - t0) / 5
        if ref is None:
            ref = out
        last = out[3] if what == "tangent" else out[0]
        first = ref[3] if what == "tangent" else ref[0]
        diff = float(np.max(np.linalg.norm(first - last, axis=-1) / np.linalg.norm(first, axis=-1)))
        mult = K if what == "tangent" else 1
        print(f"{what} N={n}: {g} GPU(s), one process: {1e3 * dt:.2f} ms per host-buffer call = "
              f"{mult * n * (n - 1.0) / dt / 1e9:.1f} G interactions/s, max rel diff vs 1 GPU: {diff:.1e}"
              + ("" if what == "tangent" else f", tmin bit-identical: {bool(np.array_equal(ref[1], out[1]))}"))
        g *= 2
    sys.exit(0)
ref = None
g = 1
while g <= torch.cuda.device_count():
    with nbk.Context(devices=list(range(g))) as ctx:
        ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
        t0 = time.perf_counter()
        for _ in range(5):
            out = ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
        dt = (time.perf_counter() - t0) / 5
    if ref is None:
        ref = out
    diff = float(np.max(np.linalg.norm(ref[1] - out[1], axis=1) / np.linalg.norm(ref[1], axis=1)))
    print(f"{g} GPU(s), one process: {1e3 * dt:.2f} ms per host-buffer call = "
          f"{n * (n - 1.0) / dt / 1e9:.1f} G interactions/s, max rel diff of jerk vs 1 GPU: {diff:.1e}")
    g *= 2
