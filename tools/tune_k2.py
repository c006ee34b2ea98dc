#!/usr/bin/env python
"""Times every compiled tile shape of the acc+jerk sweep kernel (nbk_tune) at N (default 131072)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
variants = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, 2, 3, 0]
m, q, p = nbh.make_plummer(n, 20243)
with nbk.Context(0) as ctx:
    peak, _ = ctx.fp64_peak(20000)
    peak, _ = ctx.fp64_peak(20000)
    print(f"DFMA peak {peak:.2f} TFLOP/s")
    ref = None
    for v in variants:
        ctx.tune(v)
        best = 1e9
        for _ in range(3):
            g, dg = ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
            best = min(best, ctx.last_kernel_ms())
        if ref is None:
            ref = (g, dg)
        err = max(np.max(np.linalg.norm(g - ref[0], axis=1) / np.linalg.norm(ref[0], axis=1)),
                  np.max(np.linalg.norm(dg - ref[1], axis=1) / np.linalg.norm(ref[1], axis=1)))
        rate = n * (n - 1.0) / (best * 1e-3)
        print(f"variant {v:2d}: {best:8.3f} ms  {rate / 1e9:7.1f} G int/s  "
              f"{60 * rate / 1e12:6.2f} TF(60)  {60 * rate / 1e12 / peak * 100:5.1f}% of DFMA peak  "
              f"(max rel diff vs first {err:.1e})")
