"""Pair-once (sym.cuh) against the ordered acc+jerk sweep: per-body agreement with the oracle /
with each other and kernel times over a range of N; item ranges add up to the full result.
    python tools/time_sym.py [N ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import nbk  # noqa: E402
from oracle import oracle as O  # noqa: E402


def rel(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)))


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [2048, 3000, 5000, 16384, 32768, 65536, 131072]
    with nbk.Context(0) as ctx:
        for n in sizes:
            m, q, p = O.make_plummer(n, 1000 + n)
            for eps2 in (0.0, 1e-4):
                ctx.set_sym(0)
                ctx.grad_v_dgrad_v_sum(m, q, p, eps2)
                g0, dg0 = ctx.grad_v_dgrad_v_sum(m, q, p, eps2)
                t0 = ctx.last_kernel_ms()
                ctx.set_sym(2)
                ctx.grad_v_dgrad_v_sum(m, q, p, eps2)
                g1, dg1 = ctx.grad_v_dgrad_v_sum(m, q, p, eps2)
                t1 = ctx.last_kernel_ms()
                ctx.set_sym(1)
                g2, dg2 = ctx.grad_v_dgrad_v_sum(m, q, p, eps2)
                t2 = ctx.last_kernel_ms()
                line = (f"N={n} eps2={eps2:g}: ordered {t0:.3f} ms, pair-once {t1:.3f} ms, auto {t2:.3f} ms; "
                        f"pair-once vs ordered: acc {rel(g1, g0):.2e} jerk {rel(dg1, dg0):.2e}")
                if n <= 8192:
                    gr, dgr = O.grad_v_dgrad_v_sum(m, q, p, eps2)
                    line += (f"; vs oracle: acc {rel(g1, gr):.2e} jerk {rel(dg1, dgr):.2e} "
                             f"(ordered {rel(g0, gr):.2e} {rel(dg0, dgr):.2e})")
                print(line, flush=True)


if __name__ == "__main__":
    main()
