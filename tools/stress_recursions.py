#!/usr/bin/env python
"""Randomised stress of the device-side block recursions (nbk_adv_advance_range,
nbk_lf_advance_subsystem) against the host recursions over the same resident kernels: sizes,
spans, prefix limits and forced grid shapes drawn at random; every case is run twice on the
device (bitwise reproducibility) and compared with the host recursion (same schedule, states to
rounding).  usage: tools/stress_recursions.py [cases] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)


def rel(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=-1) / np.maximum(np.linalg.norm(b, axis=-1), 1e-300)))


bad = 0
t_start = time.perf_counter()
with nbk.Context(0) as ctx:
    for c in range(cases):
        n = int(rng.choice([3, 17, 64, 200, 513, 1025, 1400]))
        seed = int(rng.integers(1, 1000))
        eps = float(rng.choice([0.01, 0.02, 0.05]))  # softened: no pair drives the step to 1e-9
        shape = str(rng.choice(["", "", "128,1", "128,3", "256,2", "512,1", "256,5"]))
        tree_max = int(rng.choice([0, 0, 7, 33, 130, 600]))
        if shape:
            os.environ["NBK_ADT_SHAPE"] = shape
        else:
            os.environ.pop("NBK_ADT_SHAPE", None)
        try:
            m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(n, seed))
        except nbh.NbhError:  # an unbound draw: the reference's assert (ics.ml:304-331)
            continue
        # --- Advancer.A: a span short enough for a bounded number of block steps
        span = 2.0 ** -int(rng.integers(13, 18))  # (longer spans can need millions of block steps)
        s0 = nbh.AdvancerState(m, q, p)
        t0 = time.perf_counter()
        b = nbh.advancer_advance(ctx, s0, span, 1e-4, eps * eps, resident=1)
        if b.stats[0] > 60000:
            print(f"case {c}: Advancer n={n} seed={seed} skipped ({b.stats[0]} calls)")
        else:
            a = nbh.advancer_advance(ctx, s0, span, 1e-4, eps * eps, resident=2, tree_max=tree_max)
            a2 = nbh.advancer_advance(ctx, s0, span, 1e-4, eps * eps, resident=2, tree_max=tree_max)
            ok = (a.stats == b.stats and np.array_equal(a.id, b.id) and np.array_equal(a.t, b.t)
                  and np.array_equal(a.state, a2.state, equal_nan=True)
                  and rel(a.q, b.q) < 1e-9 and rel(a.p, b.p) < 1e-8)
            print(f"case {c}: Advancer n={n} seed={seed} eps={eps} span=2^{int(np.log2(span))} shape='{shape}' "
                  f"tree_max={tree_max}: {a.stats[0]} calls, dq {rel(a.q, b.q):.1e} dp {rel(a.p, b.p):.1e} "
                  f"{'ok' if ok else 'MISMATCH'} ({time.perf_counter() - t0:.2f} s)", flush=True)
            bad += 0 if ok else 1
        # --- Leapfrog
        eta = float(rng.choice([2e-3, 1e-2, 5e-2]))
        lspan = 2.0 ** -int(rng.integers(6, 10))
        ltree = int(rng.choice([0, 0, 5, 40, 300, 1024]))
        t0 = time.perf_counter()
        lb = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), lspan, eta, resident=1)
        la = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), lspan, eta, resident=2, tree_max=ltree)
        la2 = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), lspan, eta, resident=2, tree_max=ltree)
        ok = (la.nkicks == lb.nkicks and np.array_equal(la.id, lb.id) and np.array_equal(la.t, lb.t)
              and np.array_equal(la.q, la2.q) and np.array_equal(la.v, la2.v)
              and np.allclose(la.dt_max, lb.dt_max, rtol=1e-8, atol=0)
              and np.allclose(la.q, lb.q, rtol=0, atol=1e-10) and np.allclose(la.v, lb.v, rtol=0, atol=1e-9))
        print(f"case {c}: Leapfrog n={n} seed={seed} eta={eta} span=2^{int(np.log2(lspan))} tree_max={ltree}: "
              f"{la.nkicks} kick blocks {'ok' if ok else 'MISMATCH'} ({time.perf_counter() - t0:.2f} s)", flush=True)
        bad += 0 if ok else 1
        if time.perf_counter() - t_start > float(os.environ.get("STRESS_SECONDS", "40")):
            print("time budget reached")
            break
print("stress done:", "ALL OK" if bad == 0 else f"{bad} MISMATCHES")
sys.exit(1 if bad else 0)
