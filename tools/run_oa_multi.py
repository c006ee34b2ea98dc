#!/usr/bin/env python
"""BASELINE configs[3] from ONE process over several GPUs: two-mass Plummer N = 65536,
OA.advance (4 all-pairs grad_v sweeps per step) + energy + per-body E_i each step, through the
host mirror on an nbk_create_multi context - the path a single OCaml process would take."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m, q, p = nbh.make_plummer(n, 20244)
m, q, p = nbh.add_mass_disc(m, q, p, 20244)
m, q, p = nbh.adjust_frame(m, q, p)
ref = None
g = 1
while g <= torch.cuda.device_count():
    with nbk.Context(devices=list(range(g))) as ctx:
        mm, qq, pp = nbh.rescale_to_standard_units(ctx, m, q, p)
        e0 = sum(nbh.energy(ctx, mm, qq, pp))
        nbh.omelyan_advance(ctx, mm, qq, pp, 0.0, 1.0 / 256)
        t, ts, td = 0.0, 0.0, 0.0
        for _ in range(4):
            t0 = time.perf_counter()
            qq, pp, t = nbh.omelyan_advance(ctx, mm, qq, pp, t, 1.0 / 256)
            ts += time.perf_counter() - t0
            t0 = time.perf_counter()
            e = sum(nbh.energy(ctx, mm, qq, pp))
            el = nbh.bodies_to_el_phase_space(ctx, mm, qq, pp)
            td += time.perf_counter() - t0
        if g == 1:  # the same four steps with the bodies resident on the device (nbk_oa_*)
            m1, q1, p1 = nbh.rescale_to_standard_units(ctx, m, q, p)
            t1 = 0.0
            nbh.omelyan_advance_steps(ctx, m1, q1, p1, 0.0, 1.0 / 256, 1, want_energies=False)  # warm-up
            t0 = time.perf_counter()
            q2, p2, t2, _, _ = nbh.omelyan_advance_steps(ctx, m1, q1, p1, t1, 1.0 / 256, 4, want_energies=False)
            tr = (time.perf_counter() - t0) / 4
            t0 = time.perf_counter()
            q3, p3, t3, en, el3 = nbh.omelyan_advance_steps(ctx, m1, q1, p1, t1, 1.0 / 256, 4, want_el=True)
            trd = (time.perf_counter() - t0) / 4
            same = bool(np.array_equal(q2, qq) and np.array_equal(p2, pp))
            print(f"1 GPU, bodies resident on the device: OA step {1e3 * tr:.2f} ms (3 sweeps in the steady "
                  f"state), step + energy diagnostics {1e3 * trd:.2f} ms, bit-identical to the "
                  f"host-buffer steps: {same}, dE/E {abs((en[-1] - e0) / e0):.3e}")
    if ref is None:
        ref = qq
    diff = float(np.max(np.linalg.norm(qq - ref, axis=1) / np.linalg.norm(ref, axis=1)))
    print(f"{g} GPU(s), one process: OA step {1e3 * ts / 4:.2f} ms, diagnostics {1e3 * td / 4:.2f} ms, "
          f"dE/E after 4 steps {abs((e - e0) / e0):.3e}, max rel position diff vs 1 GPU {diff:.1e}")
    g *= 2
