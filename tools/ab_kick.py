#!/usr/bin/env python
"""A/B timing of the kick with the bit-exact time-scale (NBK_LIBRARY picks the build):
full sweeps at several N and the 1/8-of-the-targets share of an 8-GPU run (short runs per CTA)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402

with nbk.Context(0) as ctx:
    for n, tg in ((16384, None), (32768, None), (131072, None), (131072, (0, 16384)), (131072, (0, 2048))):
        m, q, p = nbh.make_plummer(n, 20243)
        v = p / m[:, None]
        best = 1e9
        for _ in range(3):
            dv, tm = ctx.kick_sum(m, q, v, 1e-3, 1e-3, targets=tg)
            best = min(best, ctx.last_kernel_ms())
        nt = n if tg is None else tg[1] - tg[0]
        print(f"kickt n={n} targets={nt}: {best:.3f} ms, {nt * (n - 1.0) / best / 1e6:.1f} G pairs/s, "
              f"checksum {float(np.nansum(tm)):.17g}")
