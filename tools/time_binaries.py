import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'direct-nbody_b200')]
import numpy as np, nbh, nbk
with nbk.Context(0) as ctx:
    for n in (16384, 65536, 131072):
        m, q, p = nbh.make_plummer(n, 20243)
        ctx.binary_scan(m, q, p, 0.0, 0.0)
        t = time.perf_counter(); e, pa, c = ctx.binary_scan(m, q, p, 0.0, 0.0); dt = time.perf_counter() - t
        print(f"binary_scan n={n}: {1e3*dt:.2f} ms wall, kernel {ctx.last_kernel_ms():.2f} ms, {n*(n-1.0)/ctx.last_kernel_ms()/1e6:.1f} G pairs/s, bound pairs/2={int(c.sum())//2}")
