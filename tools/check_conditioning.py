#!/usr/bin/env python
"""Worst-case parity at full size: the bodies whose acc+jerk sums differ most between two GPU
summation orders (tile shapes) are the ill-conditioned ones; for those, compare GPU, the oracle
(reference summation order) and binary128 truth."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
m, q, p = nbh.make_plummer(n, 20243)
with nbk.Context(0) as ctx:
    a = ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
    ctx.tune(2)
    b = ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)


def rel(x, y):
    return np.linalg.norm(x - y, axis=1) / np.linalg.norm(y, axis=1)


d = np.maximum(rel(a[0], b[0]), rel(a[1], b[1]))
worst = np.argsort(d)[-12:][::-1]
print("body   gpu-vs-gpu   | acc: gpu-vs-truth  oracle-vs-truth  gpu-vs-oracle | jerk: gpu-vs-truth  oracle-vs-truth  gpu-vs-oracle")
mx = np.zeros(6)
for i in worst:
    i = int(i)
    og, odg = O.grad_v_dgrad_v_sum(m, q, p, 0.0, targets=(i, i + 1))
    qg, qdg = O.grad_v_dgrad_v_sum(m, q, p, 0.0, targets=(i, i + 1), quad=True)
    r = [rel(a[0][i:i + 1], qg)[0], rel(og, qg)[0], rel(a[0][i:i + 1], og)[0],
         rel(a[1][i:i + 1], qdg)[0], rel(odg, qdg)[0], rel(a[1][i:i + 1], odg)[0]]
    mx = np.maximum(mx, r)
    print(f"{i:6d}  {d[i]:.2e}   |  {r[0]:.2e}  {r[1]:.2e}  {r[2]:.2e}  |  {r[3]:.2e}  {r[4]:.2e}  {r[5]:.2e}")
print("max over these:", " ".join(f"{x:.2e}" for x in mx))
