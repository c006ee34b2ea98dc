#!/usr/bin/env python
"""One GPU: the PEER sweep kernel with every shard local against the plain kernel, N = 131072 -
separates the cost of the PEER variant's code from the cost of the NVLink hop."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import torch  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402
from nbk.sharded import make_plan, peer_region_layout  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
ctx = nbk.Context(0)
m, q, p = nbh.make_plummer(n, 20243)
st = torch.cuda.current_stream().cuda_stream
d_m, d_q, d_p = (torch.as_tensor(a).to(dev).contiguous() for a in (m, q, p))
gathered = torch.zeros(n, 8, dtype=torch.float64, device=dev)
ctx.dev_pack(n, d_m.data_ptr(), d_q.data_ptr(), d_p.data_ptr(), False, gathered.data_ptr(), st)
plans = [make_plan(n, world, r, align=128) for r in range(world)]
total, poff, _ = peer_region_layout(plans[0].shard)
bases = [ctx.peer_alloc(total, want_handle=False)[0] for _ in range(world)]
for pl in plans:
    ctx.dev_pack(pl.count, d_m[pl.t0:].data_ptr(), d_q[pl.t0:].data_ptr(), d_p[pl.t0:].data_ptr(),
                 False, bases[pl.rank] + poff[0], st)
pl = plans[0]
g = torch.empty(pl.count, 3, dtype=torch.float64, device=dev)
dg = torch.empty_like(g)
sh = ctx.make_shards(0, pl.shard, [b + poff[0] for b in bases])


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tg, sr = (pl.t0, pl.t1), (0, n)


def plain():
    ctx.dev_grad_v_dgrad_v_sum(gathered.data_ptr(), n, 0.0, tg, sr, g.data_ptr(), dg.data_ptr(),
                               False, st)


def peer():
    ctx.dev_grad_v_dgrad_v_sum_peer(sh, n, 0.0, tg, sr, g.data_ptr(), dg.data_ptr(), False, st)


res = [(timed(plain), timed(peer)) for _ in range(3)]  # A B A B A B: no order effect
a, b = min(r[0] for r in res), min(r[1] for r in res)
print(f"N={n}, targets of rank 0 of {world}: plain kernel {a:.4f} ms, PEER kernel (local shards) "
      f"{b:.4f} ms ({100 * (b / a - 1):+.2f} %)   all: {res}")
