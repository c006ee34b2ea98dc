#!/usr/bin/env python
"""Small-size tour of the pair-once kernels (csrc/sym.cuh) for compute-sanitizer:
  compute-sanitizer --tool memcheck  python tools/sanitize_sym.py
  compute-sanitizer --tool racecheck python tools/sanitize_sym.py
Touches the plain and the masked (diagonal, ragged tail) items of the acc+jerk, acceleration and
potential kernels, their slot reductions, an item range with sharded rows behind ready flags,
the pushed row gather and the peer gather of the sums (virtual ranks on one device)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import torch  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402
from nbk.sharded import make_plan, peer_region_layout, sym_item_range  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
m, q, p = nbh.make_plummer(n, 5)
dev = torch.device("cuda", 0)
with nbk.Context(0) as ctx:
    ctx.set_sym(2)
    ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
    ctx.grad_v_sum(m, q, 1e-4)
    ctx.v_sum(m, q, 0.0)
    world = 2
    plans = [make_plan(n, world, r, align=1024) for r in range(world)]
    total, poff, _ = peer_region_layout(plans[0].npad, extra_bytes=plans[0].npad * 48)
    extra = total - plans[0].npad * 48
    bases = [ctx.peer_alloc(total, want_handle=False)[0] for _ in range(world)]
    st = torch.cuda.current_stream().cuda_stream
    d_m, d_q, d_p = (torch.as_tensor(a).to(dev).contiguous() for a in (m, q, p))
    items = ctx.sym_items(n)
    for pl in plans:
        dst = [b + poff[0] + pl.rank * pl.shard * 64 for b in bases]
        ctx.dev_pack_push(pl.count, d_m[pl.t0:].data_ptr(), d_q[pl.t0:].data_ptr(),
                          d_p[pl.t0:].data_ptr(), False, dst, st)
        ctx.dev_publish(pl.rank, bases, 1, st)
    flags2 = [b + 512 for b in bases]
    for pl in plans:
        rows = [bases[pl.rank] + poff[0] + r * pl.shard * 64 for r in range(world)]
        sh = ctx.make_shards(pl.rank, pl.shard, rows, bases[pl.rank], 1)
        ctx.dev_grad_v_dgrad_v_sym_peer(sh, n, 0.0, sym_item_range(items, world, pl.rank),
                                        bases[pl.rank] + extra, st)
        ctx.dev_publish(pl.rank, flags2, 1, st)
    for pl in plans:
        g = torch.empty(pl.count, 3, dtype=torch.float64, device=dev)
        dg = torch.empty_like(g)
        ctx.dev_sym_gather_peer(pl.rank, [b + extra for b in bases], flags2[pl.rank], 1,
                                (pl.t0, pl.t1), g.data_ptr(), dg.data_ptr(), st)
    torch.cuda.synchronize()
    for b in bases:
        ctx.peer_free(b)
print("sanitize_sym tour done")
