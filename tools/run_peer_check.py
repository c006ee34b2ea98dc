#!/usr/bin/env python
"""torchrun job: the peer-memory exchange (sweep reads remote source tiles over NVLink, CUDA IPC
regions) against the NCCL all-gather path, same shards -> results must be bit-identical.
Also times both (device events, max over ranks) and, with --tangent K, the tangent sweeps.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
      tools/run_peer_check.py --bodies 131072 --steps 10
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "direct-nbody_b200")):
    sys.path.insert(0, _p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bodies", dest="n", type=int, default=131072)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--tangent", type=int, default=0)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    import nbh
    import nbk
    from nbk.sharded import (PeerShardedAccJerk, PeerShardedTangent, ShardedAccJerk,
                             ShardedTangent, make_plan)

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = nbk.Context(local)
    n = args.n
    m, q, p = nbh.make_plummer(n, 20243)
    plan = make_plan(n, world, rank, align=128)
    stream = torch.cuda.current_stream()

    def timed(step):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ok = True
    if args.tangent == 0:
        a = ShardedAccJerk(plan, dev, ctx=ctx)
        b = PeerShardedAccJerk(plan, dev, ctx)
        a.load_local(m, q, p)
        b.load_local(m, q, p)
        ga, dga = a.step(1e-4)
        gb, dgb = b.step(1e-4)
        torch.cuda.synchronize()
        ok = torch.equal(ga, gb) and torch.equal(dga, dgb)
        # a second epoch with fresh rows in the other buffer
        q2 = q * 1.001
        a.load_local(m, q2, p)
        b.load_local(m, q2, p)
        ga, dga = a.step(0.0)
        gb, dgb = b.step(0.0)
        torch.cuda.synchronize()
        ok = ok and torch.equal(ga, gb) and torch.equal(dga, dgb)
        # load, step, step, load, step with ranks deliberately skewed: the second load must not
        # touch the row buffer a slower peer's sweep of the latest epoch is still reading
        for rep in range(4):
            if rank == rep % world:
                torch.cuda._sleep(int(3e7))  # this rank's stream lags ~15 ms
            qa, qb = q * (1.0 + 1e-3 * (rep + 1)), q * (1.0 - 1e-3 * (rep + 1))
            b.load_local(m, qa, p)
            b.step(0.0)
            gb, dgb = b.step(0.0)
            gb1, dgb1 = gb.clone(), dgb.clone()
            b.load_local(m, qb, p)
            gb, dgb = b.step(0.0)
            a.load_local(m, qa, p)
            ga, dga = a.step(0.0)
            ok = ok and torch.equal(ga, gb1) and torch.equal(dga, dgb1)
            a.load_local(m, qb, p)
            ga, dga = a.step(0.0)
            torch.cuda.synchronize()
            ok = ok and torch.equal(ga, gb) and torch.equal(dga, dgb)
        ms_nccl = timed(lambda: a.step(0.0))
        ms_peer = timed(lambda: b.step(0.0))
        inter = float(n) * (n - 1)
        # pair-once sweep over the box (block pairs dealt out to the ranks, reduce-scatter of
        # the sums): against the ordered sharded sweep per body, with skewed ranks and rows
        # alternating between the two buffers; run-to-run identical
        from nbk.sharded import PeerShardedAccJerkSym
        plan_s = make_plan(n, world, rank, align=1024)
        a2 = ShardedAccJerk(plan_s, dev, ctx=ctx)
        c = PeerShardedAccJerkSym(plan_s, dev, ctx)
        sym_rel = 0.0
        for rep in range(4):
            if rank == rep % world:
                torch.cuda._sleep(int(3e7))
            qa = q * (1.0 + 1e-3 * (rep + 1))
            c.load_local(m, qa, p)
            gc, dgc = c.step(0.0)
            gc1, dgc1 = gc.clone(), dgc.clone()
            gc, dgc = c.step(0.0)  # same rows again: bit-identical
            torch.cuda.synchronize()
            ok = ok and torch.equal(gc, gc1) and torch.equal(dgc, dgc1)
            a2.load_local(m, qa, p)
            ga, dga = a2.step(0.0)
            torch.cuda.synchronize()
            cnt = plan_s.count
            for x, y in ((gc, ga), (dgc, dga)):
                r = (torch.linalg.norm(x[:cnt] - y[:cnt], dim=1) / torch.linalg.norm(y[:cnt], dim=1)).max()
                sym_rel = max(sym_rel, float(r))
        ok = ok and sym_rel <= 1e-12
        ms_sym = timed(lambda: c.step(0.0))
        # the same sweep with an NCCL reduce-scatter of the sums instead of the ranks' peer reads
        c2 = PeerShardedAccJerkSym(plan_s, dev, ctx, reduce="nccl")
        c2.load_local(m, qa, p)
        g2, dg2 = c2.step(0.0)
        torch.cuda.synchronize()
        cnt = plan_s.count
        for x, y in ((g2, gc), (dg2, dgc)):
            r = (torch.linalg.norm(x[:cnt] - y[:cnt], dim=1) / torch.linalg.norm(y[:cnt], dim=1)).max()
            ok = ok and float(r) <= 1e-13
        ms_sym_nccl = timed(lambda: c2.step(0.0))
        if rank == 0:
            print(f"pair-once sharded sweep N={n}: {ms_sym:.3f} ms/step = {inter / ms_sym / 1e6:.1f} G interactions/s "
                  f"(sums by peer reads; {ms_sym_nccl:.3f} ms with an NCCL reduce-scatter), "
                  f"max per-body difference to the ordered sweep {sym_rel:.2e}")
        c.close()
        c2.close()
    else:
        K = args.tangent
        rng = np.random.default_rng(20245)
        dq = rng.standard_normal((K, n, 3))
        dp = rng.standard_normal((K, n, 3)) * m[None, :, None]
        a = ShardedTangent(plan, K, dev, ctx)
        b = PeerShardedTangent(plan, K, dev, ctx)
        a.load_local(m, q, p, dq, dp)
        b.load_local(m, q, p, dq, dp)
        ga, dga = a.step(1e-4)
        gb, dgb = b.step(1e-4)
        torch.cuda.synchronize()
        ok = torch.equal(ga, gb) and torch.equal(dga, dgb)
        ms_nccl = timed(lambda: a.step(1e-4))
        ms_peer = timed(lambda: b.step(1e-4))
        inter = float(K) * n * (n - 1)
    if args.tangent == 0:
        # energy (one all-reduce of two doubles) and kick + global min time-scale over the same
        # exchange, against this rank's own single-GPU evaluation of the whole system
        from nbk.sharded import PeerShardedEnergy, PeerShardedKick
        nn = min(n, 16384)
        pl2 = make_plan(nn, world, rank, align=128)
        mm, qq, pp = m[:nn], q[:nn], p[:nn]
        e = PeerShardedEnergy(pl2, dev, ctx)
        e.load_local(mm, qq, pp)
        ke, pe = e.step(1e-4)
        ke_ref, pe_ref = ctx.energy(mm, qq, pp, 1e-4)
        ok = ok and abs(ke - ke_ref) <= 1e-12 * abs(ke_ref) and abs(pe - pe_ref) <= 1e-12 * abs(pe_ref)
        e.close()
        vv = pp / mm[:, None]
        k = PeerShardedKick(pl2, dev, ctx)
        k.load_local(mm, qq, vv)
        dv, tmin, gmin = k.step(1e-3, 1e-4)
        dv_ref, tmin_ref = ctx.kick_sum(mm, qq, vv, 1e-3, 1e-4)
        torch.cuda.synchronize()
        sl = slice(pl2.t0, pl2.t1)
        dv_h = dv[:pl2.count].cpu().numpy()
        rel = np.max(np.linalg.norm(dv_h - dv_ref[sl], axis=1) / np.linalg.norm(dv_ref[sl], axis=1)) if pl2.count else 0.0
        ok = ok and rel <= 1e-12 and np.array_equal(tmin[:pl2.count].cpu().numpy(), tmin_ref[sl])
        ok = ok and gmin == float(tmin_ref.min())
        k.close()
        if rank == 0:
            print(f"sharded energy N={nn}: ke {ke:.15e} pe {pe:.15e} (single GPU {ke_ref:.15e} {pe_ref:.15e}); "
                  f"kick: max rel dv diff {rel:.1e}, tmin bit-exact, global min {gmin:.6e}")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    b.close()
    if rank == 0:
        what = "acc+jerk" if args.tangent == 0 else f"tangent K={args.tangent}"
        print(f"{what} N={n} on {world} GPUs: NCCL all-gather + sweep {ms_nccl:.3f} ms/step "
              f"({inter / ms_nccl / 1e6:.1f} G/s), peer-read sweep {ms_peer:.3f} ms/step "
              f"({inter / ms_peer / 1e6:.1f} G/s); bit-identical: {bool(flag.item())}")
        print("PEER CHECK OK" if flag.item() else "PEER CHECK FAILED")
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
