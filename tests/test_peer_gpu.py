"""Peer-memory exchange (include/nbk.h, nbk_dev_*_peer): the sweeps that read their source
tiles out of per-rank shards instead of one gathered array.

On a 1-GPU box the ranks are VIRTUAL: every rank's region lives on the same device, which
exercises everything but the NVLink hop (shard addressing, ready flags, biased target base,
tangent direction strides).  The tile decomposition of a rank's targets x all sources is the
same as in the gathered-array sweep, so results must be BIT-IDENTICAL to it.  With >= 2 GPUs a
torchrun job checks the real thing (CUDA IPC + NVLink) against the NCCL all-gather path."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _virtual_ranks(ctx, n, world, K=0):
    """Regions for `world` virtual ranks on ctx's device -> (plans, bases, layout)."""
    from nbk.sharded import make_plan, peer_region_layout
    plans = [make_plan(n, world, r, align=128) for r in range(world)]
    total, poff, aoff = peer_region_layout(plans[0].shard, K)
    bases = [ctx.peer_alloc(total, want_handle=False)[0] for _ in range(world)]
    return plans, bases, (total, poff, aoff)


@pytest.mark.parametrize("n,world", [(5000, 3), (1000, 8), (16384, 2), (300, 4)])
def test_virtual_ranks_bit_identical_to_gathered_sweep(ctx, oracle, n, world):
    import torch
    dev = torch.device("cuda", 0)
    m, q, p = oracle.make_plummer(n, 20243)
    eps2 = 1e-4
    plans, bases, (_, poff, _) = _virtual_ranks(ctx, n, world)
    st = torch.cuda.current_stream().cuda_stream
    d_m, d_q, d_p = (torch.as_tensor(a).to(dev).contiguous() for a in (m, q, p))
    gathered = torch.zeros(n, 8, dtype=torch.float64, device=dev)
    ctx.dev_pack(n, d_m.data_ptr(), d_q.data_ptr(), d_p.data_ptr(), False, gathered.data_ptr(), st)
    try:
        for buf, epoch in ((1, 1), (0, 2)):  # both row buffers, flags advancing
            for pl in plans:
                if pl.count:
                    ctx.dev_pack(pl.count, d_m[pl.t0:].data_ptr(), d_q[pl.t0:].data_ptr(),
                                 d_p[pl.t0:].data_ptr(), False, bases[pl.rank] + poff[buf], st)
                ctx.dev_publish(pl.rank, bases, epoch, st)
            for pl in plans:
                if not pl.count:
                    continue
                sh = ctx.make_shards(pl.rank, pl.shard, [b + poff[buf] for b in bases],
                                     bases[pl.rank], epoch)
                tg, sr = (pl.t0, pl.t1), (0, n)
                out = {k: torch.empty(pl.count, 3, dtype=torch.float64, device=dev) for k in "abcd"}
                phi, phi_ref = (torch.empty(pl.count, dtype=torch.float64, device=dev) for _ in "ab")
                ctx.dev_grad_v_dgrad_v_sum_peer(sh, n, eps2, tg, sr, out["a"].data_ptr(),
                                                out["b"].data_ptr(), False, st)
                ctx.dev_grad_v_dgrad_v_sum(gathered.data_ptr(), n, eps2, tg, sr,
                                           out["c"].data_ptr(), out["d"].data_ptr(), False, st)
                torch.cuda.synchronize()
                assert torch.equal(out["a"], out["c"]) and torch.equal(out["b"], out["d"])
                ctx.dev_grad_v_sum_peer(sh, n, eps2, tg, sr, out["a"].data_ptr(), False, st)
                ctx.dev_grad_v_sum(gathered.data_ptr(), n, eps2, tg, sr, out["c"].data_ptr(), False, st)
                ctx.dev_v_sum_peer(sh, n, eps2, tg, sr, phi.data_ptr(), False, st)
                ctx.dev_v_sum(gathered.data_ptr(), n, eps2, tg, sr, phi_ref.data_ptr(), False, st)
                torch.cuda.synchronize()
                assert torch.equal(out["a"], out["c"]) and torch.equal(phi, phi_ref)
                if epoch == 1:  # and against the oracle itself
                    g_ref, dg_ref = oracle.grad_v_dgrad_v_sum(m, q, p, eps2, targets=tg)
                    ctx.dev_grad_v_dgrad_v_sum_peer(sh, n, eps2, tg, sr, out["a"].data_ptr(),
                                                    out["b"].data_ptr(), False, st)
                    torch.cuda.synchronize()
                    err = np.linalg.norm(out["b"].cpu().numpy() - dg_ref, axis=1) / \
                        np.linalg.norm(dg_ref, axis=1)
                    assert err.max() <= 1e-12
    finally:
        torch.cuda.synchronize()
        for b in bases:
            ctx.peer_free(b)


@pytest.mark.parametrize("K", [1, 4])
def test_virtual_ranks_tangent(ctx, oracle, K):
    import torch
    dev = torch.device("cuda", 0)
    n, world, eps2 = 3000, 3, 1e-4
    m, q, p = oracle.make_plummer(n, 41)
    rng = np.random.default_rng(20245)
    dq = rng.standard_normal((K, n, 3))
    dp = rng.standard_normal((K, n, 3)) * m[None, :, None]
    plans, bases, (_, poff, aoff) = _virtual_ranks(ctx, n, world, K)
    st = torch.cuda.current_stream().cuda_stream
    d_m, d_q, d_p = (torch.as_tensor(a).to(dev).contiguous() for a in (m, q, p))
    d_dq, d_dp = torch.as_tensor(dq).to(dev).contiguous(), torch.as_tensor(dp).to(dev).contiguous()
    gathered = torch.zeros(n, 8, dtype=torch.float64, device=dev)
    gaux = torch.zeros(K, n, 8, dtype=torch.float64, device=dev)
    ctx.dev_pack(n, d_m.data_ptr(), d_q.data_ptr(), d_p.data_ptr(), False, gathered.data_ptr(), st)
    for k in range(K):
        ctx.dev_pack_aux(n, d_m.data_ptr(), d_dq[k].data_ptr(), d_dp[k].data_ptr(),
                         gaux[k].data_ptr(), st)
    try:
        shard = plans[0].shard
        for pl in plans:
            ctx.dev_pack(pl.count, d_m[pl.t0:].data_ptr(), d_q[pl.t0:].data_ptr(),
                         d_p[pl.t0:].data_ptr(), False, bases[pl.rank] + poff[0], st)
            for k in range(K):
                ctx.dev_pack_aux(pl.count, d_m[pl.t0:].data_ptr(), d_dq[k, pl.t0:].data_ptr(),
                                 d_dp[k, pl.t0:].data_ptr(),
                                 bases[pl.rank] + aoff[0] + k * shard * 64, st)
        for pl in plans:
            sh = ctx.make_shards(pl.rank, shard, [b + poff[0] for b in bases],
                                 aux=[b + aoff[0] for b in bases])  # no flags: one stream orders it
            a, b, c, d = (torch.empty(K, pl.count, 3, dtype=torch.float64, device=dev)
                          for _ in range(4))
            tg, sr = (pl.t0, pl.t1), (0, n)
            ctx.dev_grad_v_dgrad_v_sum_tangent_peer(sh, n, K, eps2, tg, sr, a.data_ptr(),
                                                    b.data_ptr(), st)
            ctx.dev_grad_v_dgrad_v_sum_tangent(gathered.data_ptr(), gaux.data_ptr(), n, K, eps2,
                                               tg, sr, c.data_ptr(), d.data_ptr(), st)
            torch.cuda.synchronize()
            assert torch.equal(a, c) and torch.equal(b, d)
    finally:
        torch.cuda.synchronize()
        for b in bases:
            ctx.peer_free(b)


def test_sharded_energy_and_kick_drivers_world1(ctx, oracle):
    """PeerShardedEnergy / PeerShardedKick (nbk/sharded.py) on one rank: the peer sweeps, the
    local reductions and the (trivial) all-reduce against the single-context calls."""
    import torch
    from nbk.sharded import PeerShardedEnergy, PeerShardedKick, make_plan
    n = 3000
    m, q, p = oracle.make_plummer(n, 20244)
    dev = torch.device("cuda", 0)
    plan = make_plan(n, 1, 0, align=128)
    e = PeerShardedEnergy(plan, dev, ctx)
    e.load_local(m, q, p)
    ke, pe = e.step(1e-4)
    ke_ref, pe_ref = ctx.energy(m, q, p, 1e-4)
    assert abs(ke - ke_ref) <= 1e-13 * abs(ke_ref) and abs(pe - pe_ref) <= 1e-13 * abs(pe_ref)
    phi_ref, _ = ctx.v_sum(m, q, 1e-4)
    assert np.array_equal(e.phi.cpu().numpy(), phi_ref)
    e.close()
    v = p / m[:, None]
    k = PeerShardedKick(plan, dev, ctx)
    k.load_local(m, q, v)
    dv, tmin, gmin = k.step(1e-3, 1e-4)
    dv_ref, tmin_ref = ctx.kick_sum(m, q, v, 1e-3, 1e-4)
    assert np.array_equal(dv.cpu().numpy(), dv_ref) and np.array_equal(tmin.cpu().numpy(), tmin_ref)
    assert gmin == tmin_ref.min()
    k.close()


def test_peer_argument_errors(ctx):
    import nbk
    base, _ = ctx.peer_alloc(1 << 20, want_handle=False)
    try:
        ok = ctx.make_shards(0, 128, [base, base + 128 * 64])
        bad_rows = ctx.make_shards(0, 100, [base, base + 128 * 64])
        with pytest.raises(nbk.NbkError):  # shard_rows not whole tiles
            ctx.dev_grad_v_sum_peer(bad_rows, 200, 0.0, (0, 100), (0, 200), base, False, 0)
        with pytest.raises(nbk.NbkError):  # targets outside this rank's shard
            ctx.dev_grad_v_sum_peer(ok, 200, 0.0, (100, 200), (0, 200), base, False, 0)
        with pytest.raises(nbk.NbkError):  # sources do not start on a tile
            ctx.dev_grad_v_sum_peer(ok, 200, 0.0, (0, 100), (5, 200), base, False, 0)
        with pytest.raises(nbk.NbkError):  # shards too small for n_total
            ctx.dev_grad_v_sum_peer(ok, 300, 0.0, (0, 100), (0, 300), base, False, 0)
    finally:
        ctx.peer_free(base)


def test_two_gpu_peer_exchange_matches_nccl_path():
    """Real NVLink + CUDA IPC: tools/run_peer_check.py under torchrun on 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "tools", "run_peer_check.py"), "--bodies", "20000", "--steps", "3"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PEER CHECK OK" in r.stdout
