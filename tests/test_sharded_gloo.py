"""CPU, world_size 2 over gloo: the multi-GPU plumbing of nbk/sharded.py (shard plan, in-place
all-gather layout, overlap path's source split and accumulation).  The CUDA sweep is replaced by
an injected checker built on the oracle - test infrastructure standing in for the kernel, so
that the host-side logic can be verified without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_plan_covers_every_body_once():
    from nbk.sharded import make_plan
    for n in (0, 1, 7, 8, 1000, 131072, 131073):
        for world in (1, 2, 3, 4, 8):
            plans = [make_plan(n, world, r) for r in range(world)]
            assert plans[0].t0 == 0 and plans[-1].t1 == n
            for a, b in zip(plans, plans[1:]):
                assert a.t1 == b.t0 and a.shard == b.shard
            assert all(p.count <= p.shard for p in plans) and plans[0].npad >= n
    with pytest.raises(ValueError):
        make_plan(10, 2, 2)


def _oracle_sweep(packed, n, eps2, targets, sources, g, dg, accumulate):
    from oracle import oracle as O
    a = packed[:n].numpy()
    m, q, v = a[:, 3].copy(), a[:, 0:3].copy(), a[:, 4:7].copy()
    gs, dgs = O.grad_v_dgrad_v_sum(m, q, v * m[:, None], eps2, targets=targets, sources=sources)
    if accumulate:
        g += torch.from_numpy(gs)
        dg += torch.from_numpy(dgs)
    else:
        g.copy_(torch.from_numpy(gs))
        dg.copy_(torch.from_numpy(dgs))


def _worker(rank, world, port, n, overlap, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbk.sharded import ShardedAccJerk, make_plan
        from oracle import oracle as O
        m, q, p = O.make_plummer(n, 20243)
        # momenta whose p/m round-trips exactly, so the checker sees the same velocities
        v = p / m[:, None]
        plan = make_plan(n, world, rank)
        sh = ShardedAccJerk(plan, torch.device("cpu"), overlap=overlap, sweep=_oracle_sweep)
        sh.load_local(m, q, v * m[:, None])
        g, dg = sh.step(0.0)
        np.save(os.path.join(out_dir, f"g{rank}.npy"), g[:plan.count].numpy())
        np.save(os.path.join(out_dir, f"dg{rank}.npy"), dg[:plan.count].numpy())
        np.save(os.path.join(out_dir, f"all{rank}.npy"), sh.packed_all.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,overlap", [(1000, False), (1001, True), (64, True)])
def test_two_rank_sharded_sweep_equals_single_process(tmp_path, oracle, n, overlap):
    world = 2
    port = 29600 + (os.getpid() + n) % 300
    mp.spawn(_worker, args=(world, port, n, overlap, str(tmp_path)), nprocs=world, join=True)
    m, q, p = oracle.make_plummer(n, 20243)
    v = p / m[:, None]
    g_ref, dg_ref = oracle.grad_v_dgrad_v_sum(m, q, v * m[:, None], 0.0)
    g = np.concatenate([np.load(tmp_path / f"g{r}.npy") for r in range(world)])
    dg = np.concatenate([np.load(tmp_path / f"dg{r}.npy") for r in range(world)])
    if overlap:  # sources are summed in three pieces: rounding-level differences only
        for a, b in ((g, g_ref), (dg, dg_ref)):
            assert np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)) < 1e-13
    else:
        assert np.array_equal(g, g_ref) and np.array_equal(dg, dg_ref)
    # every rank ends up with the same gathered array: row i is body i
    a0, a1 = np.load(tmp_path / "all0.npy"), np.load(tmp_path / "all1.npy")
    assert np.array_equal(a0, a1)
    assert np.array_equal(a0[:n, 0:3], q) and np.array_equal(a0[:n, 3], m)


# ---- peer-memory exchange: host-side plumbing (the CUDA side is tests/test_peer_gpu.py) ------
def test_aligned_plan_keeps_tiles_inside_one_owner():
    from nbk.sharded import make_plan
    for n in (1, 127, 128, 129, 5000, 131072, 131073):
        for world in (1, 2, 3, 8):
            plans = [make_plan(n, world, r, align=128) for r in range(world)]
            assert plans[0].shard % 128 == 0 and plans[0].npad >= n
            assert plans[0].t0 == 0 and plans[-1].t1 == n
            for a, b in zip(plans, plans[1:]):
                assert a.t1 == b.t0
            for j0 in range(0, n, 128):  # every source tile has exactly one owner
                assert j0 // plans[0].shard == (min(j0 + 128, n) - 1) // plans[0].shard


def test_peer_region_layout_is_disjoint_and_aligned():
    from nbk.sharded import PEER_FLAG_BYTES, peer_region_layout
    for rows, K in ((128, 0), (16384, 0), (4096, 6)):
        total, poff, aoff = peer_region_layout(rows, K)
        assert poff[0] == PEER_FLAG_BYTES and len(poff) == 2
        row_bytes = rows * 64
        assert poff[1] - poff[0] == row_bytes * (1 + K) and total == poff[1] + row_bytes * (1 + K)
        assert all(o % 128 == 0 for o in poff + aoff)
        assert aoff[0] == poff[0] + row_bytes


def test_row_buffer_toggle_follows_loads_not_epoch_parity():
    """load, step, step, load must not pack into the buffer the peers' sweeps of the
    latest epoch read.  The toggle is relative to the buffer in use; two loads in a row re-pack
    the same (unpublished) buffer."""
    from nbk.sharded import next_row_buffer
    cur, stepped, live = 0, True, []
    for op in "LSSLSLLSSSLS":
        if op == "L":
            new = next_row_buffer(cur, stepped)
            if stepped:          # `cur` has been published: a peer may be reading it
                assert new != cur
            else:                # nobody has seen `cur` yet
                assert new == cur
            cur, stepped = new, False
        else:
            stepped = True
        live.append(cur)
    assert live == [1, 1, 1, 0, 0, 1, 1, 1, 1, 1, 0, 0]


class _FakeCtx:
    """Stands in for nbk.Context: records the region calls PeerRegions makes."""

    def __init__(self, rank):
        self.rank, self.log = rank, []

    def peer_alloc(self, nbytes, want_handle=True):
        self.log.append(("alloc", nbytes))
        return 0x1000 * (self.rank + 1), (bytes([self.rank]) * 64 if want_handle else None)

    def peer_open(self, handle):
        self.log.append(("open", handle[0]))
        return 0x100000 * (handle[0] + 1)

    def peer_close(self, ptr):
        self.log.append(("close", ptr))

    def peer_free(self, ptr):
        self.log.append(("free", ptr))


def _peer_worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbk.sharded import PeerRegions, make_plan
        ctx = _FakeCtx(rank)
        rg = PeerRegions(make_plan(1000, world, rank, align=128), ctx)  # handles over gloo
        bases = list(rg.base)
        rg.close(barrier=dist.barrier)
        np.save(os.path.join(out_dir, f"bases{rank}.npy"), np.array(bases))
        with open(os.path.join(out_dir, f"log{rank}.txt"), "w") as f:
            f.write(repr(ctx.log))
    finally:
        dist.destroy_process_group()


def test_two_rank_peer_region_handle_exchange(tmp_path):
    world = 2
    port = 29300 + os.getpid() % 300
    mp.spawn(_peer_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        bases = np.load(tmp_path / f"bases{rank}.npy")
        other = 1 - rank
        assert bases[rank] == 0x1000 * (rank + 1)          # own allocation
        assert bases[other] == 0x100000 * (other + 1)      # the peer's handle, opened
        log = eval(open(tmp_path / f"log{rank}.txt").read())
        kinds = [k for k, _ in log]
        assert kinds == ["alloc", "open", "close", "free"]  # unmap peers before freeing
        assert log[1] == ("open", other)


class _FailingCtx(_FakeCtx):
    def peer_open(self, handle):
        if self.rank == 1:
            raise RuntimeError("cudaIpcOpenMemHandle failed (simulated)")
        return super().peer_open(handle)


def _peer_fail_worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbk.sharded import PeerRegions, make_plan
        ctx = _FailingCtx(rank)
        try:
            PeerRegions(make_plan(1000, world, rank, align=128), ctx)
            outcome = "mapped"
        except RuntimeError as e:
            outcome = "error: " + str(e)
        with open(os.path.join(out_dir, f"fail{rank}.txt"), "w") as f:
            f.write(outcome + "\n" + repr([k for k, _ in ctx.log]))
    finally:
        dist.destroy_process_group()


def test_peer_region_failure_is_agreed_by_every_rank(tmp_path):
    """One rank cannot map its peer: BOTH ranks get the RuntimeError (nobody is left waiting in
    a collective), unmap what they mapped and free their own region - bench.py then falls back
    to the NCCL all-gather on every rank."""
    world = 2
    port = 29350 + os.getpid() % 200
    mp.spawn(_peer_fail_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        outcome, kinds = open(tmp_path / f"fail{rank}.txt").read().split("\n")
        assert outcome.startswith("error: peer-memory exchange unavailable") and "rank 1" in outcome
        kinds = eval(kinds)
        assert kinds[0] == "alloc" and kinds[-1] == "free"
        assert kinds.count("open") == kinds.count("close")


# ---- pair-once sharded sweep: item ranges, reduce-scatter, split (CUDA side: test_sym_gpu.py) --


class _FakeSymCtx(_FakeCtx):
    def dev_publish(self, *a):
        self.log.append(("publish",))

    def make_shards(self, *a):
        return None


def _sym_worker(rank, world, port, n, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nbk.sharded import PeerShardedAccJerkSym, make_plan, sym_item_blocks, SYM_BLOCK
        from oracle import oracle as O
        m, q, p = O.make_plummer(n, 20243)
        plan = make_plan(n, world, rank, align=SYM_BLOCK)

        def sweep(sh, n_, eps2, item_range, sum6):
            # what nbk_dev_grad_v_dgrad_v_sym_peer leaves: the sums restricted to the pairs of
            # the rank's block pairs, for all bodies
            sum6.zero_()
            B = SYM_BLOCK
            for i, j in sym_item_blocks(n_, *item_range):
                ti, tj = (i * B, min((i + 1) * B, n_)), (j * B, min((j + 1) * B, n_))
                g, dg = O.grad_v_dgrad_v_sum(m, q, p, eps2, targets=ti, sources=tj)
                sum6[ti[0]:ti[1], 0:3] += torch.from_numpy(g)
                sum6[ti[0]:ti[1], 3:6] += torch.from_numpy(dg)
                if i != j:
                    g, dg = O.grad_v_dgrad_v_sum(m, q, p, eps2, targets=tj, sources=ti)
                    sum6[tj[0]:tj[1], 0:3] += torch.from_numpy(g)
                    sum6[tj[0]:tj[1], 3:6] += torch.from_numpy(dg)

        def reduce_scatter(out, inp):  # gloo has no reduce_scatter_tensor
            tot = inp.clone()
            dist.all_reduce(tot)
            out.copy_(tot[rank * plan.shard:(rank + 1) * plan.shard])

        def split(src, count, g, dg):
            g[:count] = src[:count, 0:3]
            dg[:count] = src[:count, 3:6]

        sh = PeerShardedAccJerkSym(plan, torch.device("cpu"), _FakeSymCtx(rank), sweep=sweep,
                                   reduce_scatter=reduce_scatter, split=split)
        g, dg = sh.step(0.0)
        np.save(os.path.join(out_dir, f"g{rank}.npy"), g[:plan.count].numpy())
        np.save(os.path.join(out_dir, f"dg{rank}.npy"), dg[:plan.count].numpy())
        np.save(os.path.join(out_dir, f"items{rank}.npy"), np.array(sh.item_range))
        sh.regions.close(barrier=dist.barrier)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [4096, 5000])
def test_two_rank_pair_once_plumbing(tmp_path, oracle, n):
    """PeerShardedAccJerkSym: contiguous item ranges that cover the upper triangle once, the
    reduce-scatter of the per-rank sums and the split give every rank its slice of the full
    acc+jerk sums (to summation order)."""
    from nbk.sharded import make_plan
    world = 2
    port = 29100 + (os.getpid() + n) % 300
    mp.spawn(_sym_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    m, q, p = oracle.make_plummer(n, 20243)
    g_ref, dg_ref = oracle.grad_v_dgrad_v_sum(m, q, p, 0.0)
    items = [np.load(tmp_path / f"items{r}.npy") for r in range(world)]
    nb = (n + 1023) // 1024
    assert items[0][0] == 0 and items[0][1] == items[1][0] and items[1][1] == nb * (nb + 1) // 2
    g = np.concatenate([np.load(tmp_path / f"g{r}.npy") for r in range(world)])
    dg = np.concatenate([np.load(tmp_path / f"dg{r}.npy") for r in range(world)])
    assert g.shape == (n, 3)
    for a, b in ((g, g_ref), (dg, dg_ref)):
        assert np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)) < 1e-13
    plans = [make_plan(n, world, r, align=1024) for r in range(world)]
    assert plans[0].shard % 1024 == 0 and plans[0].count + plans[1].count == n
