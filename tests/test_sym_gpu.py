"""Pair-once (symmetric) acc+jerk sweep (csrc/sym.cuh; hermite.ml:157-166 evaluates every
unordered pair once too): against the oracle per body through the C ABI, ragged sizes and the
diagonal masks, item ranges, run-to-run identity, and the sharded form on virtual ranks (every
rank's region on one device: everything but the NVLink hop)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _rel(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.fixture
def sym_ctx(ctx):
    ctx.set_sym(2)  # every full square of >= 2048 bodies takes the pair-once kernel
    yield ctx
    ctx.set_sym(1)


@pytest.mark.parametrize("n", [2048, 2049, 3000, 4096, 5000, 7169])
@pytest.mark.parametrize("eps2", [0.0, 1e-4])
def test_pair_once_matches_oracle_per_body(sym_ctx, oracle, n, eps2):
    m, q, p = oracle.make_plummer(n, 7 + n)
    g_ref, dg_ref = oracle.grad_v_dgrad_v_sum(m, q, p, eps2)
    launches = sym_ctx.launch_count
    g, dg = sym_ctx.grad_v_dgrad_v_sum(m, q, p, eps2)
    assert _rel(g, g_ref).max() <= TOL and _rel(dg, dg_ref).max() <= TOL
    # it really was the pair-once path: pack + sym sweep + slot reduction
    assert sym_ctx.launch_count - launches == 3
    g2, dg2 = sym_ctx.grad_v_dgrad_v_sum(m, q, p, eps2)
    assert np.array_equal(g, g2) and np.array_equal(dg, dg2)  # fixed summation order
    sym_ctx.set_sym(0)
    g0, dg0 = sym_ctx.grad_v_dgrad_v_sum(m, q, p, eps2)
    sym_ctx.set_sym(2)
    assert _rel(g, g0).max() <= 1e-13 and _rel(dg, dg0).max() <= 1e-12


def test_pair_once_unequal_masses_and_sub_square(sym_ctx, oracle):
    """Masses over six decades (the j side is weighted with the TARGET's mass) and a square
    over a sub-range of the bodies (targets = sources = [a, b))."""
    rng = np.random.default_rng(5)
    n = 6000
    m, q, p = oracle.make_plummer(n, 99)
    m = m * 10.0 ** rng.uniform(-3, 3, n)
    p = p * 10.0 ** rng.uniform(-1, 1, n)[:, None]
    for rg in (None, (700, 4900)):
        g_ref, dg_ref = oracle.grad_v_dgrad_v_sum(m, q, p, 1e-6, targets=rg, sources=rg)
        g, dg = sym_ctx.grad_v_dgrad_v_sum(m, q, p, 1e-6, targets=rg, sources=rg)
        assert _rel(g, g_ref).max() <= TOL and _rel(dg, dg_ref).max() <= TOL


@pytest.mark.parametrize("n", [2048, 3000, 5000, 7169])
@pytest.mark.parametrize("eps2", [0.0, 1e-4])
def test_pair_once_acceleration_and_potential_match_oracle(sym_ctx, oracle, n, eps2):
    """The pair-once forms of Base.grad_v and Base.v summed (sym_lite_kernel; b_operate /
    c_operate of the Omelyan advancer, Energy): per body against the oracle at ragged sizes,
    unequal masses, the total potential energy, run-to-run identical."""
    rng = np.random.default_rng(n)
    m, q, p = oracle.make_plummer(n, 3 + n)
    m = m * 10.0 ** rng.uniform(-2, 2, n)
    g_ref = oracle.grad_v_sum(m, q, eps2)
    phi_ref = oracle.v_sum(m, q, eps2)
    l0 = sym_ctx.launch_count
    g = sym_ctx.grad_v_sum(m, q, eps2)
    assert sym_ctx.launch_count - l0 == 3  # pack + sym_lite_kernel + its slot reduction
    l0 = sym_ctx.launch_count
    phi, tot = sym_ctx.v_sum(m, q, eps2)
    assert sym_ctx.launch_count - l0 == 4  # pack + sym_lite_kernel + reduction + the sum
    assert _rel(g, g_ref).max() <= TOL
    assert np.max(np.abs(phi - phi_ref) / np.abs(phi_ref)) <= TOL
    pe = oracle.total_potential_energy(m, q, eps2)
    ke2, pe2 = sym_ctx.energy(m, q, p, eps2)
    assert abs(pe2 - pe) <= 1e-12 * abs(pe) and abs(0.5 * tot - pe) <= 1e-12 * abs(pe)
    g2 = sym_ctx.grad_v_sum(m, q, eps2)
    phi2, tot2 = sym_ctx.v_sum(m, q, eps2)
    assert np.array_equal(g, g2) and np.array_equal(phi, phi2) and tot == tot2
    sym_ctx.set_sym(0)  # and next to the ordered sweeps
    g0 = sym_ctx.grad_v_sum(m, q, eps2)
    phi0, tot0 = sym_ctx.v_sum(m, q, eps2)
    sym_ctx.set_sym(2)
    assert _rel(g, g0).max() <= 1e-13 and np.max(np.abs(phi - phi0) / np.abs(phi0)) <= 1e-13


def test_pair_once_acceleration_and_potential_automatic_at_size(ctx, oracle):
    """At N = 32768 the acceleration and potential calls take the pair-once kernels by
    themselves; results next to the ordered sweeps."""
    ctx.set_sym(1)
    m, q, p = oracle.make_plummer(32768, 5)
    l0 = ctx.launch_count
    g = ctx.grad_v_sum(m, q, 0.0)
    assert ctx.launch_count - l0 == 3
    t_once = ctx.last_kernel_ms()
    phi, tot = ctx.v_sum(m, q, 0.0)
    ctx.set_sym(0)
    g0 = ctx.grad_v_sum(m, q, 0.0)
    t_ord = ctx.last_kernel_ms()
    phi0, tot0 = ctx.v_sum(m, q, 0.0)
    ctx.set_sym(1)
    print(f"grad_v N=32768: pair-once {t_once:.3f} ms, ordered {t_ord:.3f} ms")
    assert _rel(g, g0).max() <= 1e-13
    assert np.max(np.abs(phi - phi0) / np.abs(phi0)) <= 1e-13 and abs(tot - tot0) <= 1e-13 * abs(tot0)


def test_pair_once_momentum_conservation(sym_ctx, oracle):
    """Pair-once sums obey Newton's third law to rounding: sum_i gsum_i = 0, sum_i dgsum_i = 0
    relative to the sum of the norms."""
    n = 8192
    m, q, p = oracle.make_plummer(n, 11)
    g, dg = sym_ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
    for a in (g, dg):
        assert np.linalg.norm(a.sum(axis=0)) <= 1e-13 * np.linalg.norm(a, axis=1).sum()


def test_automatic_choice_small_stays_ordered_large_goes_pair_once(ctx, oracle):
    ctx.set_sym(1)
    m, q, p = oracle.make_plummer(4096, 3)
    l0 = ctx.launch_count
    ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
    small = ctx.launch_count - l0
    m, q, p = oracle.make_plummer(32768, 3)
    l0 = ctx.launch_count
    g, dg = ctx.grad_v_dgrad_v_sum(m, q, p, 0.0)
    large = ctx.launch_count - l0
    assert large == 3 and small in (2, 3)
    # a rectangle (targets != sources) never takes it
    l0 = ctx.launch_count
    g1, dg1 = ctx.grad_v_dgrad_v_sum(m, q, p, 0.0, targets=(0, 16384))
    assert _rel(g1, g[:16384]).max() <= 1e-13 and _rel(dg1, dg[:16384]).max() <= 1e-12


@pytest.mark.parametrize("n,world", [(5000, 3), (8192, 2), (16384, 8), (2100, 2)])
def test_pair_once_item_ranges_on_virtual_ranks(ctx, oracle, n, world):
    """nbk_dev_grad_v_dgrad_v_sym_peer on `world` virtual ranks: rows sharded over per-rank
    regions (whole 1024-blocks), every rank runs its contiguous item range behind the ready
    flags; the per-rank sums add up to the oracle's rows for every body; the same ranges
    through the gathered-array entry give bit-identical per-rank sums."""
    import torch
    from nbk.sharded import make_plan, peer_region_layout, sym_item_range
    dev = torch.device("cuda", 0)
    eps2 = 1e-4
    m, q, p = oracle.make_plummer(n, 20243)
    plans = [make_plan(n, world, r, align=1024) for r in range(world)]
    total, poff, _ = peer_region_layout(plans[0].shard)
    bases = [ctx.peer_alloc(total, want_handle=False)[0] for _ in range(world)]
    st = torch.cuda.current_stream().cuda_stream
    d_m, d_q, d_p = (torch.as_tensor(a).to(dev).contiguous() for a in (m, q, p))
    gathered = torch.zeros(n, 8, dtype=torch.float64, device=dev)
    ctx.dev_pack(n, d_m.data_ptr(), d_q.data_ptr(), d_p.data_ptr(), False, gathered.data_ptr(), st)
    items = ctx.sym_items(n)
    nb = (n + 1023) // 1024
    assert items == nb * (nb + 1) // 2
    try:
        for pl in plans:
            if pl.count:
                ctx.dev_pack(pl.count, d_m[pl.t0:].data_ptr(), d_q[pl.t0:].data_ptr(),
                             d_p[pl.t0:].data_ptr(), False, bases[pl.rank] + poff[0], st)
            ctx.dev_publish(pl.rank, bases, 1, st)
        tot = torch.zeros(n, 6, dtype=torch.float64, device=dev)
        for pl in plans:
            sh = ctx.make_shards(pl.rank, pl.shard, [b + poff[0] for b in bases], bases[pl.rank], 1)
            rng = sym_item_range(items, world, pl.rank)
            a = torch.full((n, 6), float("nan"), dtype=torch.float64, device=dev)
            b = torch.full((n, 6), float("nan"), dtype=torch.float64, device=dev)
            ctx.dev_grad_v_dgrad_v_sym_peer(sh, n, eps2, rng, a.data_ptr(), st)
            ctx.dev_grad_v_dgrad_v_sym(gathered.data_ptr(), n, eps2, rng, b.data_ptr(), st)
            torch.cuda.synchronize()
            assert torch.equal(a, b)
            tot += a
        g_ref, dg_ref = oracle.grad_v_dgrad_v_sum(m, q, p, eps2)
        t = tot.cpu().numpy()
        assert _rel(t[:, 0:3], g_ref).max() <= TOL and _rel(t[:, 3:6], dg_ref).max() <= TOL
        # split: rows [count][6] -> gsum, dgsum
        g = torch.empty(n, 3, dtype=torch.float64, device=dev)
        dg = torch.empty(n, 3, dtype=torch.float64, device=dev)
        ctx.dev_split6(tot.data_ptr(), n, g.data_ptr(), dg.data_ptr(), st)
        torch.cuda.synchronize()
        assert torch.equal(g, tot[:, 0:3]) and torch.equal(dg, tot[:, 3:6])
    finally:
        torch.cuda.synchronize()
        for b in bases:
            ctx.peer_free(b)


@pytest.mark.parametrize("n,world", [(9000, 3), (16384, 4)])
def test_pair_once_pushed_rows_on_virtual_ranks(ctx, oracle, n, world):
    """The exchange of PeerShardedAccJerkSym on virtual ranks: every rank's packing kernel
    stores its rows into its place in EVERY rank's gathered array (nbk_dev_pack_push), publishes,
    and each rank runs its item range on its own gathered array behind the flags.  Two epochs
    with moved bodies in alternating buffers; the gathered arrays equal a plain pack of all
    bodies bit for bit and the ranks' sums add up to the oracle's rows."""
    import torch
    from nbk.sharded import make_plan, peer_region_layout, sym_item_range
    dev = torch.device("cuda", 0)
    m, q, p = oracle.make_plummer(n, 77)
    plans = [make_plan(n, world, r, align=1024) for r in range(world)]
    npad = plans[0].npad
    total, poff, _ = peer_region_layout(npad)
    bases = [ctx.peer_alloc(total, want_handle=False)[0] for _ in range(world)]
    st = torch.cuda.current_stream().cuda_stream
    items = ctx.sym_items(n)
    d_m, d_p = (torch.as_tensor(a).to(dev).contiguous() for a in (m, p))
    try:
        for epoch, buf, scale in ((1, 1, 1.0), (2, 0, 1.01)):
            qq = q * scale
            d_q = torch.as_tensor(qq).to(dev).contiguous()
            ref_rows = torch.zeros(n, 8, dtype=torch.float64, device=dev)
            ctx.dev_pack(n, d_m.data_ptr(), d_q.data_ptr(), d_p.data_ptr(), False,
                         ref_rows.data_ptr(), st)
            for pl in plans:
                dst = [b + poff[buf] + pl.rank * pl.shard * 64 for b in bases]
                ctx.dev_pack_push(pl.count, d_m[pl.t0:].data_ptr(), d_q[pl.t0:].data_ptr(),
                                  d_p[pl.t0:].data_ptr(), False, dst, st)
                ctx.dev_publish(pl.rank, bases, epoch, st)
            tot = torch.zeros(n, 6, dtype=torch.float64, device=dev)
            sums = []
            for pl in plans:
                rows = [bases[pl.rank] + poff[buf] + r * pl.shard * 64 for r in range(world)]
                sh = ctx.make_shards(pl.rank, pl.shard, rows, bases[pl.rank], epoch)
                a = torch.empty(n, 6, dtype=torch.float64, device=dev)
                sums.append(a)
                ctx.dev_grad_v_dgrad_v_sym_peer(sh, n, 0.0, sym_item_range(items, world, pl.rank),
                                                a.data_ptr(), st)
                b = torch.empty(n, 6, dtype=torch.float64, device=dev)
                ctx.dev_grad_v_dgrad_v_sym(ref_rows.data_ptr(), n, 0.0,
                                           sym_item_range(items, world, pl.rank), b.data_ptr(), st)
                torch.cuda.synchronize()
                assert torch.equal(a, b)  # the gathered rows are the plain pack, bit for bit
                tot += a
            g_ref, dg_ref = oracle.grad_v_dgrad_v_sum(m, qq, p, 0.0)
            t = tot.cpu().numpy()
            assert _rel(t[:, 0:3], g_ref).max() <= TOL and _rel(t[:, 3:6], dg_ref).max() <= TOL
            # the reduce-scatter as peer reads: second flag array, every rank gathers its slice
            # of the ranks' sums in rank order (= the order `tot` was added up in)
            flags2 = [b + 512 for b in bases]
            for pl in plans:
                ctx.dev_publish(pl.rank, flags2, epoch, st)
            for pl in plans:
                g = torch.empty(max(pl.count, 1), 3, dtype=torch.float64, device=dev)
                dg = torch.empty_like(g)
                ctx.dev_sym_gather_peer(pl.rank, [a.data_ptr() for a in sums], flags2[pl.rank],
                                        epoch, (pl.t0, pl.t1), g.data_ptr(), dg.data_ptr(), st)
                torch.cuda.synchronize()
                assert torch.equal(g[:pl.count], tot[pl.t0:pl.t1, 0:3])
                assert torch.equal(dg[:pl.count], tot[pl.t0:pl.t1, 3:6])
    finally:
        torch.cuda.synchronize()
        for b in bases:
            ctx.peer_free(b)


def test_pair_once_argument_errors(ctx):
    import nbk
    import torch
    dev = torch.device("cuda", 0)
    rows = torch.zeros(2048, 8, dtype=torch.float64, device=dev)
    out = torch.zeros(2048, 6, dtype=torch.float64, device=dev)
    with pytest.raises(nbk.NbkError):
        ctx.dev_grad_v_dgrad_v_sym(rows.data_ptr(), 2048, 0.0, (0, 4), out.data_ptr())  # 3 items
    with pytest.raises(nbk.NbkError):
        ctx.dev_grad_v_dgrad_v_sym(0, 2048, 0.0, (0, 3), out.data_ptr())
    with pytest.raises(nbk.NbkError):
        ctx.set_sym(5)
    sh = ctx.make_shards(0, 1000, [rows.data_ptr()])  # shard_rows not a multiple of 1024
    with pytest.raises(nbk.NbkError):
        ctx.dev_grad_v_dgrad_v_sym_peer(sh, 1000, 0.0, (0, 1), out.data_ptr())
    with pytest.raises(nbk.NbkError):  # a null destination
        ctx.dev_pack_push(16, rows.data_ptr(), rows.data_ptr(), rows.data_ptr(), False,
                          [rows.data_ptr(), 0])
