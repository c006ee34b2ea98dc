"""GPU: the host mirror of the reference's integrators (libnbh: Hermite, Leapfrog, OA,
Advancer.A, Energy, Integrator - every pair loop on the GPU through the C ABI) against the
oracle's restatement of the same OCaml functions, on the same seeded bodies.

Sum order differs from the reference (1e-12 per sweep) and an N-body flow amplifies that, so
states are compared with tolerances that grow with the integration length; energy-error
trajectories - what north_star asks to agree - are compared directly, and the reference's own
test assertions (test/advance_test.ml, test/leapfrog_test.ml) are re-run on the GPU path."""
import numpy as np
import pytest

import nbh

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)))


def by_id(ids, *arrays):
    o = np.argsort(ids, kind="stable")
    return [a[o] for a in arrays]


@pytest.fixture(scope="module")
def plummer64(oracle):
    return oracle.rescale_to_standard_units(*oracle.make_plummer(64, 20241))


def test_energy_and_el_phase_space(ctx, oracle):
    m, q, p = oracle.make_plummer(1000, 3)
    ke, pe = nbh.energy(ctx, m, q, p, 1e-4)
    assert abs(ke - oracle.total_kinetic_energy(m, p)) <= 1e-13 * ke
    assert abs(pe - oracle.total_potential_energy(m, q, 1e-4)) <= 1e-12 * abs(pe)
    el = nbh.bodies_to_el_phase_space(ctx, m, q, p, 1e-4)
    el_ref = oracle.bodies_to_el_phase_space(m, q, p, 1e-4)
    assert np.allclose(el, el_ref, rtol=1e-11, atol=0)


def test_rescale_to_standard_units_on_gpu(ctx, oracle):
    """ics.ml:304-331 with Energy.energy on the GPU; test/ic_test.ml:47-61's assertion."""
    m, q, p = nbh.make_uniform_box(1000, 42)
    m2, q2, p2 = nbh.rescale_to_standard_units(ctx, m, q, p)
    r = oracle.rescale_to_standard_units(m, q, p)
    assert np.allclose(m2, r[0], rtol=1e-15) and np.allclose(q2, r[1], rtol=1e-12, atol=1e-15)
    ke, pe = nbh.energy(ctx, m2, q2, p2)
    assert abs(ke + pe + 0.25) <= 1e-8 and abs(m2.sum() - 1.0) <= 1e-8


def test_omelyan_step_matches_oracle(ctx, oracle, plummer64):
    m, q, p = plummer64
    q1, p1, t1 = nbh.omelyan_advance(ctx, m, q, p, 0.0, 1.0 / 64, 0.01)
    qr, pr, tr = oracle.omelyan_advance(m, q, p, 0.0, 1.0 / 64, 0.01)
    assert t1 == tr and rel(q1, qr) <= 1e-13 and rel(p1, pr) <= 1e-12
    # eps = 0 and a few steps
    qa, pa, ta = q, p, 0.0
    qb, pb, tb = q, p, 0.0
    for _ in range(8):
        qa, pa, ta = nbh.omelyan_advance(ctx, m, qa, pa, ta, 1.0 / 256)
        qb, pb, tb = oracle.omelyan_advance(m, qb, pb, tb, 1.0 / 256)
    assert ta == tb and rel(qa, qb) <= 1e-11 and rel(pa, pb) <= 1e-10


def test_omelyan_two_mass_energy_diagnostics(ctx, oracle):
    """BASELINE configs[3] in miniature: two-mass Plummer (el_mass_segregation ICs), OA steps,
    per-step energy + per-body E_i on the GPU; energy error follows the oracle's."""
    n = 512
    m, q, p = nbh.make_plummer(n, 20244)
    m, q, p = nbh.add_mass_disc(m, q, p, 20244)
    m, q, p = nbh.adjust_frame(m, q, p)
    m, q, p = nbh.rescale_to_standard_units(ctx, m, q, p, 1e-4)
    e0 = sum(nbh.energy(ctx, m, q, p, 1e-4))
    qa, pa, ta, qb, pb, tb = q, p, 0.0, q, p, 0.0
    for _ in range(4):
        qa, pa, ta = nbh.omelyan_advance(ctx, m, qa, pa, ta, 1.0 / 256, 1e-4)
        qb, pb, tb = oracle.omelyan_advance(m, qb, pb, tb, 1.0 / 256, 1e-4)
        ea = sum(nbh.energy(ctx, m, qa, pa, 1e-4))
        eb = oracle.energy(m, qb, pb, 1e-4)
        assert abs(ea - eb) <= 1e-11 * abs(eb)
        el = nbh.bodies_to_el_phase_space(ctx, m, qa, pa, 1e-4)
        assert abs(el[:, 0].sum() - (ea + sum(nbh.energy(ctx, m, qa, pa, 1e-4)[1:]))) <= 1e-9
    assert abs((ea - e0) / e0) < 1e-5


@pytest.mark.parametrize("n,eps", [(1, 0.0), (2, 0.0), (300, 0.0), (2500, 0.01)])
def test_omelyan_resident_steps_bit_identical_to_host_buffer_steps(ctx, oracle, n, eps):
    """nbk_oa_* (bodies, operators and sweeps resident on the device, a step's first b_operate
    reusing the previous step's last sums) against stepping through nbk_grad_v_sum from the
    host: identical bits, energies and per-body E_i from one potential sweep included."""
    m, q, p = oracle.make_plummer(n, 20244)
    eps2, h, nsteps = eps * eps, 1.0 / 128, 3
    qa, pa, ta, en_ref = q, p, 0.0, []
    for _ in range(nsteps):
        qa, pa, ta = nbh.omelyan_advance(ctx, m, qa, pa, ta, h, eps2)
        en_ref.append(sum(nbh.energy(ctx, m, qa, pa, eps2)))
    el_ref = nbh.bodies_to_el_phase_space(ctx, m, qa, pa, eps2)
    qb, pb, tb, en, el = nbh.omelyan_advance_steps(ctx, m, q, p, 0.0, h, nsteps, eps2, want_el=True)
    assert tb == ta and np.array_equal(qb, qa) and np.array_equal(pb, pa)
    assert np.array_equal(en, np.array(en_ref)) and np.array_equal(el, el_ref)
    if n > 1:  # and the oracle, to rounding
        qo, po, to = q, p, 0.0
        for _ in range(nsteps):
            qo, po, to = oracle.omelyan_advance(m, qo, po, to, h, eps2)
        assert rel(qb, qo) <= 1e-12 and rel(pb, po) <= 1e-11
        assert abs(en[-1] - oracle.energy(m, qo, po, eps2)) <= 1e-11 * abs(en[-1])
    # a second call continues from a fresh upload: no stale sums survive
    qc, pc, tc, _, _ = nbh.omelyan_advance_steps(ctx, m, qb, pb, tb, h, 1, eps2)
    qd, pd, td = nbh.omelyan_advance(ctx, m, qa, pa, ta, h, eps2)
    assert np.array_equal(qc, qd) and np.array_equal(pc, pd)


def _hot100(oracle, seed):
    m, q, p = oracle.make_hot_spherical(100, seed)
    m, q, p = oracle.adjust_frame(m, q, p)
    return oracle.rescale_to_standard_units(m, q, p)


@pytest.mark.parametrize("seed", [1, 2])
def test_leapfrog_const_dt_matches_oracle_and_reference_test(ctx, oracle, seed):
    """test/leapfrog_test.ml:7-15 on the GPU path; with eta = infinity the pre-kick order is
    exact, so states agree with the oracle to rounding (including the reference's sub-step
    ladder on fresh bodies)."""
    m, q, p = _hot100(oracle, seed)
    de = []
    for dt in (1e-4, 5e-5):
        a = nbh.leapfrog_advance_const_dt(ctx, nbh.LeapfrogState(m, q, p), dt, 1)
        b = oracle.leapfrog_advance_const_dt(oracle.LeapfrogState(m, q, p), dt, 1)
        assert np.array_equal(a.id, b.id) and np.allclose(a.t, b.t, rtol=0, atol=1e-18)
        assert rel(a.q, b.q) <= 1e-13 and rel(a.v, b.v) <= 1e-12
        ke, pe = nbh.energy(ctx, a.m, a.q, a.p)
        de.append(abs((ke + pe + 0.25) / 0.25))
    assert np.sqrt(32.0) < de[0] / de[1] < np.sqrt(128.0)


def test_leapfrog_const_dt_device_resident_steps(ctx, oracle):
    """Steady-state const-dt steps on the device (nbk_resident_leapfrog_steps) against the host
    recursion with one sweep per step, and against the oracle."""
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(777, 20241))
    s = nbh.LeapfrogState(m, q, p)
    s.dt_max[:] = np.inf
    a = nbh.leapfrog_advance_const_dt(ctx, s, 1.0 / 512, 12, resident=True)
    b = nbh.leapfrog_advance_const_dt(ctx, s, 1.0 / 512, 12, resident=False)
    so = oracle.LeapfrogState(m, q, p)
    so.dt_max[:] = np.inf
    c = oracle.leapfrog_advance_const_dt(so, 1.0 / 512, 12)
    assert np.array_equal(a.t, b.t) and np.array_equal(a.t, c.t) and np.all(np.isinf(a.dt_max))
    assert np.array_equal(a.q, b.q) and np.array_equal(a.v, b.v)   # same kernels, same order
    assert rel(a.q, c.q) <= 1e-13 and rel(a.v, c.v) <= 1e-11


def test_leapfrog_initial_timescales_bit_exact(ctx, oracle):
    """Leapfrog.advance's dt = 0 sweep (leapfrog.ml:151-155): dt_max from the GPU equals the
    sequential reference loop bit for bit, hence the same sorted order."""
    m, q, p = _hot100(oracle, 3)
    a = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), 0.0, 1e-3)
    b = oracle.leapfrog_advance(oracle.LeapfrogState(m, q, p), 0.0, 1e-3)
    assert np.array_equal(a.id, b.id)


@pytest.mark.parametrize("n,span,eta", [(100, 1.0 / 32, 1e-3), (1500, 1.0 / 64, 2e-3), (1, 0.1, 1e-3),
                                        (64, 0.0, 1e-3)])
def test_leapfrog_adaptive_device_resident_is_bit_identical_to_host_path(ctx, oracle, n, span, eta):
    """Leapfrog.advance with the bodies resident on the device (nbk_lf_*: drifts, kick blocks,
    dt_max updates and sort moves on the GPU, block recursion on the host) against the
    host-buffer path (one nbk_kick_block_sum per block): same order, same bits."""
    m, q, p = (oracle.rescale_to_standard_units(*oracle.make_hot_spherical(n, 5)) if n > 1
               else oracle.make_hot_spherical(n, 5))
    a = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span, eta, resident=1)
    b = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span, eta, resident=0)
    assert a.nkicks == b.nkicks and np.array_equal(a.id, b.id)
    assert np.array_equal(a.t, b.t) and np.array_equal(a.dt_max, b.dt_max)
    assert np.array_equal(a.q, b.q) and np.array_equal(a.v, b.v)
    if n > 1:
        assert a.nkicks > 1 or span == 0.0


@pytest.mark.parametrize("n,span,eta,tree_max", [(100, 1.0 / 32, 1e-3, 0), (512, 1.0 / 64, 2e-3, 0),
                                                 (1024, 1.0 / 128, 2e-3, 1024), (1500, 1.0 / 64, 2e-3, 0),
                                                 (1500, 1.0 / 64, 2e-3, 100), (1, 0.1, 1e-3, 0),
                                                 (64, 0.0, 1e-3, 0), (700, 1.0 / 256, float("inf"), 1024)])
def test_leapfrog_block_recursion_on_device(ctx, oracle, n, span, eta, tree_max):
    """nbk_lf_advance_subsystem: advance_subsystem (leapfrog.ml:119-142) in ONE launch for every
    subtree whose prefix fits (one CTA, bodies in shared memory).  Same kick blocks, same order
    and times as the host recursion; dt_max and states to rounding (the velocity sums are
    ordered differently; the pair time-scales themselves are bit-exact); reproducible."""
    m, q, p = (oracle.rescale_to_standard_units(*oracle.make_hot_spherical(n, 5)) if n > 1
               else oracle.make_hot_spherical(n, 5))
    a = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span, eta, resident=2, tree_max=tree_max)
    b = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span, eta, resident=1)
    assert a.nkicks == b.nkicks and np.array_equal(a.id, b.id)
    assert np.array_equal(a.t, b.t)
    assert np.allclose(a.dt_max, b.dt_max, rtol=1e-9, atol=0)
    assert np.allclose(a.q, b.q, rtol=0, atol=1e-11) and np.allclose(a.v, b.v, rtol=0, atol=1e-10)
    a2 = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span, eta, resident=2, tree_max=tree_max)
    assert np.array_equal(a.q, a2.q) and np.array_equal(a.v, a2.v) and np.array_equal(a.dt_max, a2.dt_max)


@pytest.mark.parametrize("n,iend", [(1, 1), (300, 300), (5000, 4097), (20000, 20000)])
def test_leapfrog_sort_sub_on_device_is_ocamls_stable_sort(ctx, n, iend):
    """nbk_lf_sort_sub against Array.fast_sort (stable) with OCaml's polymorphic compare on floats
    (leapfrog.ml:106-111): nan below everything and equal to itself, -0.0 = 0.0, +inf on top, many
    ties (dt_max = infinity is what a freshly reset block holds)."""
    rng = np.random.default_rng(n + iend)
    key = rng.choice([np.inf, -np.inf, np.nan, 0.0, -0.0, 1.0, 2.0 ** -30, -3.5], size=n)
    rnd = rng.random(n) < 0.5
    key[rnd] = 10.0 ** rng.uniform(-12, 3, int(rnd.sum()))
    m = rng.uniform(0.5, 2.0, n)
    q, v = rng.standard_normal((n, 3)), rng.standard_normal((n, 3))
    ctx.lf_upload(key, m, q, v)
    perm = ctx.lf_sort_sub(iend)
    # the reference order: stable sort with cmp(x, y) = nan-lowest total order
    import functools

    def ocaml_compare(i, j):
        x, y = key[i], key[j]
        if x < y:
            return -1
        if x > y:
            return 1
        if x == y:
            return 0
        xn, yn = np.isnan(x), np.isnan(y)
        return 0 if (xn and yn) else (-1 if xn else 1)
    want = sorted(range(iend), key=functools.cmp_to_key(ocaml_compare))  # sorted() is stable
    assert list(perm) == want
    dt2, q2, v2 = ctx.lf_download(n)
    full = np.concatenate([perm, np.arange(iend, n)])
    assert np.array_equal(dt2.view(np.uint64), key[full].view(np.uint64))   # rows beyond iend untouched
    assert np.array_equal(q2, q[full]) and np.array_equal(v2, v[full])


def test_leapfrog_adaptive_energy_error_beside_oracle(ctx, oracle):
    """Leapfrog.advance with block time steps (leapfrog.ml:119-158) on the GPU path beside the
    oracle's sequential-kick run of the same bodies: same elapsed time, energy errors of the same
    size (statistical agreement - dt_max sees pre-kick velocities), and the
    reference test's scaling with eta (test/leapfrog_test.ml:17-25) over a shortened interval
    (the full unit-time run makes ~10^6 tiny kick sweeps; it is latency-, not kernel-bound)."""
    m, q, p = _hot100(oracle, 2)
    out = []
    for eta in (1e-3, 5e-4):
        a = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), 1.0 / 32, eta)
        b = oracle.leapfrog_advance(oracle.LeapfrogState(m, q, p), 1.0 / 32, eta)
        assert np.allclose(a.t, 1.0 / 32, rtol=0, atol=1e-14) and np.allclose(b.t, a.t, atol=1e-14)
        ke, pe = nbh.energy(ctx, a.m, a.q, a.p)
        da = abs((ke + pe + 0.25) / 0.25)
        db = abs((oracle.energy(b.m, b.q, b.p) + 0.25) / 0.25)
        assert 0.2 < da / db < 5.0, (da, db)
        qa, = by_id(a.id, a.q)
        qb, = by_id(b.id, b.q)
        assert rel(qa, qb) < 1e-6
        out.append(da)
    assert 2.0 < out[0] / out[1] < 8.0


def test_leapfrog_unit_time_energy_error_trajectory_beside_oracle(ctx, oracle):
    """north_star: leapfrog energy-error trajectories agree over ONE N-body time unit.
    test/leapfrog_test.ml:17-25 at its own length - hot sphere N = 100, Leapfrog.advance bs 1.0 eta,
    eta in {1e-3, 5e-4} - on the GPU path (device-side advance_subsystem recursion) beside the
    oracle's sequential-kick run of the same bodies: both stop at t = 1, the GPU path's energy
    errors sit within 2x of the oracle's, its eta-scaling ratio falls in the reference test's
    band (sqrt 8, sqrt 32), and the error trajectory sampled every quarter time unit
    (four successive advances of 1/4) stays within 2x of the oracle's at every checkpoint."""
    m, q, p = _hot100(oracle, 2)
    de_gpu, de_ref = [], []
    for eta in (1e-3, 5e-4):
        a = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), 1.0, eta)
        b = oracle.leapfrog_advance(oracle.LeapfrogState(m, q, p), 1.0, eta)
        assert np.allclose(a.t, 1.0, rtol=0, atol=1e-12) and np.allclose(b.t, 1.0, rtol=0, atol=1e-12)
        ke, pe = nbh.energy(ctx, a.m, a.q, a.p)
        de_gpu.append(abs((ke + pe + 0.25) / 0.25))
        de_ref.append(abs((oracle.energy(b.m, b.q, b.p) + 0.25) / 0.25))
        print(f"leapfrog N=100 t=1 eta={eta:g}: dE gpu {de_gpu[-1]:.3e} oracle {de_ref[-1]:.3e} "
              f"kick blocks {a.nkicks} / {b.nkicks}")
        assert 0.5 < de_gpu[-1] / de_ref[-1] < 2.0, (eta, de_gpu, de_ref)
    r_gpu, r_ref = de_gpu[0] / de_gpu[1], de_ref[0] / de_ref[1]
    assert np.sqrt(8.0) < r_ref < np.sqrt(32.0), r_ref
    assert np.sqrt(8.0) < r_gpu < np.sqrt(32.0), r_gpu
    # the trajectory: |E(t) + 1/4| / (1/4) at t = 1/4, 1/2, 3/4, 1 from successive advances
    sa, sb = nbh.LeapfrogState(m, q, p), oracle.LeapfrogState(m, q, p)
    for k in range(4):
        sa = nbh.leapfrog_advance(ctx, sa, 0.25, 1e-3)
        sb = oracle.leapfrog_advance(sb, 0.25, 1e-3)
        ke, pe = nbh.energy(ctx, sa.m, sa.q, sa.p)
        da = abs((ke + pe + 0.25) / 0.25)
        db = abs((oracle.energy(sb.m, sb.q, sb.p) + 0.25) / 0.25)
        print(f"  t = {0.25 * (k + 1):.2f}: dE gpu {da:.3e} oracle {db:.3e}")
        assert np.allclose(sa.t, 0.25 * (k + 1), rtol=0, atol=1e-12)
        assert 0.5 < da / db < 2.0, (k, da, db)


@pytest.mark.parametrize("mode", [1, 2])
def test_hermite_short_run_matches_oracle(ctx, oracle, plummer64, mode):
    """mode 1: host scheduler + one launch per advance_body; mode 2: the whole loop in one
    persistent kernel (hermite_resident_kernel)."""
    m, q, p = plummer64
    a = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), 0.0625, 1e-5, mode=mode)
    b = oracle.hermite_advance(oracle.HermiteState(m, q, p), 0.0625, 1e-5)
    assert np.array_equal(a.id, b.id) and np.array_equal(a.t, b.t)
    assert abs(a.nsteps - b.nsteps) <= max(2, b.nsteps // 1000)
    assert rel(a.q, b.q) <= 1e-9 and rel(a.p, b.p) <= 1e-8


def test_hermite_energy_error_trajectory_agrees_with_oracle(ctx, oracle):
    """north_star: Hermite energy-error trajectories agree - N = 64 over 1 N-body time unit,
    checked at 4 points (BASELINE configs[0] at N = 1024 is the next test)."""
    eta = 1e-4
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(64, 20240))
    e0 = oracle.energy(m, q, p)
    sa, sb = nbh.HermiteState(m, q, p), oracle.HermiteState(m, q, p)
    for _ in range(4):
        sa = nbh.hermite_advance(ctx, sa, 0.25, eta)
        sb = oracle.hermite_advance(sb, 0.25, eta)
        ea = abs((sum(nbh.energy(ctx, sa.m, sa.q, sa.p)) - e0) / e0)
        eb = abs((oracle.energy(sb.m, sb.q, sb.p) - e0) / e0)
        assert ea < 1e-5
        assert 0.1 < (ea + 1e-13) / (eb + 1e-13) < 10.0, (ea, eb)
        assert abs(sa.nsteps - sb.nsteps) <= max(3, sb.nsteps // 200)
    assert np.all(sa.t == 1.0)


def test_hermite_tight_tolerance_single_advance(ctx, oracle):
    """eta = 1e-6 in ONE advance call.  (Chaining calls at this eta breaks the reference's own
    algorithm: finish_bs leaves a body with a tiny step whose corrector-predictor distance is
    rounding noise, and timescale (hermite.ml:87-93) then shrinks h without bound - the oracle
    loops forever on it too; both sides carry a step cap for that reason.)"""
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(64, 20240))
    e0 = oracle.energy(m, q, p)
    a = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), 0.25, 1e-6)
    b = oracle.hermite_advance(oracle.HermiteState(m, q, p), 0.25, 1e-6)
    ea = abs((sum(nbh.energy(ctx, a.m, a.q, a.p)) - e0) / e0)
    eb = abs((oracle.energy(b.m, b.q, b.p) - e0) / e0)
    assert ea < 1e-9 and eb < 1e-9 and abs(a.nsteps - b.nsteps) <= 20
    assert rel(a.q, b.q) <= 1e-9


def test_hermite_config0_n1024_energy_error(ctx, oracle):
    """BASELINE configs[0]: Plummer N = 1024 equal-mass, 4th-order Hermite for 1 N-body time
    unit, energy error against the oracle's run of the same bodies (device-resident loop)."""
    import time
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(1024, 20240))
    e0 = oracle.energy(m, q, p)
    t0 = time.perf_counter()
    a = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), 1.0, 1e-4)
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    b = oracle.hermite_advance(oracle.HermiteState(m, q, p), 1.0, 1e-4)
    t_cpu = time.perf_counter() - t0
    ea = abs((sum(nbh.energy(ctx, a.m, a.q, a.p)) - e0) / e0)
    eb = abs((oracle.energy(b.m, b.q, b.p) - e0) / e0)
    print(f"config0: steps gpu {a.nsteps} cpu {b.nsteps}; dE/E gpu {ea:.3e} cpu {eb:.3e}; "
          f"wall gpu {t_gpu:.2f}s cpu {t_cpu:.2f}s")
    # At eta = 1e-4 and eps = 0 the reference's own step criterion loses a close encounter
    # (dE/E ~ 1.7 in the oracle): what is asserted is that the GPU run reproduces the
    # reference's error, not that the error is small.
    assert np.all(a.t == 1.0)
    assert abs(a.nsteps - b.nsteps) <= 0.02 * b.nsteps
    assert 0.5 < (ea + 1e-12) / (eb + 1e-12) < 2.0


@pytest.mark.parametrize("seed", [5, 6, 7])
def test_advancer_energy_error_reference_test(ctx, oracle, seed):
    """test/advance_test.ml:9-31 on the GPU path, and the final state beside the oracle's."""
    n = 10
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, seed))
    eps2 = (4.0 / n) ** 2
    e0 = oracle.energy(m, q, p, eps2)
    de = []
    for sf in (1e-4, 1e-3):
        a = nbh.advancer_advance(ctx, nbh.AdvancerState(m, q, p), 1.0, sf, eps2)
        b = oracle.advancer_advance(oracle.AdvancerState(m, q, p), 1.0, sf, eps2)
        assert np.all(a.t == 1.0) and a.stats == b.stats
        qa, pa = by_id(a.id, a.q, a.p)
        qb, pb = by_id(b.id, b.q, b.p)
        assert rel(qa, qb) <= 1e-9 and rel(pa, pb) <= 1e-9
        de.append(abs(sum(nbh.energy(ctx, a.m, a.q, a.p, eps2)) - e0))
    ratio = de[1] / de[0] / 10.0 ** (4.0 / 3.0)
    assert np.sqrt(0.1) < ratio < np.sqrt(10.0)
    assert de[0] <= 1e-6 * abs(e0)


def test_advancer_external_potential_reference_test(ctx, oracle):
    """test/advance_test.ml:34-52 on the GPU path."""
    n = 100
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 1))
    eps2 = (4.0 / n) ** 2
    e0 = oracle.energy(m, q, p, eps2) + nbh.extpot_plummer_test_energy(m, q)
    a = nbh.advancer_advance(ctx, nbh.AdvancerState(m, q, p), 1.0, 1e-2, eps2, extpot="plummer_test")
    e1 = sum(nbh.energy(ctx, a.m, a.q, a.p, eps2)) + nbh.extpot_plummer_test_energy(a.m, a.q)
    assert abs((e1 - e0) / e0) < 1e-2
    b = oracle.advancer_advance(oracle.AdvancerState(m, q, p), 1.0, 1e-2, eps2, extpot="plummer_test")
    e1b = oracle.energy(b.m, b.q, b.p, eps2) + oracle.extpot_plummer_test_energy(b.m, b.q)
    assert abs(e1 - e1b) <= 1e-6 * abs(e1b)


def test_integrator_checkpoints_and_eg_err(ctx, oracle):
    """Integrator.constant_sf_integrate (integrator.ml:58-84): steps of ci, energy check after
    each, Eg_err with the last good state when the bound is exceeded."""
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(32, 9))
    s0 = nbh.AdvancerState(m, q, p)
    s1, errs = nbh.constant_sf_integrate(ctx, s0, 0.5, ci=0.125, sf=1e-3, max_eg_err=1e-2, eps2=0.01,
                                         resident=1)
    assert len(errs) == 4 and np.all(s1.t == 0.5) and errs.max() < 1e-4
    with pytest.raises(nbh.EgErr) as ei:
        nbh.constant_sf_integrate(ctx, s0, 0.5, ci=0.125, sf=1e-3, max_eg_err=1e-16, eps2=0.01,
                                  resident=1)
    assert np.all(ei.value.state.t == 0.0) and len(ei.value.errors) == 1
    # records resident + energy from the resident records == host-buffer path, bit for bit
    s2, errs2 = nbh.constant_sf_integrate(ctx, s0, 0.5, ci=0.125, sf=1e-3, max_eg_err=1e-2, eps2=0.01,
                                          resident=False)
    assert np.array_equal(errs, errs2) and np.array_equal(s1.id, s2.id)
    assert np.array_equal(s1.state, s2.state, equal_nan=True)
    with pytest.raises(nbh.EgErr) as ej:
        nbh.constant_sf_integrate(ctx, s0, 0.5, ci=0.125, sf=1e-3, max_eg_err=1e-16, eps2=0.01,
                                  resident=False)
    assert np.array_equal(ej.value.errors, ei.value.errors)
    # default: the block recursion on the device too - same checkpoints, to rounding
    s3, errs3 = nbh.constant_sf_integrate(ctx, s0, 0.5, ci=0.125, sf=1e-3, max_eg_err=1e-2, eps2=0.01)
    assert len(errs3) == 4 and np.all(s3.t == 0.5) and np.array_equal(s3.id, s1.id)
    assert np.allclose(errs3, errs, rtol=1e-4, atol=1e-15)
    assert rel(s3.q, s1.q) <= 1e-9 and rel(s3.p, s1.p) <= 1e-9


def test_hermite_cluster_kernel_matches_host_scheduler_and_oracle(ctx, oracle):
    """1536 < N <= 12288: the stepping loop runs on an 8-CTA thread-block cluster with the state
    split over the CTAs' shared memories (DSMEM).  Same steps as the launch-per-step path and
    the oracle."""
    n = 2000
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 20246))
    a = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), 0.004, 1e-4, mode=2)   # cluster
    b = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), 0.004, 1e-4, mode=1)   # host scheduler
    c = oracle.hermite_advance(oracle.HermiteState(m, q, p), 0.004, 1e-4)
    assert np.all(a.t == 0.004) and np.array_equal(a.id, c.id)
    assert abs(a.nsteps - c.nsteps) <= 2 and abs(b.nsteps - c.nsteps) <= 2
    assert rel(a.q, c.q) <= 1e-10 and rel(a.p, c.p) <= 1e-8
    assert rel(a.q, b.q) <= 1e-10 and rel(a.p, b.p) <= 1e-8


def _mirrored(oracle, half, seed):
    """A point-symmetric system: body i + half is body i reflected through the origin, so the two
    always have the same next time and the schedule is decided by the tie rule alone."""
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(half, seed))
    return np.concatenate([m, m]) * 0.5, np.concatenate([q, -q]), np.concatenate([p, -p]) * 0.5


@pytest.mark.parametrize("n,threads", [(64, "0"), (100, "128"), (333, "0"), (777, "192"), (1024, "0"),
                                       (1024, "512"), (1536, "0"), (1300, "1024"),
                                       (1537, "0"), (3001, "0"), (6000, "0")])  # the last three: cluster
def test_hermite_pipelined_loop_matches_plain_loop_and_oracle(oracle, monkeypatch, n, threads):
    """The single-CTA stepping loop is software-pipelined (corrector of step k beside the search
    and the force sum of step k+1, integer-key arg-min); NBK_HERMITE_SPEC=0 is the plain loop.
    Same schedule - same step count as each other and as the oracle - and the same state."""
    import nbk
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 20250 + n))
    span = 0.005 if n > 1536 else (0.01 if n > 512 else 0.03)
    if threads != "0":
        monkeypatch.setenv("NBK_HERMITE_THREADS", threads)
    with nbk.Context(0) as c1:
        a = nbh.hermite_advance(c1, nbh.HermiteState(m, q, p), span, 1e-4, mode=2)
        a2 = nbh.hermite_advance(c1, a, span, 1e-4, mode=2)          # resumed
    monkeypatch.setenv("NBK_HERMITE_SPEC", "0")
    with nbk.Context(0) as c0:
        b = nbh.hermite_advance(c0, nbh.HermiteState(m, q, p), span, 1e-4, mode=2)
        b2 = nbh.hermite_advance(c0, b, span, 1e-4, mode=2)
    c = oracle.hermite_advance(oracle.HermiteState(m, q, p), span, 1e-4)
    assert np.all(a.t == span) and np.array_equal(a.id, c.id)
    assert abs(a.nsteps - b.nsteps) <= 2 and abs(a.nsteps - c.nsteps) <= 2 and a.nsteps >= n
    # the six sums are added in another order: rounding differences, amplified by close encounters
    assert rel(a.q, b.q) <= 1e-10 and rel(a.p, b.p) <= 1e-8 and rel(a.f, b.f) <= 1e-7
    assert rel(a.q, c.q) <= 1e-10 and rel(a.p, c.p) <= 1e-8
    assert abs(a2.nsteps - b2.nsteps) <= 2 and rel(a2.q, b2.q) <= 1e-9 and np.all(a2.t == 2 * span)


@pytest.mark.parametrize("spec", ["1", "0"])
def test_hermite_device_loops_break_ties_like_the_reference(oracle, monkeypatch, spec):
    """Every body starts with a twin that has exactly the same next time; which one steps first is
    the reference's list order (stable sort, List.merge puts the re-inserted body first,
    hermite.ml:178,187).  The exact ties last only until rounding separates the twins (the sums
    run over differently ordered lists), after which their order is decided at the 1e-16 level and
    costs ~1e-10 in the positions - so: identical step counts on every path, states to 1e-8."""
    import nbk
    m, q, p = _mirrored(oracle, 96, 20251)
    monkeypatch.setenv("NBK_HERMITE_SPEC", spec)
    with nbk.Context(0) as cx:
        a = nbh.hermite_advance(cx, nbh.HermiteState(m, q, p), 0.05, 1e-4, mode=2)
        h = nbh.hermite_advance(cx, nbh.HermiteState(m, q, p), 0.05, 1e-4, mode=1)  # host scheduler
    c = oracle.hermite_advance(oracle.HermiteState(m, q, p), 0.05, 1e-4)
    assert a.nsteps == c.nsteps == h.nsteps and a.nsteps > 400
    assert rel(a.q, c.q) <= 1e-8 and rel(a.p, c.p) <= 1e-7
    assert rel(h.q, c.q) <= 1e-8 and rel(a.q, h.q) <= 1e-8
    assert np.abs(a.q[:96] + a.q[96:]).max() <= 1e-7 * np.abs(a.q).max()   # still twins


@pytest.mark.parametrize("n,ctas", [(2000, "0"), (2000, "3"), (300, "0"), (7, "0")])
def test_hermite_grid_kernel_matches_cluster_kernel_and_oracle(oracle, monkeypatch, n, ctas):
    """N > 12288 steps on the whole grid (cooperative launch: state split over the SMs' shared
    memories, candidates and partial sums through L2, two grid barriers per step).  Forced here at
    small N (NBK_HERMITE_THREADS=-148) so that the oracle and the other device loops can run
    beside it: same steps, same state."""
    import nbk
    monkeypatch.setenv("NBK_HERMITE_THREADS", "-148")
    if ctas != "0":
        monkeypatch.setenv("NBK_HERMITE_GRID_CTAS", ctas)
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 20246))
    span = 0.004 if n > 100 else 0.05
    with nbk.Context(0) as grid_ctx:
        a = nbh.hermite_advance(grid_ctx, nbh.HermiteState(m, q, p), span, 1e-4, mode=2)
    monkeypatch.delenv("NBK_HERMITE_THREADS")
    with nbk.Context(0) as plain_ctx:
        b = nbh.hermite_advance(plain_ctx, nbh.HermiteState(m, q, p), span, 1e-4, mode=2)
    c = oracle.hermite_advance(oracle.HermiteState(m, q, p), span, 1e-4)
    assert np.all(a.t == span) and np.array_equal(a.id, c.id)
    assert abs(a.nsteps - c.nsteps) <= 2 and a.nsteps == b.nsteps
    assert rel(a.q, c.q) <= 1e-10 and rel(a.p, c.p) <= 1e-8
    assert rel(a.q, b.q) <= 1e-11 and rel(a.p, b.p) <= 1e-9


def test_hermite_grid_kernel_large_system_beside_host_scheduler(ctx, oracle):
    """N = 20000 (> 12288: the grid kernel is the automatic choice) against the host scheduler's
    one-launch-per-step path over the same short span, and the oracle's step count."""
    n = 20000
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 20247))
    span = 2.0 ** -12
    a = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), span, 1e-3, mode=0)   # grid kernel
    b = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), span, 1e-3, mode=1)   # host scheduler
    assert np.all(a.t == span) and np.all(b.t == span)
    assert a.nsteps >= n and abs(a.nsteps - b.nsteps) <= 2
    assert rel(a.q, b.q) <= 1e-11 and rel(a.p, b.p) <= 1e-9
    ke, pe = nbh.energy(ctx, a.m, a.q, a.p)
    assert abs((ke + pe + 0.25) / 0.25) < 1e-6


@pytest.mark.parametrize("n,span,eps", [(10, 1.0, 0.4), (100, 0.25, 0.04), (1024, 1.0 / 64, 0.01)])
def test_advancer_device_resident_is_bit_identical_to_host_path(ctx, oracle, n, span, eps):
    """Advancer.A with the body records resident on the device (nbk_adv_*: predictors,
    interact_bodies, finish_body and the sort's row moves on the GPU, the block recursion on the
    host) against the host-buffer path: identical schedule, identical bits; and both beside the
    oracle."""
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 5))
    eps2 = eps * eps
    a = nbh.advancer_advance(ctx, nbh.AdvancerState(m, q, p), span, 1e-4, eps2, resident=1)
    b = nbh.advancer_advance(ctx, nbh.AdvancerState(m, q, p), span, 1e-4, eps2, resident=0)
    c = oracle.advancer_advance(oracle.AdvancerState(m, q, p), span, 1e-4, eps2)
    assert a.stats == b.stats == c.stats
    assert np.array_equal(a.id, b.id) and np.array_equal(a.id, c.id)
    assert np.array_equal(a.state, b.state, equal_nan=True)
    assert np.all(a.t == span)
    assert rel(a.q, c.q) <= 1e-9 and rel(a.p, c.p) <= 1e-9
    # a second advance continues from the carried state (h, hmax, dVdq* all live)
    a2 = nbh.advancer_advance(ctx, a, span, 1e-4, eps2, resident=1)
    b2 = nbh.advancer_advance(ctx, b, span, 1e-4, eps2, resident=0)
    assert np.array_equal(a2.state, b2.state, equal_nan=True) and a2.stats == b2.stats


def assert_records_close(a, b, tol=1e-9):
    """Two Advancer.A record arrays that followed the same block schedule with differently
    ordered sums: times and step sizes equal, vectors to rounding amplified by the flow; hmax
    comes out of |q - q2| (a cancellation, advancer.ml:273) and is only compared loosely."""
    assert np.array_equal(a.id, b.id)
    sa, sb = a.state, b.state
    for col in (0, 1, 8, 10):  # t, m, h, t0
        assert np.array_equal(sa[:, col], sb[:, col]), col
    assert np.allclose(sa[:, 9], sb[:, 9], rtol=1e-3, atol=0)  # hmax
    for c0 in range(2, 8, 3):  # q, p
        assert rel(sa[:, c0:c0 + 3], sb[:, c0:c0 + 3]) <= tol, c0
    for c0 in list(range(11, 35, 3)):  # dVdq0-2, p0, q0, q1, q2, cs
        d = np.max(np.abs(sa[:, c0:c0 + 3] - sb[:, c0:c0 + 3]))
        assert d <= 100 * tol * np.max(np.abs(sb[:, c0:c0 + 3])), (c0, d)


@pytest.mark.parametrize("n,span,eps,tree_max", [(10, 1.0, 0.4, 0), (100, 0.25, 0.04, 0),
                                                 (1024, 1.0 / 64, 0.01, 0), (700, 1.0 / 32, 0.01, 96),
                                                 (3000, 2.0 ** -16, 0.01, 512)])
def test_advancer_block_recursion_on_device(ctx, oracle, n, span, eps, tree_max):
    """nbk_adv_advance_range: advance_range (advancer.ml:316-350) as ONE persistent cooperative
    kernel for every subtree whose prefix fits (all of it when n <= 1024; below a host-side top
    of the recursion otherwise / when tree_max forces it).  Same block schedule and order as the
    host recursion and the oracle, states to rounding, reproducible run to run; a second
    advance continues from the carried records."""
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 5))
    eps2 = eps * eps
    s0 = nbh.AdvancerState(m, q, p)
    a = nbh.advancer_advance(ctx, s0, span, 1e-4, eps2, resident=2, tree_max=tree_max)
    b = nbh.advancer_advance(ctx, s0, span, 1e-4, eps2, resident=1)
    assert a.stats == b.stats and np.all(a.t == span)
    assert_records_close(a, b)
    if n <= 1024:
        c = oracle.advancer_advance(oracle.AdvancerState(m, q, p), span, 1e-4, eps2)
        assert a.stats == c.stats and np.array_equal(a.id, c.id)
        assert rel(a.q, c.q) <= 1e-9 and rel(a.p, c.p) <= 1e-9
    a_again = nbh.advancer_advance(ctx, s0, span, 1e-4, eps2, resident=2, tree_max=tree_max)
    assert np.array_equal(a.state, a_again.state, equal_nan=True) and np.array_equal(a.id, a_again.id)
    a2 = nbh.advancer_advance(ctx, a, span, 1e-4, eps2, resident=2, tree_max=tree_max)
    b2 = nbh.advancer_advance(ctx, b, span, 1e-4, eps2, resident=1)
    assert a2.stats == b2.stats and np.all(a2.t == 2 * span)
    assert_records_close(a2, b2, tol=1e-8)


@pytest.mark.parametrize("n,span,tree_max,shape", [(700, 1.0 / 32, 0, "128,2"), (700, 1.0 / 32, 96, "512,1"),
                                                   (700, 1.0 / 32, 0, "256,3"),
                                                   (3000, 2.0 ** -16, 512, "256,2")])
def test_advancer_block_recursion_on_device_forced_grids(ctx, oracle, monkeypatch, n, span, tree_max, shape):
    """NBK_ADT_SHAPE = "threads,ctas" forces small grids, so that the kernel's multi-pass loops
    (more slots than threads, active blocks wider than a CTA, passive slices of several
    shared-memory tiles, several warps per active slot) run at sizes a test can afford."""
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 5))
    s0 = nbh.AdvancerState(m, q, p)
    b = nbh.advancer_advance(ctx, s0, span, 1e-4, 1e-4, resident=1)
    monkeypatch.setenv("NBK_ADT_SHAPE", shape)
    a = nbh.advancer_advance(ctx, s0, span, 1e-4, 1e-4, resident=2, tree_max=tree_max)
    assert a.stats == b.stats and np.all(a.t == span)
    assert_records_close(a, b)


@pytest.mark.parametrize("n", [148 * 128 + 40, 148 * 256 + 40])
def test_advancer_block_recursion_on_device_wide_grids(ctx, oracle, n):
    """The 256- and 512-thread shapes of the device recursion (one slot per thread up to
    148 x 512 rows); subtrees of up to 512 rows under the host's top of the recursion."""
    m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(n, 20244))
    s0 = nbh.AdvancerState(m, q, p)
    span = 2.0 ** -20
    a = nbh.advancer_advance(ctx, s0, span, 1e-4, 1e-6, resident=2)
    b = nbh.advancer_advance(ctx, s0, span, 1e-4, 1e-6, resident=1)
    assert a.stats == b.stats and np.all(a.t == span) and a.stats[0] > 500
    assert_records_close(a, b)


def test_device_recursion_argument_and_state_errors():
    """nbk_adv_advance_range / nbk_lf_advance_subsystem on a context with nothing uploaded, and
    with prefixes beyond their limits: an error code and a message, never a launch."""
    import ctypes as C
    import nbk
    lib = nbk.load_library()
    with nbk.Context(0) as c:
        h = (C.c_double * 4)(1.0, 1.0, 1.0, 1.0)
        perm = (C.c_int * 4)()
        assert lib.nbk_adv_advance_range(c._h, 4, 0.5, 1e-4, 0.0, h, perm, None) == -4  # NBK_ERR_STATE
        assert b"nbk_adv_upload" in lib.nbk_last_error(c._h)
        assert lib.nbk_lf_advance_subsystem(c._h, 4, 0.5, 1e-3, h, h, perm, None) == -4
        rec = np.zeros((8, nbh.ADV_STRIDE))
        rec[:, 1] = 1.0
        assert lib.nbk_adv_upload(c._h, 8, rec.ctypes.data_as(C.POINTER(C.c_double))) == 0
        assert lib.nbk_adv_advance_range(c._h, 9, 0.5, 1e-4, 0.0, h, perm, None) == -1  # imax > n
        assert lib.nbk_adv_advance_range(c._h, 4, -1.0, 1e-4, 0.0, h, perm, None) == -1
        assert lib.nbk_adv_advance_range(c._h, 4, 0.5, 1e-4, 0.0, None, perm, None) == -1
        assert lib.nbk_adv_range_max() == 1024 and lib.nbk_lf_subsystem_max() == 1024
