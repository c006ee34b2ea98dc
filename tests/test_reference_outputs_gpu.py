"""GPU: the CUDA path against what the REFERENCE'S OWN SOURCE TEXT computes.

oracle/ref_fixtures/ref_miniml/*.hex hold the outputs of the unmodified reference modules (run by
the interpreter oracle/miniml over /root/reference, committed; tests/test_ref_fixtures.py) for the
seeded bodies under oracle/ref_fixtures/inputs/.  The other GPU tests compare the kernels with
the C restatement in oracle/ (which reproduces those files bit for bit); here the comparison is
made directly, value by value, with nothing of this repository's CPU code in between:

  per pair   Base.v / grad_v / grad_v_dgrad_v (eps = 0 and eps^2 = 0.0016), Leapfrog.kick - the
             pair time-scale BIT-EXACT, the rest <= 1e-12
  sums       Hermite.start_bs forces and first steps, energies, softened too
  runs       Hermite.advance (also resumed), Leapfrog.advance_const_dt / advance (eta = 1e-2, 1e-3),
             OA.advance, Advancer.A.advance (also resumed, softened, with the external field of
             test/advance_test.ml), Integrator.constant_sf_integrate - same step counts, same
             block order, states to the tolerances of the oracle comparisons
  analysis   binaries / binaries_threshold / tightest_binary (same pairs, same order, energies
             <= 1e-12), bodies_to_el_phase_space; Ics.adjust_frame / rescale_to_standard_units
  tangents   DFloat_pushforward.jacobian_advance, every entry of the 6n x 6n Jacobian, and
             omega_pushforward
  one N-body time unit   the energy-error trajectories of Hermite.advance and Leapfrog.advance at
             every quarter of a time unit (the north star's "must agree over 1 N-body time unit")
"""
import os
import struct

import numpy as np
import pytest

import nbh

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KIT = os.path.join(ROOT, "oracle", "ref_fixtures")


def _f(bits):
    return struct.unpack("<d", struct.pack("<Q", int(bits, 16)))[0]


class Sheet:
    def __init__(self, name):
        self.v = {}
        for ln in open(os.path.join(KIT, "ref_miniml", name)):
            key, i, j, k, bits = ln.split()
            self.v[(key, int(i), int(j), int(k))] = _f(bits)

    def scalar(self, key, i=0, j=0):
        return self.v[(key, i, j, 0)]

    def vec(self, key, i, j=0, n=3):
        return np.array([self.v[(key, i, j, k)] for k in range(n)])

    def rows(self, key, count, n=3):
        return np.array([[self.v[(key, i, 0, k)] for k in range(n)] for i in range(count)])

    def column(self, key, count):
        return np.array([self.v[(key, i, 0, 0)] for i in range(count)])

    def count(self, key):
        return 1 + max(i for (k, i, _, _) in self.v if k == key)


def read_bodies(name):
    rows = [ln.split() for ln in open(os.path.join(KIT, "inputs", name)).read().splitlines()[1:]]
    a = np.array([[_f(w) for w in r] for r in rows])
    return a[:, 0].copy(), a[:, 1:4].copy(), a[:, 4:7].copy()


@pytest.fixture(scope="module")
def big():
    return read_bodies("plummer_n64_seed20240.txt")


@pytest.fixture(scope="module")
def small():
    return read_bodies("plummer_n16_std_seed20241.txt")


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    den = np.linalg.norm(b)
    return np.linalg.norm(a - b) / den if den else np.linalg.norm(a - b)


def test_pair_functions(ctx, big):
    m, q, p = big
    s = Sheet("pairs.hex")
    v = p / m[:, None]
    for e, eps2 in enumerate((0.0, 0.0016)):
        for i in range(12):
            for j in range(i + 1, 12):
                tg, sr = (i, i + 1), (j, j + 1)
                g = ctx.grad_v_sum(m, q, eps2, targets=tg, sources=sr)[0]
                assert rel(g, s.vec(f"grad_v{e}", i, j)) <= 1e-12
                g2, dg = ctx.grad_v_dgrad_v_sum(m, q, p, eps2, targets=tg, sources=sr)
                assert rel(g2[0], s.vec(f"gvdgv_g{e}", i, j)) <= 1e-12
                assert rel(dg[0], s.vec(f"gvdgv_dg{e}", i, j)) <= 1e-12
                phi, _tot = ctx.v_sum(m, q, eps2, targets=tg, sources=sr)
                assert abs(phi[0] - s.scalar(f"v{e}", i, j)) <= 1e-12 * abs(s.scalar(f"v{e}", i, j))
    # Leapfrog.kick 1e-3 1e-4 on fresh copies: v += dv, dt_max = the pair time-scale, both ends
    for i in range(12):
        for j in range(i + 1, 12):
            for a, b, kv, kt in ((i, j, "kick_v1", "kick_dtmax1"), (j, i, "kick_v2", "kick_dtmax2")):
                dv, tmin = ctx.kick_sum(m, q, v, 1e-3, 1e-4, targets=(a, a + 1), sources=(b, b + 1))
                want = s.vec(kv, i, j)
                assert np.linalg.norm(v[a] + dv[0] - want) <= 1e-12 * np.linalg.norm(want - v[a]) + 1e-18
                assert tmin[0] == s.scalar(kt, i, j)          # bit-exact


def test_start_bs_sums_and_energies(ctx, big):
    m, q, p = big
    for sheet, prefix, eps2 in ((Sheet("sums.hex"), "", 0.0), (Sheet("wide.hex"), "eps_", 0.2 * 0.2)):
        g, dg = ctx.grad_v_dgrad_v_sum(m, q, p, eps2)
        f_ref, df_ref = sheet.rows(prefix + "start_f", 64), sheet.rows(prefix + "start_df", 64)
        for i in range(64):                                    # per body: f = -gsum, df = -dgsum
            assert rel(-g[i], f_ref[i]) <= 1e-12 and rel(-dg[i], df_ref[i]) <= 1e-12
        h = 6.0 * 1e-2 * np.linalg.norm(g, axis=1) / (m * np.linalg.norm(dg, axis=1))
        assert np.allclose(h, sheet.column(prefix + "start_h", 64), rtol=1e-12, atol=0)
        ke, pe = ctx.energy(m, q, p, eps2)
        assert abs(ke - sheet.scalar(prefix + "ke")) <= 1e-13 * ke
        assert abs(pe - sheet.scalar(prefix + "pe")) <= 1e-12 * abs(pe)
        assert abs(ke + pe - sheet.scalar(prefix + "energy")) <= 1e-12 * abs(pe)


def test_hermite_runs(ctx, small):
    m, q, p = small
    s = Sheet("traj.hex")
    a = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), 0.125, 1e-4)
    assert a.nsteps == int(s.scalar("hermite_nsteps")) == 200
    assert np.array_equal(a.t, s.column("hermite_t", 16))
    # observed: q 6e-17, p 3e-16, f 1e-15, df 1e-14
    assert rel(a.q, s.rows("hermite_q", 16)) <= 1e-13 and rel(a.p, s.rows("hermite_p", 16)) <= 1e-13
    assert rel(a.f, s.rows("hermite_f", 16)) <= 1e-12 and rel(a.df, s.rows("hermite_df", 16)) <= 1e-11
    # the next step h = h (eta h^2 |f| / (m |q_new - q_predicted|))^(1/4) (hermite.ml:87-93) divides by
    # the distance between corrector and predictor, ~ h^5: cancellation amplifies 1e-16 to ~2e-6
    assert np.allclose(a.h, s.column("hermite_h", 16), rtol=1e-4, atol=0)
    # two calls in a row, softened: the second resumes from the bodies' own h
    w = Sheet("wide.hex")
    eps2 = 0.2 * 0.2
    b = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), 0.0625, 1e-3, eps2)
    b = nbh.hermite_advance(ctx, b, 0.0625, 1e-3, eps2)
    assert b.nsteps == int(w.scalar("h2_nsteps")) and np.array_equal(b.t, w.column("h2_t", 16))
    assert rel(b.q, w.rows("h2_q", 16)) <= 1e-13 and rel(b.p, w.rows("h2_p", 16)) <= 1e-13
    assert rel(b.f, w.rows("h2_f", 16)) <= 1e-12 and rel(b.df, w.rows("h2_df", 16)) <= 1e-10


def test_leapfrog_runs(ctx, small):
    m, q, p = small
    s, w = Sheet("traj.hex"), Sheet("wide.hex")
    a = nbh.leapfrog_advance_const_dt(ctx, nbh.LeapfrogState(m, q, p), 1.0 / 256, 8)
    assert np.array_equal(a.id - 1, s.column("lfc_id", 16).astype(int))
    assert np.allclose(a.t, s.column("lfc_t", 16), rtol=0, atol=1e-17)
    assert rel(a.q, s.rows("lfc_q", 16)) <= 1e-13 and rel(a.v, s.rows("lfc_v", 16)) <= 1e-12
    for sheet, key, dt, eta in ((s, "lfa", 1.0 / 64, 1e-2), (w, "lf3", 1.0 / 32, 1e-3)):
        b = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), dt, eta)
        assert np.array_equal(b.id - 1, sheet.column(key + "_id", 16).astype(int)), key   # same order
        assert np.allclose(b.t, sheet.column(key + "_t", 16), rtol=0, atol=1e-16)
        # positions and velocities: observed 2e-16 / 2e-15.  dt_max: the sweep evaluates every pair
        # of a kick block from the velocities at its start, the reference kicks pair after pair, so
        # the pair time-scales (which depend on v) differ at O(dt) - 1e-3 here; the
        # block schedule, the order of the bodies and the trajectory are the reference's
        assert np.allclose(b.dt_max, sheet.column(key + "_dtmax", 16), rtol=1e-2, atol=0)
        assert np.allclose(b.q, sheet.rows(key + "_q", 16), rtol=0, atol=1e-13)
        assert np.allclose(b.v, sheet.rows(key + "_v", 16), rtol=0, atol=1e-12)
    ke, pe = ctx.energy(b.m, b.q, b.v * b.m[:, None], 0.0)
    assert abs(ke + pe - w.scalar("lf3_energy")) <= 1e-10 * abs(pe)


def test_omelyan_runs(ctx, small):
    m, q, p = small
    s, w = Sheet("traj.hex"), Sheet("wide.hex")
    q1, p1, t1 = nbh.omelyan_advance(ctx, m, q, p, 0.0, 1.0 / 64)
    assert t1 == s.scalar("oa_t", 0)
    assert rel(q1, s.rows("oa_q", 16)) <= 1e-13 and rel(p1, s.rows("oa_p", 16)) <= 1e-12
    qa, pa, ta = q, p, 0.0
    for _ in range(4):
        qa, pa, ta = nbh.omelyan_advance(ctx, m, qa, pa, ta, 1.0 / 128, 0.2 * 0.2)
    assert ta == w.scalar("oa4_t", 0)
    assert rel(qa, w.rows("oa4_q", 16)) <= 1e-12 and rel(pa, w.rows("oa4_p", 16)) <= 1e-11


def _check_advancer(state, sheet, key, tol=1e-9):
    n = len(state.id)
    assert np.array_equal(state.id, sheet.column(key + "_id", n).astype(int)), key     # sorted by hmax
    assert np.array_equal(state.t, sheet.column(key + "_t", n))
    assert rel(state.q, sheet.rows(key + "_q", n)) <= tol and rel(state.p, sheet.rows(key + "_p", n)) <= tol
    assert np.allclose(state.hmax, sheet.column(key + "_hmax", n), rtol=1e-3, atol=0)
    assert np.allclose(state.state[:, 8], sheet.column(key + "_h", n), rtol=0, atol=0)   # h: powers of two


@pytest.mark.parametrize("resident", [0, 2])
def test_advancer_runs(ctx, small, resident):
    m, q, p = small
    s, w = Sheet("traj.hex"), Sheet("wide.hex")
    a = nbh.advancer_advance(ctx, nbh.AdvancerState(m, q, p), 0.125, 1e-3, resident=resident)
    _check_advancer(a, s, "adv")
    eps2 = 0.2 * 0.2
    b = nbh.advancer_advance(ctx, nbh.AdvancerState(m, q, p), 0.0625, 1e-3, eps2, resident=resident)
    b = nbh.advancer_advance(ctx, b, 0.0625, 1e-3, eps2, resident=resident)
    _check_advancer(b, w, "a2")
    ke, pe = ctx.energy(b.m, b.q, b.p, eps2)
    assert abs(ke + pe - w.scalar("a2_energy")) <= 1e-10 * abs(pe)
    if resident == 0:   # an external field is host code: the per-interact_bodies path
        c = nbh.advancer_advance(ctx, nbh.AdvancerState(m, q, p), 0.0625, 1e-2, 0.0,
                                 extpot="plummer_test", resident=0)
        _check_advancer(c, w, "ax")
    d, errs = nbh.constant_sf_integrate(ctx, nbh.AdvancerState(m, q, p), 0.0625, ci=0.03125, sf=1e-3,
                                        resident=resident)
    assert len(errs) == 2
    _check_advancer(d, w, "int")


def test_analysis(ctx, big):
    m, q, p = big
    w = Sheet("wide.hex")
    for key, thr in (("bin", 0.0), ("thr", -abs(1e-4))):
        n = w.count(key + "_i")
        pi, pj, e = ctx.binary_list(m, q, p, 0.0, thr)
        # Analysis conses as it scans: its list is the loop order reversed (analysis.ml:836-861)
        assert np.array_equal(pi[::-1], w.column(key + "_i", n).astype(int))
        assert np.array_equal(pj[::-1], w.column(key + "_j", n).astype(int))
        assert np.allclose(e[::-1], w.column(key + "_e", n), rtol=1e-12, atol=0)
    emin, partner, _count = ctx.binary_scan(m, q, p, 0.0, 0.0)[:3]
    i = int(np.argmin(emin))
    assert {i, int(partner[i])} == {int(w.scalar("tight_i")), int(w.scalar("tight_j"))}
    el = nbh.bodies_to_el_phase_space(ctx, m, q, p, 0.0)
    want = np.array([[w.v[("el", i, 0, k)] for k in range(4)] for i in range(64)])
    assert np.allclose(el, want, rtol=1e-11, atol=0)


def test_ics_frame_and_standard_units(ctx, big):
    s = Sheet("ics.hex")
    fm, fq, fp = nbh.adjust_frame(*big)
    assert np.array_equal(fm, s.column("frame_m", 64))
    assert np.array_equal(fq, s.rows("frame_q", 64)) and np.array_equal(fp, s.rows("frame_p", 64))
    sm, sq, sp = nbh.rescale_to_standard_units(ctx, fm, fq, fp)
    assert np.allclose(sm, s.column("std_m", 64), rtol=1e-15, atol=0)
    assert np.allclose(sq, s.rows("std_q", 64), rtol=1e-12, atol=1e-15)
    assert np.allclose(sp, s.rows("std_p", 64), rtol=1e-12, atol=1e-15)


def test_dual_number_jacobian(ctx, small):
    m, q, p = small
    s = Sheet("dfloat.hex")
    for key, n, dt, sf in (("n3", 3, 0.03125, 1e-2), ("n5", 5, 0.015625, 1e-2)):
        jac, _out, _nsteps = nbh.dfloat_jacobian_advance(ctx, m[:n], q[:n], p[:n], dt, sf)
        want = np.array([[s.v[(key + "_jac", i, j, 0)] for j in range(6 * n)] for i in range(6 * n)])
        assert np.abs(jac - want).max() <= 1e-9 * np.abs(want).max(), key
        assert abs(nbh.omega_pushforward(jac) - s.scalar(key + "_omega")) <= 1e-9


def test_energy_error_trajectories_over_one_time_unit(ctx):
    """The reference's own Hermite (2457 steps) and adaptive Leapfrog runs over one N-body time unit
    (unit.hex: four advances of 0.25, 24-body Plummer sphere in standard units) beside the GPU
    path: energy error after every quarter, step counts, states."""
    m, q, p = read_bodies("plummer_n24_std_seed20242.txt")
    u = Sheet("unit.hex")
    n = len(m)
    e0 = u.scalar("he_energy", 0)
    assert abs(e0 + 0.25) < 1e-12
    h = nbh.HermiteState(m, q, p)
    for k in range(1, 5):
        h = nbh.hermite_advance(ctx, h, 0.25, 1e-4)
        de_ref = abs((u.scalar("he_energy", k) - e0) / e0)
        ke, pe = ctx.energy(h.m, h.q, h.p, 0.0)
        de_gpu = abs((ke + pe - e0) / e0)
        assert h.nsteps == int(u.scalar("he_nsteps", k)), k          # the same schedule
        assert abs(de_gpu - de_ref) <= 1e-3 * de_ref + 1e-13, (k, de_gpu, de_ref)
        q_ref = np.array([u.vec("he_q", k, i) for i in range(n)])
        p_ref = np.array([u.vec("he_p", k, i) for i in range(n)])
        assert np.array_equal(h.t, np.array([u.scalar("he_t", k, i) for i in range(n)]))
        assert rel(h.q, q_ref) <= 1e-9 and rel(h.p, p_ref) <= 1e-8, k
    assert 1e-10 < de_ref < 1e-6           # a real, resolved energy error, the same on both sides
    # adaptive leapfrog: positions follow the reference's to rounding while the block schedule is
    # the same; a dt_max within 1e-3 of a block boundary can flip a decision, after
    # which the runs agree in the energy error rather than step for step
    s = nbh.LeapfrogState(m, q, p)
    for k in range(1, 5):
        s = nbh.leapfrog_advance(ctx, s, 0.25, 2e-3)
        de_ref = abs((u.scalar("lf_energy", k) - e0) / e0)
        ke, pe = ctx.energy(s.m, s.q, s.v * s.m[:, None], 0.0)
        de_gpu = abs((ke + pe - e0) / e0)
        assert np.all(s.t == 0.25 * k)
        assert 0.5 < (de_gpu + 1e-12) / (de_ref + 1e-12) < 2.0, (k, de_gpu, de_ref)
        ids_ref = np.array([u.scalar("lf_id", k, r) for r in range(n)]).astype(int)
        q_ref = np.array([u.vec("lf_q", k, r) for r in range(n)])
        order = np.argsort(ids_ref)
        mine = np.argsort(s.id - 1)
        print(f"leapfrog t = {0.25 * k}: dE gpu {de_gpu:.3e} ref {de_ref:.3e}, same order "
              f"{np.array_equal(s.id - 1, ids_ref)}, max |dq| {np.abs(s.q[mine] - q_ref[order]).max():.2e}")
        assert np.abs(s.q[mine] - q_ref[order]).max() <= 1e-4
