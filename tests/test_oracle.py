"""CPU: the oracle restatement against (i) hand-computed two-body known answers, (ii) its own
structural identities (the reference's symmetric loop == per-target ascending-j sums, bit for
bit; binary128 truth), (iii) the committed golden fixtures, bit for bit, and (iv) the
product-side IC generator (host logic) bit for bit."""
import glob
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---------------------------------------------------------------- known answers (base.ml:60-110)
def test_two_body_known_answers(oracle):
    m = np.array([1.0, 1.0])
    q = np.array([[0.0, 0.0, 0.0], [3.0, 4.0, 0.0]])
    p = np.array([[1.0, 0.0, 0.0], [0.0, 2.0, 0.0]])
    # r = 5: v = -1/5; grad_v = (q1-q2)/125
    assert oracle.total_potential_energy(m, q) == -0.2
    g, dg = oracle.grad_v_dgrad_v_sum(m, q, p)
    assert np.allclose(g[0], [-0.024, -0.032, 0.0], rtol=2e-16, atol=0)
    assert np.array_equal(g[1], -g[0])
    # w = (1,-2,0), r.w = 5: dgv = w/125 - (3*5/3125) r = (0.0224, 0.0032, 0)
    assert np.allclose(dg[0], [0.0224, 0.0032, 0.0], rtol=1e-15, atol=0)
    assert np.array_equal(dg[1], -dg[0])
    # softened: |q1-q2|^2 = 9, eps2 = 16 -> s = 5 again
    q2 = np.array([[0.0, 0.0, 0.0], [3.0, 0.0, 0.0]])
    assert oracle.total_potential_energy(m, q2, 16.0) == -0.2
    assert np.allclose(oracle.grad_v_sum(m, q2, 16.0)[0], [-0.024, 0.0, 0.0], rtol=2e-16, atol=0)


def test_kick_known_answer(oracle):
    """leapfrog.ml:70-104 on a pair 2 apart.  At rest: c = dt/8, and the time-scale is NaN
    (tc = eta*sqrt(r2/0) = inf, dtc = 0*inf = NaN) so dt_max keeps its value (:102-104).
    With a perpendicular relative velocity 1: vdr = 0, tff = eta*sqrt(8/mtot), tc = eta*2."""
    m = np.array([1.0, 3.0])
    q = np.array([[0.0, 0.0, 0.0], [2.0, 0.0, 0.0]])
    v = np.zeros((2, 3))
    v2, dtm = oracle.kick_all_pairs(m, q, v, np.full(2, np.inf), 0.5, 0.25)
    assert np.array_equal(v2, [[3 * 0.25 / 8 * 2, 0, 0], [-1 * 0.25 / 8 * 2, 0, 0]])
    assert np.all(np.isinf(dtm))
    dv, tmin = oracle.kick_sum(m, q, v, 0.5, 0.25)
    assert np.array_equal(dv, v2) and np.all(np.isinf(tmin))
    v[1, 1] = 1.0
    _, dtm = oracle.kick_all_pairs(m, q, v, np.full(2, np.inf), 0.5, 0.0)
    assert np.array_equal(dtm, np.full(2, 0.5 * np.sqrt(8.0 / 4.0)))
    _, tmin = oracle.kick_sum(m, q, v, 0.5, 0.0)
    assert np.array_equal(tmin, dtm)


# ---------------------------------------------------------------- structural identities
@pytest.mark.parametrize("n,eps", [(2, 0.0), (3, 0.0), (257, 0.0), (1000, 0.04)])
def test_symmetric_loop_equals_per_target_sums_bitwise(oracle, n, eps):
    """grad_v_dgrad_v is antisymmetric bit for bit, so Hermite.start_bs'
    symmetric i<j loop (hermite.ml:154-167) equals ascending-j per-target sums exactly."""
    m, q, p = oracle.make_plummer(n, 5)
    f, df = oracle.start_bs_sweep(m, q, p, eps * eps)
    g, dg = oracle.grad_v_dgrad_v_sum(m, q, p, eps * eps)
    assert np.array_equal(f, -g) and np.array_equal(df, -dg)
    fo, dfo = oracle.start_bs_sweep(m, q, p, eps * eps, omp=True)
    assert np.allclose(fo, f, rtol=0, atol=1e-13 * np.abs(f).max())


def test_rectangles_add_up(oracle):
    m, q, p = oracle.make_plummer(300, 9)
    full = oracle.grad_v_dgrad_v_sum(m, q, p, 0.0, targets=(40, 90))
    a = oracle.grad_v_dgrad_v_sum(m, q, p, 0.0, targets=(40, 90), sources=(0, 120))
    b = oracle.grad_v_dgrad_v_sum(m, q, p, 0.0, targets=(40, 90), sources=(120, 300))
    for k in range(2):
        assert np.allclose(a[k] + b[k], full[k], rtol=1e-13, atol=0)


def test_against_binary128_truth(oracle):
    """Both the reference order and any other order must sit well inside 1e-12 of the
    binary128 sums."""
    m, q, p = oracle.make_plummer(2000, 20243)
    g, dg = oracle.grad_v_dgrad_v_sum(m, q, p, 0.0)
    gq, dgq = oracle.grad_v_dgrad_v_sum(m, q, p, 0.0, quad=True)
    assert np.max(np.linalg.norm(g - gq, axis=1) / np.linalg.norm(gq, axis=1)) < 1e-13
    assert np.max(np.linalg.norm(dg - dgq, axis=1) / np.linalg.norm(dgq, axis=1)) < 1e-13
    phi, phiq = oracle.v_sum(m, q, 0.0), oracle.v_sum(m, q, 0.0, quad=True)
    assert np.max(np.abs(phi - phiq) / np.abs(phiq)) < 1e-13
    pe, peq = oracle.total_potential_energy(m, q), oracle.total_potential_energy(m, q, quad=True)
    assert abs(pe - peq) < 1e-13 * abs(peq)
    assert abs(0.5 * phi.sum() - pe) < 1e-13 * abs(pe)


def test_newtons_third_law(oracle):
    m, q, p = oracle.make_plummer(500, 3)
    f, df = oracle.start_bs_sweep(m, q, p, 0.0)
    assert np.linalg.norm(f.sum(0)) < 1e-13 * np.linalg.norm(f, axis=1).sum()
    assert np.linalg.norm(df.sum(0)) < 1e-13 * np.linalg.norm(df, axis=1).sum()


def test_kick_sum_is_the_prekick_variant(oracle):
    """With dt = 0 the sequential sweep (leapfrog.ml:151-155) never changes v, so its dt_max
    must equal the Jacobi-order min BIT FOR BIT; with dt > 0 the velocity sums agree to
    rounding and dt_max differs at O(dt)."""
    m, q, p = oracle.make_hot_spherical(200, 4)
    v = p / m[:, None]
    _, dtm = oracle.kick_all_pairs(m, q, v, np.full(200, np.inf), 1e-3, 0.0)
    dv, tmin = oracle.kick_sum(m, q, v, 1e-3, 0.0)
    assert np.array_equal(tmin, dtm) and not dv.any()
    v2, dtm2 = oracle.kick_all_pairs(m, q, v, np.full(200, np.inf), 1e-3, 1e-4)
    dv, tmin = oracle.kick_sum(m, q, v, 1e-3, 1e-4)
    assert np.allclose(v + dv, v2, rtol=0, atol=1e-13 * np.abs(v2).max())
    assert np.allclose(tmin, dtm2, rtol=1e-2)
    # eta = infinity: every time-scale is NaN/inf and dt_max stays +inf (leapfrog.ml:160-165)
    _, dtm3 = oracle.kick_all_pairs(m, q, v, np.full(200, np.inf), np.inf, 1e-4)
    assert np.all(np.isinf(dtm3))


def test_el_phase_space(oracle):
    m, q, p = oracle.make_plummer(100, 8)
    el = oracle.bodies_to_el_phase_space(m, q, p)
    phi = oracle.v_sum(m, q)
    ke = (p * p).sum(1) / (2 * m)
    assert np.allclose(el[:, 0], ke + phi, rtol=1e-14)
    assert np.allclose(el[:, 1], el[:, 0] / m) and np.allclose(el[:, 3], el[:, 2] / m)
    assert np.allclose(el[:, 2], np.linalg.norm(np.cross(q, p), axis=1), rtol=1e-13)


# ---------------------------------------------------------------- golden fixtures, bit for bit
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "sweeps_*.npz"))))
def test_sweep_fixtures_bitwise(oracle, path):
    z = np.load(path)
    m, q, p, eps2 = z["m"], z["q"], z["p"], float(z["eps2"])
    g, dg = oracle.grad_v_dgrad_v_sum(m, q, p, eps2)
    f, df = oracle.start_bs_sweep(m, q, p, eps2)
    assert np.array_equal(g, z["gsum"]) and np.array_equal(dg, z["dgsum"])
    assert np.array_equal(f, z["f"]) and np.array_equal(df, z["df"])
    assert np.array_equal(oracle.grad_v_sum(m, q, eps2), z["grad"])
    assert np.array_equal(oracle.v_sum(m, q, eps2), z["phi"])
    assert oracle.total_potential_energy(m, q, eps2) == float(z["pe"])
    assert oracle.total_kinetic_energy(m, p) == float(z["ke"])
    dv, tmin = oracle.kick_sum(m, q, p / m[:, None], 1e-3, 1e-4)
    assert np.array_equal(dv, z["kick_dv"]) and np.array_equal(tmin, z["kick_tmin"])
    assert np.array_equal(oracle.bodies_to_el_phase_space(m, q, p, eps2), z["el"])


def test_trajectory_fixture_bitwise(oracle):
    z = np.load(os.path.join(GOLD, "trajectories_n16.npz"))
    m, q, p = z["m"], z["q"], z["p"]
    h = oracle.hermite_advance(oracle.HermiteState(m, q, p), 0.125, 1e-4)
    assert np.array_equal(h.q, z["hermite_q"]) and np.array_equal(h.p, z["hermite_p"])
    assert np.array_equal(h.h, z["hermite_h"]) and h.nsteps == int(z["hermite_nsteps"])
    lf = oracle.leapfrog_advance_const_dt(oracle.LeapfrogState(m, q, p), 1.0 / 256, 8)
    assert np.array_equal(lf.q, z["leapfrog_q"]) and np.array_equal(lf.v, z["leapfrog_v"])
    la = oracle.leapfrog_advance(oracle.LeapfrogState(m, q, p), 1.0 / 64, 1e-2)
    assert np.array_equal(la.id, z["leapfrog_adapt_id"])
    assert np.array_equal(la.q, z["leapfrog_adapt_q"])
    assert np.array_equal(la.dt_max, z["leapfrog_adapt_dtmax"])
    oq, op, ot = oracle.omelyan_advance(m, q, p, 0.0, 1.0 / 64)
    assert np.array_equal(oq, z["omelyan_q"]) and np.array_equal(op, z["omelyan_p"])
    a = oracle.advancer_advance(oracle.AdvancerState(m, q, p), 0.125, 1e-3, 0.0)
    assert np.array_equal(a.id, z["advancer_id"])
    assert np.array_equal(a.state, z["advancer_state"], equal_nan=True)


# ---------------------------------------------------------------- host logic: IC generator
@pytest.mark.parametrize("kind", ["plummer", "hot_spherical", "cold_spherical", "uniform_box"])
def test_product_ic_generator_matches_oracle_bitwise(oracle, kind):
    import nbh
    a = getattr(nbh, "make_" + kind)(777, 20243)
    b = getattr(oracle, "make_" + kind)(777, 20243)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    m, q, p = a
    m2, q2, p2 = nbh.add_mass_disc(m, q, p, 99)
    m3, q3, p3 = oracle.add_mass_disc(m, q, p, 99)
    assert np.array_equal(m2, m3) and np.array_equal(p2, p3)
    assert set(np.unique(m2)) <= {1.0, 10.0} and (m2 == 10.0).sum() > 0
    f1, f2 = nbh.adjust_frame(m2, q2, p2), oracle.adjust_frame(m2, q2, p2)
    assert all(np.array_equal(x, y) for x, y in zip(f1, f2))
