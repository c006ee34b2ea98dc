"""CPU: the compiled oracle (oracle/*.c, *.cc) against oracle/twin.py, a second restatement of
the same OCaml functions written independently in plain Python (IEEE binary64, no FMA, libm pow
and sqrt - the arithmetic of an ocamlopt build).  Bit-for-bit agreement on small systems: a
transcription slip in either text, or a compiler that fuses or reorders the C one, shows up as a
differing bit.  (Neither is the OCaml text itself: that comparison is tests/test_ref_fixtures.py)"""
import numpy as np
import pytest

from oracle import twin


def fl(a):  # plain Python floats: the twin's arithmetic must not go through numpy scalars
    return [float(x) for x in a]


def lists(a):
    return [[float(x) for x in row] for row in np.asarray(a)]


@pytest.fixture(scope="module", params=[(7, 3, 0.0), (12, 20241, 0.05)])
def bodies(request, oracle):
    n, seed, eps = request.param
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, seed))
    return m, q, p, eps * eps


def test_energies_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    assert oracle.total_potential_energy(m, q, eps2) == twin.total_potential_energy(fl(m), lists(q), eps2)
    assert oracle.total_kinetic_energy(m, p) == twin.total_kinetic_energy(fl(m), lists(p))


def test_hermite_start_bs_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    f, df = oracle.start_bs_sweep(m, q, p, eps2)
    tf, tdf, _ = twin.hermite_start_bs(1e-3, fl(m), lists(q), lists(p), eps2)
    assert np.array_equal(f, np.array(tf)) and np.array_equal(df, np.array(tdf))


def test_hermite_advance_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    a = oracle.hermite_advance(oracle.HermiteState(m, q, p), 1.0 / 32, 1e-3, eps2)
    bs, nsteps = twin.hermite_advance(fl(m), lists(q), lists(p), 1.0 / 32, 1e-3, eps2)
    assert a.nsteps == nsteps and [b["id"] for b in bs] == list(a.id)
    for k in ("q", "p", "f", "df"):
        assert np.array_equal(getattr(a, k), np.array([b[k] for b in bs])), k
    assert np.array_equal(a.t, np.array([b["t"] for b in bs]))
    assert np.array_equal(a.h, np.array([b["h"] for b in bs]))


def test_leapfrog_advance_bitwise(oracle, bodies):
    m, q, p, _ = bodies
    a = oracle.leapfrog_advance(oracle.LeapfrogState(m, q, p), 1.0 / 64, 1e-2)
    bs, nk = twin.leapfrog_advance(fl(m), lists(q), lists(p), 1.0 / 64, 1e-2)
    assert [b["id"] for b in bs] == list(a.id) and a.nkicks == nk
    assert np.array_equal(a.q, np.array([b["q"] for b in bs]))
    assert np.array_equal(a.v, np.array([b["v"] for b in bs]))
    assert np.array_equal(a.dt_max, np.array([b["dt_max"] for b in bs]))
    assert np.array_equal(a.t, np.array([b["t"] for b in bs]))


def test_omelyan_advance_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    oq, op, ot = oracle.omelyan_advance(m, q, p, 0.0, 1.0 / 64, eps2)
    tq, tp, tt = twin.omelyan_advance(fl(m), lists(q), lists(p), 0.0, 1.0 / 64, eps2)
    assert ot == tt and np.array_equal(oq, np.array(tq)) and np.array_equal(op, np.array(tp))


def test_advancer_advance_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    a = oracle.advancer_advance(oracle.AdvancerState(m, q, p), 1.0 / 16, 1e-3, eps2)
    bs, stats = twin.advancer_advance(fl(m), lists(q), lists(p), 1.0 / 16, 1e-3, eps2)
    assert tuple(stats) == tuple(a.stats) and [b["id"] for b in bs] == list(a.id)
    cols = {"t": 0, "m": 1, "h": 8, "hmax": 9, "t0": 10}
    for k, c in cols.items():
        assert np.array_equal(a.state[:, c], np.array([b[k] for b in bs])), k
    vecs = {"q": 2, "p": 5, "dVdq0": 11, "dVdq1": 14, "dVdq2": 17, "p0": 20, "q0": 23, "q1": 26,
            "q2": 29, "cs": 32}
    for k, c in vecs.items():
        assert np.array_equal(a.state[:, c:c + 3], np.array([b[k] for b in bs])), k


def test_analysis_bitwise(oracle, bodies):
    """Analysis.binary_energy / binaries_threshold / bodies_to_el_phase_space (analysis.ml:809-919)."""
    m, q, p, eps2 = bodies
    el = oracle.bodies_to_el_phase_space(m, q, p, eps2)
    assert np.array_equal(el, np.array(twin.bodies_to_el_phase_space(fl(m), lists(q), lists(p), eps2)))
    pi, pj, e = oracle.binaries(m, q, p, eps2, 0.0)
    ref = twin.binaries_threshold(fl(m), lists(q), lists(p), eps2, 0.0)
    assert list(zip(pi.tolist(), pj.tolist())) == [(i, j) for i, j, _ in ref]
    assert np.array_equal(e, np.array([x for _, _, x in ref]))


def test_ics_frame_and_units_bitwise(oracle):
    """Ics.adjust_frame and rescale_to_standard_units (ics.ml:161-175, 304-331) on raw draws."""
    m, q, p = oracle.make_plummer(9, 77)
    a = oracle.adjust_frame(m, q, p)
    t = twin.adjust_frame(fl(m), lists(q), lists(p))
    for x, y in zip(a, t):
        assert np.array_equal(np.asarray(x), np.array(y))
    b = oracle.rescale_to_standard_units(*a, eps2=0.0025)
    u = twin.rescale_to_standard_units(*t, 0.0025)
    for x, y in zip(b, u):
        assert np.array_equal(np.asarray(x), np.array(y))


def _closed_form_tangent(m, q, p, dq, dp, eps2):
    """SURVEY.md §8 a14, in extended precision: per target i the sums over j != i of
    g = mu r/s^3, dg = mu [w/s^3 - 3 (r.w) r/s^5] and of their first-order variations
      dg_t  = mu [dr/s^3 - 3 (r.dr) r/s^5]
      ddg_t = mu [dw/s^3 - 3 (r.dr) w/s^5 - 3 ((dr.w + r.dw) r + (r.w) dr)/s^5 + 15 (r.w)(r.dr) r/s^7]
    with r = q_i - q_j, w = p_i/m_i - p_j/m_j, dr, dw their variations, s^2 = |r|^2 + eps2."""
    L = np.longdouble
    m, q, p, dq, dp = (np.asarray(a, dtype=L) for a in (m, q, p, dq, dp))
    n = len(m)
    out = [np.zeros((n, 3), dtype=L) for _ in range(4)]
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            mu = m[i] * m[j]
            r, w = q[i] - q[j], p[i] / m[i] - p[j] / m[j]
            dr, dw = dq[i] - dq[j], dp[i] / m[i] - dp[j] / m[j]
            s2 = r @ r + L(eps2)
            s = np.sqrt(s2)
            s3, s5, s7 = s * s2, s * s2 * s2, s * s2 * s2 * s2
            rw, rdr = r @ w, r @ dr
            out[0][i] += mu * r / s3
            out[1][i] += mu * (w / s3 - 3 * rw * r / s5)
            out[2][i] += mu * (dr / s3 - 3 * rdr * r / s5)
            out[3][i] += mu * (dw / s3 - 3 * rdr * w / s5 - 3 * ((dr @ w + r @ dw) * r + rw * dr) / s5
                               + 15 * rw * rdr * r / s7)
    return [np.asarray(o, dtype=np.float64) for o in out]


@pytest.mark.parametrize("eps2", [0.0, 1e-3])
def test_dual_number_tangents_match_closed_forms(oracle, eps2):
    """The oracle's stand-in for Diff.DFloat (oracle/num.hh: value + one tangent, textbook rules)
    pushed through DFloatBase.grad_v_dgrad_v (dFloat_advancer.ml:66-81) against the closed-form
    variations evaluated in extended precision - an independent derivation, so neither the dual
    arithmetic nor the formula is taken on trust.  And against central differences of the value
    sums."""
    rng = np.random.default_rng(5)
    n = 24
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 11))
    dq = rng.standard_normal((n, 3))
    dp = rng.standard_normal((n, 3)) * m[:, None]
    g, dg, gt, dgt = oracle.grad_v_dgrad_v_sum_tangent(m, q, p, dq, dp, eps2)
    cg, cdg, cgt, cdgt = _closed_form_tangent(m, q, p, dq, dp, eps2)
    for a, b in ((g, cg), (dg, cdg), (gt, cgt), (dgt, cdgt)):
        assert np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)) <= 1e-13
    g1, gt1 = oracle.grad_v_sum_tangent(m, q, dq, eps2)
    assert np.array_equal(g1, g) and np.max(np.abs(gt1 - gt)) <= 1e-13 * np.max(np.abs(gt))
    h = 1e-6  # central differences of the value sums along (dq, dp)
    gp, dgp = oracle.grad_v_dgrad_v_sum(m, q + h * dq, p + h * dp, eps2)
    gm, dgm = oracle.grad_v_dgrad_v_sum(m, q - h * dq, p - h * dp, eps2)
    assert np.max(np.abs((gp - gm) / (2 * h) - gt)) <= 1e-6 * np.max(np.abs(gt))
    assert np.max(np.abs((dgp - dgm) / (2 * h) - dgt)) <= 1e-6 * np.max(np.abs(dgt))


def test_leapfrog_const_dt_bitwise(oracle, bodies):
    """advance_const_dt on FRESH bodies (dt_max = 0): the geometric ladder of sub-steps the
    reference actually performs (leapfrog.ml:113-142) - both restatements follow it."""
    m, q, p, _ = bodies
    a = oracle.leapfrog_advance_const_dt(oracle.LeapfrogState(m, q, p), 1.0 / 128, 3)
    bs = twin.leapfrog_advance_const_dt(fl(m), lists(q), lists(p), 1.0 / 128, 3)
    assert [b["id"] for b in bs] == list(a.id)
    assert np.array_equal(a.q, np.array([b["q"] for b in bs]))
    assert np.array_equal(a.v, np.array([b["v"] for b in bs]))
    assert np.array_equal(a.t, np.array([b["t"] for b in bs]))


@pytest.mark.parametrize("n,seed", [(1, 1), (33, 20243), (200, 7)])
def test_ic_recipes_bitwise(oracle, n, seed):
    """make_plummer / make_hot_spherical (ics.ml:211-239, with random_vector, random_from_dist and
    adjust_frame) over the fixed splitmix64 stream of SURVEY.md §8d: the inputs of every BASELINE
    config, draw for draw."""
    for mk_o, mk_t in ((oracle.make_plummer, twin.make_plummer),
                       (oracle.make_hot_spherical, twin.make_hot_spherical)):
        a, b = mk_o(n, seed), mk_t(n, seed)
        for x, y in zip(a, b):
            assert np.array_equal(np.asarray(x), np.array(y))


def test_dual_number_hermite_jacobian_bitwise(oracle):
    """DFloat_pushforward.jacobian_advance over DFloat_hermite (dFloat_pushforward.ml:39-76,
    dFloat_hermite.ml:53-209 = hermite.ml over `diff`): the compiled template instantiated with
    the dual number against the Python Hermite functions run on twin.Dual - same rules, so the
    whole 6n x 6n Jacobian must agree bit for bit; and the map is symplectic to rounding."""
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(3, 4))
    jac, out = oracle.hermite_jacobian_advance(m, q, p, 1.0 / 16, 1e-3, 0.01)
    tj, to = twin.hermite_jacobian_advance(fl(m), lists(q), lists(p), 1.0 / 16, 1e-3, 0.01)
    assert np.array_equal(out, np.array(to))
    assert np.array_equal(jac, np.array(tj))
    assert abs(oracle.omega_pushforward(jac) - 1.0) < 1e-3


def test_golden_trajectories_reproduced_by_the_twin():
    """tests/golden/trajectories_n16.npz was written by the compiled oracle; the Python twin,
    started from the stored inputs, lands on the same bits - so the committed fixtures the GPU
    parity tests read are vouched for by both restatements."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "trajectories_n16.npz"))
    m, q, p = fl(g["m"]), lists(g["q"]), lists(g["p"])
    bs, nsteps = twin.hermite_advance(m, q, p, 0.125, 1e-4, 0.0)
    assert nsteps == int(g["hermite_nsteps"])
    assert np.array_equal(g["hermite_q"], np.array([b["q"] for b in bs]))
    assert np.array_equal(g["hermite_p"], np.array([b["p"] for b in bs]))
    assert np.array_equal(g["hermite_h"], np.array([b["h"] for b in bs]))
    la, _ = twin.leapfrog_advance(m, q, p, 1.0 / 64, 1e-2)
    assert [b["id"] for b in la] == list(g["leapfrog_adapt_id"])
    assert np.array_equal(g["leapfrog_adapt_q"], np.array([b["q"] for b in la]))
    assert np.array_equal(g["leapfrog_adapt_v"], np.array([b["v"] for b in la]))
    assert np.array_equal(g["leapfrog_adapt_dtmax"], np.array([b["dt_max"] for b in la]))
    oq, op, ot = twin.omelyan_advance(m, q, p, 0.0, 1.0 / 64, 0.0)
    assert ot == float(g["omelyan_t"])
    assert np.array_equal(g["omelyan_q"], np.array(oq)) and np.array_equal(g["omelyan_p"], np.array(op))
    ab, stats = twin.advancer_advance(m, q, p, 0.125, 1e-3, 0.0)
    assert list(g["advancer_stats"]) == stats and [b["id"] for b in ab] == list(g["advancer_id"])
    st = g["advancer_state"]
    assert np.array_equal(st[:, 2:5], np.array([b["q"] for b in ab]))
    assert np.array_equal(st[:, 5:8], np.array([b["p"] for b in ab]))
    assert np.array_equal(st[:, 9], np.array([b["hmax"] for b in ab]))


@pytest.mark.parametrize("name", ["sweeps_n64_eps0.npz", "sweeps_n257_eps004.npz"])
def test_golden_sweeps_reproduced_by_the_twin(name):
    """The sweep fixtures (acc+jerk in start_bs order, potential, kinetic energy, E-L phase space)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name))
    m, q, p, eps2 = fl(g["m"]), lists(g["q"]), lists(g["p"]), float(g["eps2"])
    f, df, _ = twin.hermite_start_bs(1e-3, m, q, p, eps2)
    assert np.array_equal(g["f"], np.array(f)) and np.array_equal(g["df"], np.array(df))
    assert float(g["pe"]) == twin.total_potential_energy(m, q, eps2)
    assert float(g["ke"]) == twin.total_kinetic_energy(m, p)
    assert np.array_equal(g["el"], np.array(twin.bodies_to_el_phase_space(m, q, p, eps2)))
