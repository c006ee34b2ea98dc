"""CPU: the compiled oracle (oracle/*.c, *.cc) against oracle/twin.py, a second restatement of
the same OCaml functions written independently in plain Python (IEEE binary64, no FMA, libm pow
and sqrt - the arithmetic of an ocamlopt build).  Bit-for-bit agreement on small systems: a
transcription slip in either text, or a compiler that fuses or reorders the C one, shows up as a
differing bit.  (Neither is the OCaml binary: DESIGN.md §6, "parity unpinned".)"""
import numpy as np
import pytest

from oracle import twin


def lists(a):
    return [[float(x) for x in row] for row in np.asarray(a)]


@pytest.fixture(scope="module", params=[(7, 3, 0.0), (12, 20241, 0.05)])
def bodies(request, oracle):
    n, seed, eps = request.param
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, seed))
    return m, q, p, eps * eps


def test_energies_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    assert oracle.total_potential_energy(m, q, eps2) == twin.total_potential_energy(list(m), lists(q), eps2)
    assert oracle.total_kinetic_energy(m, p) == twin.total_kinetic_energy(list(m), lists(p))


def test_hermite_start_bs_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    f, df = oracle.start_bs_sweep(m, q, p, eps2)
    tf, tdf, _ = twin.hermite_start_bs(1e-3, list(m), lists(q), lists(p), eps2)
    assert np.array_equal(f, np.array(tf)) and np.array_equal(df, np.array(tdf))


def test_hermite_advance_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    a = oracle.hermite_advance(oracle.HermiteState(m, q, p), 1.0 / 32, 1e-3, eps2)
    bs, nsteps = twin.hermite_advance(list(m), lists(q), lists(p), 1.0 / 32, 1e-3, eps2)
    assert a.nsteps == nsteps and [b["id"] for b in bs] == list(a.id)
    for k in ("q", "p", "f", "df"):
        assert np.array_equal(getattr(a, k), np.array([b[k] for b in bs])), k
    assert np.array_equal(a.t, np.array([b["t"] for b in bs]))
    assert np.array_equal(a.h, np.array([b["h"] for b in bs]))


def test_leapfrog_advance_bitwise(oracle, bodies):
    m, q, p, _ = bodies
    a = oracle.leapfrog_advance(oracle.LeapfrogState(m, q, p), 1.0 / 64, 1e-2)
    bs, nk = twin.leapfrog_advance(list(m), lists(q), lists(p), 1.0 / 64, 1e-2)
    assert [b["id"] for b in bs] == list(a.id) and a.nkicks == nk
    assert np.array_equal(a.q, np.array([b["q"] for b in bs]))
    assert np.array_equal(a.v, np.array([b["v"] for b in bs]))
    assert np.array_equal(a.dt_max, np.array([b["dt_max"] for b in bs]))
    assert np.array_equal(a.t, np.array([b["t"] for b in bs]))


def test_omelyan_advance_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    oq, op, ot = oracle.omelyan_advance(m, q, p, 0.0, 1.0 / 64, eps2)
    tq, tp, tt = twin.omelyan_advance(list(m), lists(q), lists(p), 0.0, 1.0 / 64, eps2)
    assert ot == tt and np.array_equal(oq, np.array(tq)) and np.array_equal(op, np.array(tp))


def test_advancer_advance_bitwise(oracle, bodies):
    m, q, p, eps2 = bodies
    a = oracle.advancer_advance(oracle.AdvancerState(m, q, p), 1.0 / 16, 1e-3, eps2)
    bs, stats = twin.advancer_advance(list(m), lists(q), lists(p), 1.0 / 16, 1e-3, eps2)
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
    assert np.array_equal(el, np.array(twin.bodies_to_el_phase_space(list(m), lists(q), lists(p), eps2)))
    pi, pj, e = oracle.binaries(m, q, p, eps2, 0.0)
    ref = twin.binaries_threshold(list(m), lists(q), lists(p), eps2, 0.0)
    assert list(zip(pi.tolist(), pj.tolist())) == [(i, j) for i, j, _ in ref]
    assert np.array_equal(e, np.array([x for _, _, x in ref]))


def test_ics_frame_and_units_bitwise(oracle):
    """Ics.adjust_frame and rescale_to_standard_units (ics.ml:161-175, 304-331) on raw draws."""
    m, q, p = oracle.make_plummer(9, 77)
    a = oracle.adjust_frame(m, q, p)
    t = twin.adjust_frame(list(m), lists(q), lists(p))
    for x, y in zip(a, t):
        assert np.array_equal(np.asarray(x), np.array(y))
    b = oracle.rescale_to_standard_units(*a, eps2=0.0025)
    u = twin.rescale_to_standard_units(*t, 0.0025)
    for x, y in zip(b, u):
        assert np.array_equal(np.asarray(x), np.array(y))
