"""GPU: the OCaml binding's C stubs (ocaml/nbk_stubs.c) and ocaml/nbk.ml, RUN against the device.

tests/caml_shim/ compiles the stubs unchanged against a stand-in for the OCaml runtime headers;
here they are called with OCaml-shaped values (tagged ints, boxed doubles, tuples for the
(t0, t1, s0, s1) ranges, Bigarray.Array1 over numpy memory) through BOTH calling conventions the
OCaml toolchains use - native (all arguments) and bytecode (argv, argn; externals of more than five
arguments) - and must return exactly what the same entry point of the C ABI returns when called
directly: the stubs are mechanical, so any difference is an argument in the wrong place.

Then the last link: ocaml/nbk.ml itself (its `external` declarations, pack1 / pack3 / vec / ctx) is
run by the interpreter oracle/miniml over those real stubs - OCaml source -> C stub -> libnbk ->
kernel, with miniml standing in for ocamlopt and caml_shim for the runtime."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "caml_shim"))
import harness  # noqa: E402

pytestmark = pytest.mark.gpu
CONVENTIONS = [False, True]


@pytest.fixture(scope="module")
def stubs():
    return harness.Stubs()


@pytest.fixture(scope="module")
def mlctx(stubs):
    h = stubs.call("create", 0)
    yield h
    stubs.finalize(h)


@pytest.fixture(scope="module")
def bodies(oracle):
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(777, 424242))
    return m, q, p


def flat(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1)


@pytest.mark.parametrize("bytecode", CONVENTIONS)
def test_host_buffer_sweeps(ctx, stubs, mlctx, bodies, bytecode):
    m, q, p = bodies
    n = len(m)
    v = p / m[:, None]
    rng = (100, 613, 5, 777)
    nt = rng[1] - rng[0]
    g, dg = np.zeros(3 * nt), np.zeros(3 * nt)
    stubs.call("grad_v_dgrad_v_sum", mlctx, m, flat(q), flat(p), 1e-4, rng, g, dg, bytecode=bytecode)
    g0, dg0 = ctx.grad_v_dgrad_v_sum(m, q, p, 1e-4, targets=rng[:2], sources=rng[2:])
    assert np.array_equal(g.reshape(-1, 3), g0) and np.array_equal(dg.reshape(-1, 3), dg0)
    g1 = np.zeros(3 * nt)
    stubs.call("grad_v_sum", mlctx, m, flat(q), 1e-4, rng, g1, bytecode=bytecode)
    assert np.array_equal(g1.reshape(-1, 3), ctx.grad_v_sum(m, q, 1e-4, targets=rng[:2], sources=rng[2:]))
    phi = np.zeros(nt)
    tot = stubs.call("v_sum", mlctx, m, flat(q), 1e-4, rng, phi, bytecode=bytecode)
    phi0, tot0 = ctx.v_sum(m, q, 1e-4, targets=rng[:2], sources=rng[2:])
    assert np.array_equal(phi, phi0) and tot == tot0
    dv, tm = np.zeros(3 * nt), np.zeros(nt)
    stubs.call("kick_sum", mlctx, m, flat(q), flat(v), 1e-3, 1e-4, rng, dv, tm, bytecode=bytecode)
    dv0, tm0 = ctx.kick_sum(m, q, v, 1e-3, 1e-4, targets=rng[:2], sources=rng[2:])
    assert np.array_equal(dv.reshape(-1, 3), dv0) and np.array_equal(tm, tm0)
    s = np.zeros(3 * n)
    stubs.call("interact_sum", mlctx, m, flat(q), 300, 1e-4, s, bytecode=bytecode)
    assert np.array_equal(s.reshape(-1, 3), ctx.interact_sum(m, q, 300, 1e-4))
    dvb, tmb = np.zeros(3 * n), np.zeros(n)
    stubs.call("kick_block_sum", mlctx, m, flat(q), flat(v), 200, 1e-3, 1e-4, dvb, tmb, bytecode=bytecode)
    dvb0, tmb0 = ctx.kick_block_sum(m, q, v, 1e-3, 1e-4, 200)
    assert np.array_equal(dvb.reshape(-1, 3), dvb0) and np.array_equal(tmb, tmb0)


def test_tuple_results_and_int32_arrays(ctx, stubs, mlctx, bodies):
    m, q, p = bodies
    n = len(m)
    ke, pe = stubs.call("energy", mlctx, m, flat(q), flat(p), 1e-4)       # float * float
    assert (ke, pe) == ctx.energy(m, q, p, 1e-4)
    assert stubs.call("kinetic_energy", mlctx, m, flat(p)) == ke
    assert stubs.call("potential_energy", mlctx, m, flat(q), 1e-4) == pe
    emin, partner, count = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
    for bytecode in CONVENTIONS:
        stubs.call("binary_scan", mlctx, m, flat(q), flat(p), 0.0, 0.0, (0, n, 0, n), emin, partner, count,
                   bytecode=bytecode)
        e0, p0, c0 = ctx.binary_scan(m, q, p, 0.0, 0.0)
        assert np.array_equal(emin, e0) and np.array_equal(partner, p0) and np.array_equal(count, c0)
    npairs = int(count.sum()) // 2
    pi, pj, e = np.zeros(npairs, np.int32), np.zeros(npairs, np.int32), np.zeros(npairs)
    found = stubs.call("binary_list", mlctx, m, flat(q), flat(p), 0.0, 0.0, pi, pj, e)
    pi0, pj0, e0 = ctx.binary_list(m, q, p, 0.0, 0.0)
    assert found == npairs == len(pi0) and np.array_equal(pi, pi0) and np.array_equal(pj, pj0)
    assert np.array_equal(e, e0)
    assert stubs.call("device_count", mlctx) == 1


@pytest.mark.parametrize("bytecode", CONVENTIONS)
def test_tangent_sums_with_tuple_arguments(ctx, stubs, mlctx, bodies, bytecode):
    m, q, p = (a[:200].copy() for a in bodies)
    n, K = 200, 3
    rs = np.random.default_rng(5)
    dq, dp = rs.normal(size=(K, n, 3)), rs.normal(size=(K, n, 3))
    gt, dgt = np.zeros(K * 3 * n), np.zeros(K * 3 * n)
    stubs.call("grad_v_dgrad_v_sum_tangent", mlctx, m, flat(q), flat(p), K, flat(dq), flat(dp), 0.0,
               (0, n, 0, n), gt, dgt, bytecode=bytecode)
    _g, _dg, gt0, dgt0 = ctx.grad_v_dgrad_v_sum_tangent(m, q, p, dq, dp, 0.0)
    assert np.array_equal(gt.reshape(K, n, 3), gt0) and np.array_equal(dgt.reshape(K, n, 3), dgt0)
    # the advance_body-shaped entry: three tuples of bigarrays
    t = np.zeros(n)
    f, df = rs.normal(size=(n, 3)), rs.normal(size=(n, 3))
    t_tan = np.zeros((K, n))
    f_tan, df_tan = rs.normal(size=(K, n, 3)), rs.normal(size=(K, n, 3))
    nt_tan = rs.normal(size=K)
    outs = [np.zeros(3), np.zeros(3), np.zeros(3 * K), np.zeros(3 * K)]
    stubs.call("grad_v_dgrad_v_sum_predicted_tangent", mlctx,
               (t, m, flat(q), flat(p), flat(f), flat(df)), K,
               (flat(t_tan), flat(dq), flat(dp), flat(f_tan), flat(df_tan)), 1e-3, nt_tan, 0.0,
               (7, 8, 0, n), tuple(outs), bytecode=bytecode)
    ref = ctx.grad_v_dgrad_v_sum_predicted_tangent(t, m, q, p, f, df, t_tan, dq, dp, f_tan, df_tan,
                                                   1e-3, nt_tan, 0.0, targets=(7, 8))
    for a, b in zip(outs, ref):
        assert np.array_equal(a, np.asarray(b).reshape(-1))


def test_resident_protocols(ctx, stubs, mlctx, bodies, oracle):
    m, q, p = (a[:96].copy() for a in bodies)
    n = 96
    # Omelyan rows
    stubs.call("oa_upload", mlctx, m, flat(q), flat(p))
    stubs.call("oa_steps", mlctx, 1.0 / 128, 1e-4, 3)
    qo, po = np.zeros(3 * n), np.zeros(3 * n)
    stubs.call("oa_download", mlctx, qo, po)
    ctx.oa_upload(m, q, p)
    ctx.oa_steps(1.0 / 128, 1e-4, 3)
    q0, p0 = ctx.oa_download()
    assert np.array_equal(qo.reshape(-1, 3), q0) and np.array_equal(po.reshape(-1, 3), p0)
    el = np.zeros(4 * n)
    ke, pe = stubs.call("oa_diagnostics", mlctx, 1e-4, el)
    ke0, pe0, el0 = ctx.oa_diagnostics(1e-4)
    assert (ke, pe) == (ke0, pe0) and np.array_equal(el.reshape(-1, 4), el0)
    # Hermite: upload, the whole stepping loop in one call, download
    f, df = oracle.start_bs_sweep(m, q, p, 0.0)
    h = 6.0 * 1e-3 * np.linalg.norm(f, axis=1) / (m * np.linalg.norm(df, axis=1))
    t = np.zeros(n)
    stubs.call("hermite_upload", mlctx, t, m, flat(q), flat(p), h, flat(f), flat(df))
    steps = stubs.call("hermite_advance_resident", mlctx, 0.25, 1e-3, 0.0, 1 << 40)
    outs = [np.zeros(n), np.zeros(3 * n), np.zeros(3 * n), np.zeros(n), np.zeros(3 * n), np.zeros(3 * n)]
    stubs.call("hermite_download", mlctx, *outs)
    ctx.hermite_upload(t, m, q, p, f, df, h)
    steps0 = ctx.hermite_advance_resident(0.25, 1e-3, 0.0)
    ref = ctx.hermite_download(n)
    assert steps == steps0 > n
    for a, b in zip(outs, ref):
        assert np.array_equal(a, np.asarray(b).reshape(-1))
    # Leapfrog rows: upload, a kick block with its drift, the device sort, download
    v = p / m[:, None]
    dtm = np.linspace(1.0, 2.0, n)[::-1].copy()
    stubs.call("lf_upload", mlctx, dtm, m, flat(q), flat(v))
    back = np.zeros(n)
    stubs.call("lf_kick_block", mlctx, n, 40, 1e-2, 1e-3, True, 5e-4, back)
    perm = np.zeros(n, np.int32)
    stubs.call("lf_sort_sub", mlctx, perm)
    d1, q1, v1 = np.zeros(n), np.zeros(3 * n), np.zeros(3 * n)
    stubs.call("lf_download", mlctx, d1, q1, v1)
    assert sorted(perm.tolist()) == list(range(n)) and np.all(np.diff(d1) >= 0)
    assert np.array_equal(np.sort(back), d1)
    ctx.lf_upload(dtm, m, q, v)
    lib = ctx._lib
    import ctypes as C
    back0 = np.zeros(n)
    assert lib.nbk_lf_kick_block(ctx._h, n, 40, C.c_double(1e-2), C.c_double(1e-3), 1, 1, C.c_double(5e-4),
                                 back0.ctypes.data_as(C.POINTER(C.c_double))) == 0
    perm0 = ctx.lf_sort_sub(n)
    d0, q0, v0 = ctx.lf_download(n)
    assert np.array_equal(back, back0) and np.array_equal(perm, perm0)
    assert np.array_equal(d1, d0) and np.array_equal(q1.reshape(-1, 3), q0) and np.array_equal(v1.reshape(-1, 3), v0)


def test_errors_from_the_library_become_ocaml_failures(stubs, mlctx, bodies):
    m, q, p = bodies
    with pytest.raises(harness.OCamlFailure) as ei:   # target range outside the body array
        stubs.call("grad_v_sum", mlctx, m, flat(q), 0.0, (0, len(m) + 5, 0, len(m)), np.zeros(3 * (len(m) + 5)))
    assert "Failure:" in str(ei.value)
    # the context is still usable afterwards
    g = np.zeros(3)
    stubs.call("grad_v_sum", mlctx, m, flat(q), 0.0, (0, 1, 0, len(m)), g)
    assert np.isfinite(g).all() and np.abs(g).sum() > 0


@pytest.mark.parametrize("bytecode", CONVENTIONS)
def test_nbk_ml_runs_over_the_real_stubs(ctx, stubs, bodies, bytecode):
    """ocaml/nbk.ml interpreted by oracle/miniml with its externals routed to the compiled stubs:
    Nbk.ctx (lazy create), Nbk.vec / pack1 / pack3, the range tuple, unit and float * float results."""
    sys.path.insert(0, ROOT)
    from oracle.miniml import Interp, call
    it = Interp(search_path=[os.path.join(ROOT, "ocaml")])
    harness.bind_miniml(it, stubs, bytecode=bytecode)
    it.load_file(os.path.join(ROOT, "ocaml", "nbk.ml"))
    prog = it.load_source("""
type body = { m : float; q : float array; p : float array }
let forces bs eps2 =
  let n = Array.length bs in
  let m = Nbk.pack1 (fun b -> b.m) bs and q = Nbk.pack3 (fun b -> b.q) bs
  and p = Nbk.pack3 (fun b -> b.p) bs in
  let g = Nbk.vec (3 * n) and dg = Nbk.vec (3 * n) in
  Nbk.grad_v_dgrad_v_sum (Nbk.ctx ()) m q p eps2 (0, n, 0, n) g dg;
  let ke, pe = Nbk.energy (Nbk.ctx ()) m q p eps2 in
  (Array.init n (fun i -> Array.init 3 (fun k -> g.{3 * i + k})),
   Array.init n (fun i -> Array.init 3 (fun k -> dg.{3 * i + k})), ke +. pe,
   Nbk.device_count (Nbk.ctx ()), Nbk.lf_subsystem_max ())
""", "Prog")
    m, q, p = (a[:300] for a in bodies)
    from oracle.miniml import Record
    bs = [Record(dict(m=float(m[i]), q=[float(x) for x in q[i]], p=[float(x) for x in p[i]]))
          for i in range(300)]
    g, dg, e, ndev, lfmax = call(prog.vals["forces"][0], bs, 1e-4)
    g0, dg0 = ctx.grad_v_dgrad_v_sum(m, q, p, 1e-4)
    ke0, pe0 = ctx.energy(m, q, p, 1e-4)
    assert np.array_equal(np.array(g), g0) and np.array_equal(np.array(dg), dg0)
    assert e == ke0 + pe0 and ndev == 1 and lfmax == ctx._lib.nbk_lf_subsystem_max()
