"""GPU: the dual-number Hermite step - DFloat_hermite.advance_body's loop
(dFloat_hermite.ml:115-150: every body predicted over duals, then DFloatBase.grad_v_dgrad_v
summed) through nbk_grad_v_dgrad_v_sum_predicted_tangent, and DFloat_pushforward.jacobian_advance
(dFloat_pushforward.ml:39-76) through the host mirror above it, against the oracle's template over
Dual.  Diff.DFloat itself is third-party and unpinned upstream, so the
tangent parity here is against the textbook dual-number rules the oracle restates."""
import numpy as np
import pytest

import nbh

pytestmark = pytest.mark.gpu


def _dual_state(oracle, n, K, seed):
    """A mid-run Hermite state (bodies at different times, f/df from a real sweep) with K random
    tangent directions of every field."""
    rng = np.random.default_rng(seed)
    m, q, p = oracle.make_plummer(n, seed)
    g, dg = oracle.grad_v_dgrad_v_sum(m, q, p, 0.0)
    f, df = -g, -dg
    t = rng.uniform(0.0, 0.01, n)
    h = rng.uniform(0.005, 0.02, n)
    tan = dict(t=rng.standard_normal((K, n)) * 1e-2, h=rng.standard_normal((K, n)) * 1e-3,
               q=rng.standard_normal((K, n, 3)), p=rng.standard_normal((K, n, 3)) * m[None, :, None],
               f=rng.standard_normal((K, n, 3)) * np.abs(f).mean(),
               df=rng.standard_normal((K, n, 3)) * np.abs(df).mean())
    return m, q, p, t, h, f, df, tan


@pytest.mark.parametrize("n,K,ib", [(2, 1, 0), (9, 3, 4), (300, 4, 299), (1500, 6, 77)])
def test_predicted_tangent_sums_match_dual_number_oracle(ctx, oracle, n, K, ib):
    """One advance_body force evaluation: target ib, all other bodies as sources, tangents of
    t, q, p, f, df on every body and of the target time nt = t_ib + h_ib."""
    m, q, p, t, h, f, df, tan = _dual_state(oracle, n, K, 100 + n)
    eps2 = 1e-6
    nt = t[ib] + h[ib]
    nt_tan = tan["t"][:, ib] + tan["h"][:, ib]
    g, dg, gt, dgt = ctx.grad_v_dgrad_v_sum_predicted_tangent(
        t, m, q, p, f, df, tan["t"], tan["q"], tan["p"], tan["f"], tan["df"], nt, nt_tan, eps2,
        targets=(ib, ib + 1))
    for k in range(K):
        val, d = oracle.dfloat_hermite_advance_body(
            t, m, q, p, h, f, df, tan["t"][k], tan["q"][k], tan["p"][k], tan["h"][k], tan["f"][k],
            tan["df"][k], ib, 0.01, eps2)
        # fp = -gsum, dfp = -dgsum (dFloat_hermite.ml:131-134)
        for mine, ref in ((-g[0], val["f"]), (-dg[0], val["df"]), (-gt[k, 0], d["f"]),
                          (-dgt[k, 0], d["df"])):
            assert np.linalg.norm(mine - ref) <= 1e-11 * np.linalg.norm(ref), (k, mine, ref)
    # values are those of the plain predicted sweep, bit for bit
    g0, dg0 = ctx.grad_v_dgrad_v_sum_predicted(t, m, q, p, f, df, nt, eps2, targets=(ib, ib + 1))
    assert np.array_equal(g, g0) and np.array_equal(dg, dg0)


def test_predicted_tangent_all_targets_and_zero_time_tangents(ctx, oracle):
    """All-pairs shape (what a shared-step dual Hermite or BASELINE configs[4] on predicted bodies
    uses): with every body at the target time and no time tangents the prediction is the
    identity, so the sums must equal the start_bs-shape tangent sums."""
    n, K = 700, 3
    m, q, p, t, h, f, df, tan = _dual_state(oracle, n, K, 7)
    t[:] = 0.25
    g, dg, gt, dgt = ctx.grad_v_dgrad_v_sum_predicted_tangent(
        t, m, q, p, f, df, None, tan["q"], tan["p"], tan["f"], tan["df"], 0.25, None, 1e-4)
    g1, dg1, gt1, dgt1 = ctx.grad_v_dgrad_v_sum_tangent(m, q, p, tan["q"], tan["p"], 1e-4)
    assert np.array_equal(g, g1) and np.array_equal(dg, dg1)
    assert np.array_equal(gt, gt1) and np.array_equal(dgt, dgt1)


def test_predicted_tangent_matches_finite_differences(ctx, oracle):
    """Independent of any dual-number library: central differences of the plain predicted sweep
    along one direction of (t, q, p, f, df, nt)."""
    n, ib = 40, 11
    m, q, p, t, h, f, df, tan = _dual_state(oracle, n, 1, 5)
    nt = t[ib] + h[ib]
    nt_tan = tan["t"][:, ib] + tan["h"][:, ib]
    _, _, gt, dgt = ctx.grad_v_dgrad_v_sum_predicted_tangent(
        t, m, q, p, f, df, tan["t"], tan["q"], tan["p"], tan["f"], tan["df"], nt, nt_tan, 0.0,
        targets=(ib, ib + 1))
    e = 1e-6

    def at(s):
        return ctx.grad_v_dgrad_v_sum_predicted(
            t + s * tan["t"][0], m, q + s * tan["q"][0], p + s * tan["p"][0], f + s * tan["f"][0],
            df + s * tan["df"][0], nt + s * nt_tan[0], 0.0, targets=(ib, ib + 1))
    (gp, dgp), (gm, dgm) = at(e), at(-e)
    assert np.linalg.norm((gp - gm) / (2 * e) - gt[0]) <= 1e-6 * np.linalg.norm(gt[0])
    assert np.linalg.norm((dgp - dgm) / (2 * e) - dgt[0]) <= 1e-6 * np.linalg.norm(dgt[0])


@pytest.mark.parametrize("n,dt,sf,eps", [(3, 0.05, 0.02, 0.0), (8, 0.03125, 0.01, 0.05),
                                         (16, 0.015625, 0.01, 0.05)])
def test_jacobian_advance_matches_oracle_and_is_symplectic(ctx, oracle, n, dt, sf, eps):
    """DFloat_pushforward.jacobian_advance on the GPU path (all 6n unit directions in one pass)
    against the oracle's column-by-column template over Dual: same step count, every column
    <= 1e-9, and omega_pushforward ~ 1 (dFloat_pushforward.ml:141-150)."""
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 20245))
    jac_ref, out_ref = oracle.hermite_jacobian_advance(m, q, p, dt, sf, eps * eps)
    jac, out, nsteps = nbh.dfloat_jacobian_advance(ctx, m, q, p, dt, sf, eps * eps)
    assert nsteps >= n
    assert np.allclose(out, out_ref, rtol=1e-11, atol=1e-13)
    col = np.linalg.norm(jac_ref, axis=0)
    assert np.max(np.linalg.norm(jac - jac_ref, axis=0) / col) <= 1e-9
    om, om_ref = nbh.omega_pushforward(jac), oracle.omega_pushforward(jac_ref)
    assert abs(om - om_ref) <= 1e-9
    assert abs(om - 1.0) <= 5e-3  # 4th-order Hermite is not symplectic: close to 1, not equal


def test_dfloat_hermite_advance_k_directions_equal_single_passes(ctx, oracle):
    """K directions carried at once give what K separate one-direction runs give (they do not
    interact), and the values equal the plain Hermite run's on the host-scheduler path."""
    n, dt, sf = 12, 0.02, 0.01
    m, q, p = oracle.rescale_to_standard_units(*oracle.make_plummer(n, 9))
    rng = np.random.default_rng(3)
    dq, dp = rng.standard_normal((3, n, 3)), rng.standard_normal((3, n, 3)) * m[None, :, None]
    allk = nbh.dfloat_hermite_advance(ctx, m, q, p, dq, dp, dt, sf, 1e-4)
    for k in range(3):
        one = nbh.dfloat_hermite_advance(ctx, m, q, p, dq[k:k + 1], dp[k:k + 1], dt, sf, 1e-4)
        assert one["nsteps"] == allk["nsteps"]
        for key in ("q", "p", "t", "h"):
            assert np.array_equal(one[key], allk[key])
        for key in ("q_tan", "p_tan", "t_tan", "h_tan", "f_tan", "df_tan"):
            a, b = one[key][0], allk[key][k]
            assert np.allclose(a, b, rtol=1e-12, atol=1e-14 * np.abs(b).max()), key


def test_sharded_predicted_tangent_driver_single_rank(ctx, oracle):
    """PeerShardedTangent.load_local_predicted (device-side dual predictor into the peer region)
    + step() against the host-buffer entry point, one rank."""
    import torch
    from nbk.sharded import PeerShardedTangent, make_plan
    n, K = 3000, 4
    m, q, p, t, h, f, df, tan = _dual_state(oracle, n, K, 21)
    nt, nt_tan = 0.03, np.linspace(-1e-3, 1e-3, K)
    sh = PeerShardedTangent(make_plan(n, 1, 0, align=128), K, torch.device("cuda", 0), ctx)
    try:
        sh.load_local_predicted(t, m, q, p, f, df, tan["t"], tan["q"], tan["p"], tan["f"],
                                tan["df"], nt, nt_tan)
        gt, dgt = sh.step(1e-6)
        torch.cuda.synchronize()
        _, _, gt_ref, dgt_ref = ctx.grad_v_dgrad_v_sum_predicted_tangent(
            t, m, q, p, f, df, tan["t"], tan["q"], tan["p"], tan["f"], tan["df"], nt, nt_tan, 1e-6)
        for a, b in ((gt.cpu().numpy(), gt_ref), (dgt.cpu().numpy(), dgt_ref)):
            assert np.max(np.linalg.norm(a - b, axis=-1) / np.linalg.norm(b, axis=-1)) <= 1e-13
    finally:
        sh.close()
