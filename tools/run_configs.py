#!/usr/bin/env python
"""BASELINE.json configs[1], [3], [4] (and [0]) at their stated sizes on one B200, each beside a
bounded run of the oracle: timings and energy errors, one JSON object per config
(-> profiles/r2_configs.json).  configs[2] is bench.py.

usage: tools/run_configs.py [out.json] [--quick]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "direct-nbody_b200")]
import numpy as np  # noqa: E402
import nbh  # noqa: E402
import nbk  # noqa: E402
from oracle import oracle as O  # noqa: E402  (checker only: energy errors beside the GPU's)

quick = "--quick" in sys.argv
only = next((a.split("=", 1)[1].split(",") for a in sys.argv[1:] if a.startswith("--only=")), None)
out_path = next((a for a in sys.argv[1:] if not a.startswith("--")), None)


def want(key):  # --only=0b,1b runs just those sections
    return only is None or key in only


O.set_step_cap(50_000_000)
res = []
with nbk.Context(0) as ctx:
    # ---- configs[0]: Plummer N=1024, Hermite, 1 N-body time unit
    if want("0"):
        m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(1024, 20240))
        e0 = sum(nbh.energy(ctx, m, q, p))
        for eta in (1e-4,) if quick else (1e-4, 1e-5):
            t0 = time.perf_counter()
            a = nbh.hermite_advance(ctx, nbh.HermiteState(m, q, p), 1.0, eta)
            tg = time.perf_counter() - t0
            t0 = time.perf_counter()
            b = O.hermite_advance(O.HermiteState(m, q, p), 1.0, eta)
            tc = time.perf_counter() - t0
            res.append({"config": 0, "what": "Hermite.advance, Plummer N=1024, 1 time unit, device-resident loop",
                        "eta": eta, "steps_gpu": a.nsteps, "steps_oracle": b.nsteps,
                        "dE_gpu": abs((sum(nbh.energy(ctx, a.m, a.q, a.p)) - e0) / e0),
                        "dE_oracle": abs((O.energy(b.m, b.q, b.p) - e0) / e0),
                        "wall_gpu_s": tg, "wall_oracle_1core_s": tc, "us_per_step_gpu": 1e6 * tg / a.nsteps})
            print(res[-1], flush=True)

    # ---- configs[0] as bin/nbody.ml actually wires it: Advancer.A, sf = 1e-4
    if want("0b"):
        na = 256 if quick else 1024
        m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(na, 20240))
        e0 = sum(nbh.energy(ctx, m, q, p))
        # eps = 0.01: with eps = 0 this seed holds a pair whose step drops to 4e-9 and the
        # reference's block recursion then needs millions of interact_bodies calls per 1/256.
        span, eps2a = 1.0 / 16, 1e-4
        nbh.advancer_advance(ctx, nbh.AdvancerState(m, q, p), span / 64, 1e-4, eps2a)  # warm-up (first launch)
        t0 = time.perf_counter()
        a = nbh.advancer_advance(ctx, nbh.AdvancerState(m, q, p), span, 1e-4, eps2a)
        tg = time.perf_counter() - t0
        t0 = time.perf_counter()
        b = O.advancer_advance(O.AdvancerState(m, q, p), span, 1e-4, eps2a)
        tc = time.perf_counter() - t0
        e0 = sum(nbh.energy(ctx, m, q, p, eps2a))
        res.append({"config": "0b", "what": f"Advancer.A.advance, Plummer N={na}, eps=0.01, {span} time units, sf=1e-4 (block recursion on the device, nbk_adv_advance_range)",
                    "interact_bodies_calls": a.stats[0], "pair_interactions": a.stats[1],
                    "same_schedule_as_oracle": a.stats == b.stats,
                    "dE_gpu": abs((sum(nbh.energy(ctx, a.m, a.q, a.p, eps2a)) - e0) / e0),
                    "dE_oracle": abs((O.energy(b.m, b.q, b.p, eps2a) - e0) / e0),
                    "wall_gpu_s": tg, "wall_oracle_1core_s": tc,
                    "us_per_interact_call_gpu": 1e6 * tg / max(a.stats[0], 1)})
        print(res[-1], flush=True)

    # ---- configs[1]: Plummer N=16384, leapfrog (const dt DKD), 1024 steps of 1/1024 = one N-body
    # time unit on the GPU; the oracle (the reference's sequential symmetric kick loop, one core)
    # beside it AT THE SAME N for the first 64 steps: positions and energy error at that point
    if want("1"):
        n = 16384
        m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(n, 20241))
        s = nbh.LeapfrogState(m, q, p)
        s.dt_max[:] = np.inf  # skip the reference's sub-step ladder on fresh bodies
        nst_o, nst = (8, 64) if quick else (64, 1024)
        nbh.leapfrog_advance_const_dt(ctx, s, 1.0 / 1024, 2)  # warm-up (first use of the kernels)
        t0 = time.perf_counter()
        a64 = nbh.leapfrog_advance_const_dt(ctx, s, 1.0 / 1024, nst_o)
        a = nbh.leapfrog_advance_const_dt(ctx, a64, 1.0 / 1024, nst - nst_o)
        tg = time.perf_counter() - t0
        e64 = sum(nbh.energy(ctx, a64.m, a64.q, a64.p))
        ke, pe = nbh.energy(ctx, a.m, a.q, a.p)
        so = O.LeapfrogState(m, q, p)
        so.dt_max[:] = np.inf
        t0 = time.perf_counter()
        bo = O.leapfrog_advance_const_dt(so, 1.0 / 1024, nst_o)
        tc = time.perf_counter() - t0
        res.append({"config": 1, "what": f"Leapfrog.advance_const_dt dt=1/1024 x {nst}, Plummer N=16384",
                    "t_end": float(a.t[0]), "dE_gpu_t_end": abs((ke + pe + 0.25) / 0.25),
                    "ms_per_step_gpu": 1e3 * tg / nst,
                    "oracle_steps_same_n": nst_o,
                    f"dE_gpu_after_{nst_o}": abs((e64 + 0.25) / 0.25),
                    f"dE_oracle_after_{nst_o}": abs((O.energy(bo.m, bo.q, bo.p) + 0.25) / 0.25),
                    f"max_rel_pos_diff_after_{nst_o}": float(np.max(np.linalg.norm(a64.q - bo.q, axis=1) /
                                                                   np.linalg.norm(bo.q, axis=1))),
                    f"max_rel_vel_diff_after_{nst_o}": float(np.max(np.linalg.norm(a64.v - bo.v, axis=1) /
                                                                   np.linalg.norm(bo.v, axis=1))),
                    "ms_per_step_oracle_1core": 1e3 * tc / nst_o})
        print(res[-1], flush=True)

    # ---- configs[1], adaptive: Leapfrog.advance with block time steps, coarse eta; the oracle's
    # sequential-kick run beside it on the first 1024 bodies (statistical agreement: dt_max sees
    # pre-kick velocities on the GPU path)
    if want("1b"):
        m, q, p = nbh.rescale_to_standard_units(ctx, *nbh.make_plummer(16384, 20241))
        span, eta = (1.0 / 1024, 2e-2) if quick else (1.0 / 256, 2e-2)
        nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span / 64, eta)  # warm-up (first launch)
        t0 = time.perf_counter()
        a = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(m, q, p), span, eta)
        tg = time.perf_counter() - t0
        ke, pe = nbh.energy(ctx, a.m, a.q, a.p)
        n_o = 1024
        mo, qo, po = O.rescale_to_standard_units(m[:n_o], q[:n_o], p[:n_o])
        t0 = time.perf_counter()
        bo = O.leapfrog_advance(O.LeapfrogState(mo, qo, po), span, eta)
        tc = time.perf_counter() - t0
        bg = nbh.leapfrog_advance(ctx, nbh.LeapfrogState(mo, qo, po), span, eta)
        res.append({"config": "1b", "what": f"Leapfrog.advance (block steps) span={span:g}, eta={eta:g}, Plummer N=16384",
                    "kick_blocks_n16384": a.nkicks, "wall_gpu_s_n16384": tg,
                    "dE_gpu_n16384": abs((ke + pe + 0.25) / 0.25),
                    "n_oracle": n_o, "kick_blocks_gpu_n1024": bg.nkicks,
                    "dE_gpu_n1024": abs((sum(nbh.energy(ctx, bg.m, bg.q, bg.p)) + 0.25) / 0.25),
                    "dE_oracle_n1024": abs((O.energy(bo.m, bo.q, bo.p) + 0.25) / 0.25),
                    "wall_oracle_1core_s_n1024": tc})
        print(res[-1], flush=True)

    # ---- configs[3]: two-mass Plummer N=65536, Omelyan advancer, per-step energy diagnostics
    if want("3"):
        n = 16384 if quick else 65536
        m, q, p = nbh.make_plummer(n, 20243 + 1)
        m, q, p = nbh.add_mass_disc(m, q, p, 20244)
        m, q, p = nbh.adjust_frame(m, q, p)
        m, q, p = nbh.rescale_to_standard_units(ctx, m, q, p)
        e0 = sum(nbh.energy(ctx, m, q, p))
        # device-resident: bodies, operators, sweeps and diagnostics on the GPU (nbk_oa_*)
        nbh.omelyan_advance_steps(ctx, m, q, p, 0.0, 1.0 / 256, 1)  # warm-up
        t0 = time.perf_counter()
        qr, pr, tr, en_r, el_r = nbh.omelyan_advance_steps(ctx, m, q, p, 0.0, 1.0 / 256, 4, want_el=True)
        t_res = (time.perf_counter() - t0) / 4
        t, errs, t_step, t_diag = 0.0, [], 0.0, 0.0
        for _ in range(4):
            t0 = time.perf_counter()
            q, p, t = nbh.omelyan_advance(ctx, m, q, p, t, 1.0 / 256)
            t_step += time.perf_counter() - t0
            t0 = time.perf_counter()
            e = sum(nbh.energy(ctx, m, q, p))
            el = nbh.bodies_to_el_phase_space(ctx, m, q, p)
            t_diag += time.perf_counter() - t0
            errs.append(abs((e - e0) / e0))
        res.append({"config": 3, "what": f"OA.advance h=1/256 x 4, two-mass Plummer N={n}, energy + per-body E_i each step",
                    "dE_per_step": errs, "ms_per_step_gpu": 1e3 * t_step / 4,
                    "ms_diagnostics_per_step_gpu": 1e3 * t_diag / 4,
                    "ms_per_step_with_diagnostics_device_resident": 1e3 * t_res,
                    "device_resident_bit_identical": bool(np.array_equal(qr, q) and np.array_equal(pr, p)
                                                          and np.array_equal(el_r, el) and en_r[-1] == e),
                    "sum_Ei_minus_KE_minus_2PE": float(abs(el[:, 0].sum() - (e + nbh.energy(ctx, m, q, p)[1])))})
        print(res[-1], flush=True)

    # ---- configs[4]: tangent (variational) forces, N=32768
    if want("4"):
        n = 8192 if quick else 32768
        m, q, p = nbh.make_plummer(n, 20245)
        rng = np.random.default_rng(20245)
        for K in (1, 6):
            dq = rng.standard_normal((K, n, 3))
            dp = rng.standard_normal((K, n, 3)) * m[None, :, None]
            ctx.grad_v_dgrad_v_sum_tangent(m, q, p, dq, dp, 0.0)
            t0 = time.perf_counter()
            g, dg, gt, dgt = ctx.grad_v_dgrad_v_sum_tangent(m, q, p, dq, dp, 0.0)
            tg = time.perf_counter() - t0
            sl = (1000, 1016)
            rg, rdg, rgt, rdgt = O.grad_v_dgrad_v_sum_tangent(m, q, p, dq[0], dp[0], 0.0, targets=sl)
            err = float(np.max(np.linalg.norm(dgt[0][sl[0]:sl[1]] - rdgt, axis=1) / np.linalg.norm(rdgt, axis=1)))
            res.append({"config": 4, "what": f"acc+jerk tangent sums, K={K} directions, Plummer N={n} (host-buffer call)",
                        "K": K, "wall_s": tg, "tangent_interactions_per_s": K * n * (n - 1.0) / tg,
                        "max_rel_err_vs_dual_oracle_16_targets": err})
            print(res[-1], flush=True)
if out_path:
    json.dump(res, open(out_path, "w"), indent=1)
