"""CPU: the OCaml side of the drop-in (ocaml/*.ml), EXECUTED.

Nothing here can compile OCaml, and the reference tree does not travel to the GPU box, so the
maintainer-side modules (Hermite_gpu, Omelyan_gpu, Leapfrog_gpu, Advancer_gpu, Energy_gpu,
Analysis_gpu, Dfloat_hermite_gpu) are run the only way this image allows: by the interpreter oracle/miniml, next to
the UNMODIFIED reference modules they replace, with every `external` of ocaml/nbk.ml bound to a
Python stand-in that does what include/nbk.h says the C entry point does.  The stand-ins are
built from the reference's own functions (Base.grad_v, Leapfrog.kick, A.interact_bodies,
A.finish_body, ... called through the interpreter), i.e. they are an executable reading of
nbk.h's contract "phase X of the reference, on the resident rows".

What this checks is the ORCHESTRATION the .ml files add - packing, ranges, signs, the resident
protocols (pending updates, permutations returned by the device, hmax / dt_max bookkeeping on
the host), the block recursions walked around per-phase calls: with stand-ins that perform the
reference's phases exactly, Leapfrog_gpu.advance and Advancer_gpu.advance must return the
reference's result BIT FOR BIT, and the sweeps-based ones (different summation order) to 1e-12.
What it does not check: the C stubs (ocaml/nbk_stubs.c) and OCaml typing.  The kernels behind the
same entry points are checked on the GPU against the oracle (tests/test_*_gpu.py)."""
import math
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_fixtures"))
from oracle.miniml import Builtin, Ctor, Interp, Record, call, from_list, run_deep, to_list  # noqa: E402

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="needs the reference tree")
UNIT = ()
INF = math.inf
NONE_ = Ctor("None")


def bodies(name, n=None):
    import make_inputs
    m, q, p = make_inputs.read_bodies(os.path.join(ROOT, "oracle", "ref_fixtures", "inputs", name))
    n = n or len(m)
    return [(float(m[i]), [float(x) for x in q[i]], [float(x) for x in p[i]]) for i in range(n)]


class World:
    """An interpreter with the reference, ocaml/ and the stand-ins for ocaml/nbk.ml's externals."""

    def __init__(self, hermite_resident_max=0, adv_range_max=0, lf_subsystem_max=0):
        it = self.it = Interp(search_path=[REF, os.path.join(ROOT, "ocaml"),
                                           os.path.join(ROOT, "oracle", "ref_fixtures", "diff_restated")],
                              stdout=open(os.devnull, "w"), stderr=open(os.devnull, "w"))
        self.limits = dict(hermite_resident_max=hermite_resident_max, adv_range_max=adv_range_max,
                           lf_subsystem_max=lf_subsystem_max)
        self.calls = {}
        for f in ("base", "hermite", "leapfrog", "advancer", "omelyan_advancer", "energy", "analysis",
                  "dFloat", "dFloat_advancer", "dFloat_hermite"):
            it.load_file(os.path.join(REF, f + ".ml"))
        self.An = it.load_source("module An = Analysis.Make(Advancer.A)", "AnHolder").mods["An"]
        self.register()
        for f in ("nbk", "hermite_gpu", "omelyan_gpu", "leapfrog_gpu", "advancer_gpu", "energy_gpu",
                  "analysis_gpu", "dfloat_hermite_gpu"):
            it.load_file(os.path.join(ROOT, "ocaml", f + ".ml"))
        # Analysis.load_state(s) and the stale half of dFloat_advancer.ml: not used here
        assert it.missing <= {"input_value", "zero"}, it.missing

    # -- access to interpreted values
    def fn(self, path):
        *mods, name = path.split(".")
        m = self.it.root.mods[mods[0]]
        for sub in mods[1:]:
            m = m.mods[sub]
        return m.vals[name][0]

    def run(self, path, *args):
        return run_deep(call, self.fn(path), *args)

    def set_eps2(self, eps2):
        self.fn("Base.eps2").f["contents"] = eps2

    # -- the stand-ins ---------------------------------------------------------------------------
    def register(self):
        def ext(names, arity):
            def deco(f):
                def counted(*a):
                    self.calls[names[0]] = self.calls.get(names[0], 0) + 1
                    return f(*a)
                for nm in names:
                    self.it.externals[nm] = Builtin(counted, arity, nm)
                return f
            return deco

        base = lambda name: self.it.root.mods["Base"].vals[name][0]  # noqa: E731
        st = self.state = {}

        @ext(["nbk_ml_create"], 1)
        def _create(dev):
            return {"device": dev}

        for prim, key in (("nbk_ml_hermite_resident_max", "hermite_resident_max"),
                          ("nbk_ml_adv_range_max", "adv_range_max"),
                          ("nbk_ml_lf_subsystem_max", "lf_subsystem_max")):
            ext([prim], 1)(lambda u, key=key: self.limits[key])

        # ---- host-buffer sweeps: per target, sources in index order (include/nbk.h)
        @ext(["nbk_ml_grad_v_sum", "nbk_ml_grad_v_sum_bc"], 6)
        def _grad_v_sum(ctx, m, q, eps2, rng, out):
            assert eps2 == self.fn("Base.eps2").f["contents"]
            t0, t1, s0, s1 = rng
            gv = [0.0] * 3
            for i in range(t0, t1):
                acc = [0.0] * 3
                for j in range(s0, s1):
                    if j != i:
                        call(base("grad_v"), m[i], q[3 * i:3 * i + 3], m[j], q[3 * j:3 * j + 3], gv)
                        for k in range(3):
                            acc[k] += gv[k]
                out[3 * (i - t0):3 * (i - t0) + 3] = acc
            return UNIT

        @ext(["nbk_ml_grad_v_dgrad_v_sum", "nbk_ml_grad_v_dgrad_v_sum_bc"], 8)
        def _gvdgv_sum(ctx, m, q, p, eps2, rng, g, dg):
            t0, t1, s0, s1 = rng
            gv, dgv = [0.0] * 3, [0.0] * 3
            for i in range(t0, t1):
                a, b = [0.0] * 3, [0.0] * 3
                for j in range(s0, s1):
                    if j != i:
                        call(base("grad_v_dgrad_v"), m[i], q[3 * i:3 * i + 3], p[3 * i:3 * i + 3],
                             m[j], q[3 * j:3 * j + 3], p[3 * j:3 * j + 3], gv, dgv)
                        for k in range(3):
                            a[k] += gv[k]
                            b[k] += dgv[k]
                g[3 * (i - t0):3 * (i - t0) + 3] = a
                dg[3 * (i - t0):3 * (i - t0) + 3] = b
            return UNIT

        @ext(["nbk_ml_interact_sum", "nbk_ml_interact_sum_bc"], 6)
        def _interact_sum(ctx, m, q, na, eps2, s):
            # active rows [0, na) against everybody, passive rows against the active ones
            n = len(m)
            _grad_v_sum(ctx, m, q, eps2, (0, na, 0, n), s)
            tail = [0.0] * (3 * (n - na))
            _grad_v_sum(ctx, m, q, eps2, (na, n, 0, na), tail)
            s[3 * na:] = tail
            return UNIT

        # ---- energies in the reference's own summation order (energy.ml:57-78)
        def kinetic(m, p):
            e = 0.0
            for i in range(len(m)):
                x, y, z = p[3 * i:3 * i + 3]
                e = e + (x * x + y * y + z * z) / (2.0 * m[i])
            return e

        def potential(m, q):
            pe = 0.0
            for i in range(len(m)):
                for j in range(i + 1, len(m)):
                    pe = pe + call(base("v"), m[i], q[3 * i:3 * i + 3], m[j], q[3 * j:3 * j + 3])
            return pe
        ext(["nbk_ml_kinetic_energy"], 3)(lambda ctx, m, p: kinetic(m, p))
        ext(["nbk_ml_potential_energy"], 4)(lambda ctx, m, q, eps2: potential(m, q))
        ext(["nbk_ml_energy"], 5)(lambda ctx, m, q, p, eps2: (kinetic(m, p), potential(m, q)))

        # ---- resident Hermite state (nbk_hermite_upload / nbk_hermite_body_sum)
        @ext(["nbk_ml_hermite_upload", "nbk_ml_hermite_upload_bc"], 8)
        def _hermite_upload(ctx, t, m, q, p, h, f, df):
            n = len(m)
            st["hermite"] = [Record(dict(id=i, t=t[i], m=m[i], q=q[3 * i:3 * i + 3], p=p[3 * i:3 * i + 3],
                                         h=h[i], f=f[3 * i:3 * i + 3], df=df[3 * i:3 * i + 3]))
                             for i in range(n)]
            return UNIT

        @ext(["nbk_ml_hermite_body_sum", "nbk_ml_hermite_body_sum_bc"], 8)
        def _hermite_body_sum(ctx, i, nt, eps2, upd, us, g, dg):
            hs = st["hermite"]
            if upd >= 0:
                hs[upd].f.update(t=us[0], q=us[1:4], p=us[4:7], f=us[7:10], df=us[10:13])
            predq, predp = self.fn("Hermite.predict_position_into"), self.fn("Hermite.predict_momentum_into")
            qp = call(predq, [0.0] * 3, hs[i], nt)
            pp = call(predp, [0.0] * 3, hs[i], nt)
            a, b, gv, dgv = [0.0] * 3, [0.0] * 3, [0.0] * 3, [0.0] * 3
            for j, bj in enumerate(hs):
                if j != i:
                    call(base("grad_v_dgrad_v"), hs[i].f["m"], qp, pp, bj.f["m"],
                         call(predq, [0.0] * 3, bj, nt), call(predp, [0.0] * 3, bj, nt), gv, dgv)
                    for k in range(3):
                        a[k] += gv[k]
                        b[k] += dgv[k]
            g[:], dg[:] = a, b
            return UNIT

        @ext(["nbk_ml_hermite_advance_resident"], 5)
        def _hermite_advance_resident(ctx, stop_time, eta, eps2, max_steps):
            # hermite.ml:171-178 + finish_bs on the resident rows; returns the advance_body calls made
            hs = st["hermite"]
            before = self.fn("Base.nsteps").f["contents"]
            lst = call(self.fn("List.fast_sort"), self.fn("Hermite.compare_next_time"), from_list(hs))
            out = call(self.fn("Hermite.advance_bodies"), eta, stop_time, lst)   # sorted by id = row
            st["hermite"] = list(out)
            steps = self.fn("Base.nsteps").f["contents"] - before
            self.fn("Base.nsteps").f["contents"] = before    # the caller adds them (hermite_gpu.ml)
            return steps

        @ext(["nbk_ml_hermite_download", "nbk_ml_hermite_download_bc"], 7)
        def _hermite_download(ctx, t, q, p, h, f, df):
            for i, b in enumerate(st["hermite"]):
                t[i], h[i] = b.f["t"], b.f["h"]
                for name, dst in (("q", q), ("p", p), ("f", f), ("df", df)):
                    dst[3 * i:3 * i + 3] = b.f[name]
            return UNIT

        # ---- resident Advancer.A records (nbk_adv_*): each phase IS the reference's phase
        A = lambda name: self.fn("Advancer.A." + name)  # noqa: E731
        FIELDS = (("t", 1), ("m", 1), ("q", 3), ("p", 3), ("h", 1), ("hmax", 1), ("t0", 1), ("dVdq0", 3),
                  ("dVdq1", 3), ("dVdq2", 3), ("p0", 3), ("q0", 3), ("q1", 3), ("q2", 3), ("cs", 3))

        def to_row(b):
            row = []
            for name, w in FIELDS:
                row.extend([b.f[name]] if w == 1 else b.f[name])
            assert len(row) == 35
            return row

        @ext(["nbk_ml_adv_upload"], 2)
        def _adv_upload(ctx, r):
            rows = []
            for i in range(len(r) // 35):
                b = call(A("make_body_id"), i, 0.0, 1.0, [0.0] * 3, [0.0] * 3)
                o = 35 * i
                for name, w in FIELDS:
                    b.f[name] = r[o] if w == 1 else r[o:o + w]
                    o += w
                rows.append(b)
            st["adv"] = rows
            return UNIT

        @ext(["nbk_ml_adv_download"], 4)
        def _adv_download(ctx, i0, i1, r):
            out = []
            for b in st["adv"][i0:i1]:
                out.extend(to_row(b))
            r[:len(out)] = out
            return UNIT

        @ext(["nbk_ml_adv_permute"], 2)
        def _adv_permute(ctx, perm):
            old = list(st["adv"])
            for s, o in enumerate(perm):
                st["adv"][s] = old[o]
            return UNIT

        @ext(["nbk_ml_adv_begin_step"], 4)
        def _adv_begin(ctx, imin, imax, hnew):
            for b in st["adv"][imin:imax]:
                call(A("predict_q012"), b, hnew)
                call(A("clear_before_step"), b)
            return UNIT

        @ext(["nbk_ml_adv_interact", "nbk_ml_adv_interact_bc"], 6)
        def _adv_interact(ctx, imin, imax, t, vC, eps2):
            call(A("interact_bodies"), st["adv"], imin, imax, t, vC)
            return UNIT

        @ext(["nbk_ml_adv_finish"], 5)
        def _adv_finish(ctx, imin, imax, t, r):
            out = []
            for b in st["adv"][imin:imax]:
                call(A("finish_body"), b, t)
                out.extend(to_row(b))
            r[:len(out)] = out
            return UNIT
        ext(["nbk_ml_adv_get_t"], 2)(lambda ctx, i: st["adv"][i].f["t"])

        @ext(["nbk_ml_adv_init_first"], 2)
        def _adv_init_first(ctx, eps2):
            # advancer.ml:357-393 (the sums); the time steps it also assigns are the host's business
            call(A("initialize_before_first_step"), st["adv"], 1.0)
            return UNIT

        @ext(["nbk_ml_adv_advance_range", "nbk_ml_adv_advance_range_bc"], 6)
        def _adv_advance_range(ctx, dt, sf, eps2, hmax, perm):
            # advance_range bs imax dt sf None (advancer.ml:316-350) for the prefix, in place
            imax = len(hmax)
            rows = st["adv"]
            old = list(rows[:imax])
            for b, h in zip(old, hmax):
                b.f["hmax"] = h
            call(A("advance_range"), rows, imax, dt, sf, NONE_)   # extpot = None
            for s_ in range(imax):
                hmax[s_] = rows[s_].f["hmax"]
                perm[s_] = next(k for k, b in enumerate(old) if b is rows[s_])
            return UNIT

        # ---- resident Leapfrog.b rows (nbk_lf_*)
        Lf = lambda name: self.fn("Leapfrog." + name)  # noqa: E731

        @ext(["nbk_ml_lf_upload"], 5)
        def _lf_upload(ctx, dtm, m, q, v):
            st["lf"] = [Record(dict(id=i, t=0.0, dt_max=dtm[i], m=m[i], q=q[3 * i:3 * i + 3],
                                    v=v[3 * i:3 * i + 3])) for i in range(len(m))]
            return UNIT

        @ext(["nbk_ml_lf_download"], 4)
        def _lf_download(ctx, dtm, q, v):
            for i, b in enumerate(st["lf"]):
                dtm[i] = b.f["dt_max"]
                q[3 * i:3 * i + 3] = b.f["q"]
                v[3 * i:3 * i + 3] = b.f["v"]
            return UNIT

        @ext(["nbk_ml_lf_drift"], 4)
        def _lf_drift(ctx, i0, i1, dt):
            for b in st["lf"][i0:i1]:
                call(Lf("drift"), dt, b)
            return UNIT

        @ext(["nbk_ml_lf_kick_block", "nbk_ml_lf_kick_block_bc"], 8)
        def _lf_kick_block(ctx, iend, ibegin, eta, dt, with_tscale, drift_after, dtm):
            rows = st["lf"]
            for b in rows[ibegin:iend]:
                b.f["dt_max"] = INF
            for i in range(ibegin, iend):          # leapfrog.ml:132-136, sequentially
                for j in range(i):
                    call(Lf("kick"), eta, dt, rows[i], rows[j])
            for b in rows[ibegin:iend]:
                call(Lf("drift"), drift_after, b)
            for i in range(iend):
                dtm[i] = rows[i].f["dt_max"]
            return UNIT

        @ext(["nbk_ml_lf_sort_sub"], 2)
        def _lf_sort_sub(ctx, perm):
            iend = len(perm)
            keyed = from_list(list(range(iend)))
            cmp = Builtin(lambda a, b: call(self.fn("Pervasives.compare"), st["lf"][a].f["dt_max"],
                                            st["lf"][b].f["dt_max"]), 2)
            order = to_list(call(self.fn("List.stable_sort"), cmp, keyed))
            old = list(st["lf"])
            for s, o in enumerate(order):
                st["lf"][s] = old[o]
                perm[s] = o
            return UNIT

        @ext(["nbk_ml_lf_advance_subsystem", "nbk_ml_lf_advance_subsystem_bc"], 6)
        def _lf_advance_subsystem(ctx, dt, eta, t, dtm, perm):
            # advance_subsystem dt eta bs iend (leapfrog.ml:119-142) for the prefix, in place
            iend = len(t)
            rows = st["lf"]
            old = list(rows[:iend])
            for b, ti in zip(old, t):
                b.f["t"] = ti
            call(Lf("advance_subsystem"), dt, eta, rows, iend)
            for s_ in range(iend):
                t[s_], dtm[s_] = rows[s_].f["t"], rows[s_].f["dt_max"]
                perm[s_] = next(k for k, b in enumerate(old) if b is rows[s_])
            return UNIT

        @ext(["nbk_ml_kick_sum", "nbk_ml_kick_sum_bc"], 9)
        def _kick_sum(ctx, m, q, v, eta, dt, rng, dv, tm):
            # per target: every pair from the velocities at sweep start, min of the pair time-scales
            t0, t1, s0, s1 = rng
            mk = lambda i: Record(dict(id=i, t=0.0, dt_max=INF, m=m[i], q=q[3 * i:3 * i + 3],  # noqa: E731
                                       v=v[3 * i:3 * i + 3]))
            for i in range(t0, t1):
                tmin, acc = INF, [0.0] * 3
                for j in range(s0, s1):
                    if j != i:
                        bi, bj = mk(i), mk(j)
                        call(Lf("kick"), eta, dt, bi, bj)
                        tmin = min(tmin, bi.f["dt_max"])
                        for k in range(3):
                            acc[k] += bi.f["v"][k] - v[3 * i + k]
                tm[i - t0] = tmin
                dv[3 * (i - t0):3 * (i - t0) + 3] = acc
            return UNIT

        # ---- dual-number sums (nbk_grad_v_dgrad_v_sum_tangent / _predicted_tangent): K directions
        # through the reference's DFloatBase / DFloat_hermite, one direction at a time
        def D(v, d):
            return Ctor("D", (v, d))

        def dual3(vals, tans, i):
            return [D(vals[3 * i + c], tans[3 * i + c]) for c in range(3)]

        def parts(x):
            return (x.arg, 0.0) if x.name == "C" else x.arg

        @ext(["nbk_ml_tangent", "nbk_ml_tangent_bc"], 11)
        def _tangent(ctx, m, q, p, K, q_tan, p_tan, eps2, rng, gt, dgt):
            t0, t1, s0, s1 = rng
            n, nt_ = len(m), t1 - t0
            gvd = self.fn("DFloat_advancer.DFloatBase.grad_v_dgrad_v")
            for k in range(K):
                qt, pt = q_tan[3 * n * k:3 * n * (k + 1)], p_tan[3 * n * k:3 * n * (k + 1)]
                for i in range(t0, t1):
                    a, b = [0.0] * 3, [0.0] * 3
                    for j in range(s0, s1):
                        if j == i:
                            continue
                        gv, dgv = [Ctor("C", 0.0)] * 3, [Ctor("C", 0.0)] * 3
                        call(gvd, Ctor("C", m[i]), dual3(q, qt, i), dual3(p, pt, i),
                             Ctor("C", m[j]), dual3(q, qt, j), dual3(p, pt, j), gv, dgv)
                        for c in range(3):
                            a[c] += parts(gv[c])[1]
                            b[c] += parts(dgv[c])[1]
                    o = 3 * nt_ * k + 3 * (i - t0)
                    gt[o:o + 3], dgt[o:o + 3] = a, b
            return UNIT

        @ext(["nbk_ml_predicted_tangent", "nbk_ml_predicted_tangent_bc"], 9)
        def _predicted_tangent(ctx, st_, K, tans, nt, nt_tan, eps2, rng, out):
            t, m, q, p, f, df = st_
            t_tan, q_tan, p_tan, f_tan, df_tan = tans
            g, dg, gt, dgt = out
            t0, t1, s0, s1 = rng
            n = len(m)
            predq = self.fn("DFloat_hermite.predict_position_into")
            predp = self.fn("DFloat_hermite.predict_momentum_into")
            gvd = self.fn("DFloat_advancer.DFloatBase.grad_v_dgrad_v")
            C0 = Ctor("C", 0.0)
            for k in range(K):
                def body(i):
                    sl3 = slice(3 * n * k, 3 * n * (k + 1))
                    return Record(dict(id=i, t=D(t[i], t_tan[n * k + i]), m=Ctor("C", m[i]),
                                       q=dual3(q, q_tan[sl3], i), p=dual3(p, p_tan[sl3], i), h=C0,
                                       f=dual3(f, f_tan[sl3], i), df=dual3(df, df_tan[sl3], i)))
                ntd = D(nt, nt_tan[k])
                for i in range(t0, t1):
                    bi = body(i)
                    qp, pp = call(predq, [C0] * 3, bi, ntd), call(predp, [C0] * 3, bi, ntd)
                    av, bv, at, bt = [0.0] * 3, [0.0] * 3, [0.0] * 3, [0.0] * 3
                    for j in range(s0, s1):
                        if j == i:
                            continue
                        bj = body(j)
                        gv, dgv = [C0] * 3, [C0] * 3
                        call(gvd, bi.f["m"], qp, pp, bj.f["m"], call(predq, [C0] * 3, bj, ntd),
                             call(predp, [C0] * 3, bj, ntd), gv, dgv)
                        for c in range(3):
                            av[c] += parts(gv[c])[0]
                            bv[c] += parts(dgv[c])[0]
                            at[c] += parts(gv[c])[1]
                            bt[c] += parts(dgv[c])[1]
                    nt_ = t1 - t0
                    o = 3 * nt_ * k + 3 * (i - t0)
                    gt[o:o + 3], dgt[o:o + 3] = at, bt
                    if k == 0:
                        g[3 * (i - t0):3 * (i - t0) + 3] = av
                        dg[3 * (i - t0):3 * (i - t0) + 3] = bv
            return UNIT

        # ---- Analysis: binaries scan / list and the E-L diagnostics of the resident OA rows
        def abody(m, q, p, i):
            return call(A("make_body_id"), i, 0.0, m[i], q[3 * i:3 * i + 3], p[3 * i:3 * i + 3])

        def pair_energies(m, q, p):
            bs = [abody(m, q, p, i) for i in range(len(m))]
            be = self.An.vals["binary_energy"][0]
            return [(i, j, call(be, bs[i], bs[j])) for i in range(len(m)) for j in range(i + 1, len(m))]

        @ext(["nbk_ml_binary_scan", "nbk_ml_binary_scan_bc"], 10)
        def _binary_scan(ctx, m, q, p, eps2, thr, rng, emin, partner, count):
            n = len(m)
            emin[:], partner[:], count[:] = [INF] * n, [-1] * n, [0] * n
            for i, j, e in pair_energies(m, q, p):
                for a, b in ((i, j), (j, i)):
                    if e < emin[a]:
                        emin[a], partner[a] = e, b
                    if e < thr:
                        count[a] += 1
            return UNIT

        @ext(["nbk_ml_binary_list", "nbk_ml_binary_list_bc"], 9)
        def _binary_list(ctx, m, q, p, eps2, thr, pi, pj, e):
            k = 0
            for i, j, en in pair_energies(m, q, p):
                if en < thr:
                    pi[k], pj[k], e[k] = i, j, en
                    k += 1
            return k

        @ext(["nbk_ml_oa_upload"], 4)
        def _oa_upload(ctx, m, q, p):
            st["oa"] = [abody(m, q, p, i) for i in range(len(m))]
            return UNIT

        @ext(["nbk_ml_oa_diagnostics"], 3)
        def _oa_diag(ctx, eps2, el):
            rows = call(self.An.vals["bodies_to_el_phase_space"][0], st["oa"])
            for i, r in enumerate(rows):
                el[4 * i:4 * i + 4] = r
            return (0.0, 0.0)


def rel(a, b):
    num = math.sqrt(sum((x - y) ** 2 for x, y in zip(a, b)))
    den = math.sqrt(sum(y * y for y in b))
    return num / den if den else num


@pytest.fixture(scope="module")
def world():
    return World()


@pytest.fixture(scope="module")
def world_subtrees():
    """The same modules with the device-side limits a GPU reports being non-zero: narrow subtrees
    of the block recursions (and the whole Hermite loop) become ONE call each."""
    return World(hermite_resident_max=12288, adv_range_max=7, lf_subsystem_max=5)


def hermite_bodies(w, bs):
    return [w.run("Hermite.make", 0.0, m, q, p) for m, q, p in bs]


def test_hermite_start_bs_and_advance_through_the_resident_protocol(world):
    w = world
    w.set_eps2(0.0)
    small = bodies("plummer_n16_std_seed20241.txt")
    ref = to_list(w.run("Hermite.start_bs", 1e-2, from_list(hermite_bodies(w, small))))
    got = to_list(w.run("Hermite_gpu.start_bs", 1e-2, from_list(hermite_bodies(w, small))))
    for r, g in zip(ref, got):
        assert rel(g.f["f"], r.f["f"]) < 1e-13 and rel(g.f["df"], r.f["df"]) < 1e-13
        assert abs(g.f["h"] - r.f["h"]) <= 1e-13 * abs(r.f["h"])
    # Hermite.advance: upload once, host scheduler, one nbk_hermite_body_sum per advance_body with the
    # previous step's state riding along (hermite_resident_max = 0 forces this path)
    w.fn("Base.nsteps").f["contents"] = 0
    ref = w.run("Hermite.advance", hermite_bodies(w, small), 0.0625, 1e-3)
    nref = w.fn("Base.nsteps").f["contents"]
    w.fn("Base.nsteps").f["contents"] = 0
    w.calls.clear()
    got = w.run("Hermite_gpu.advance", hermite_bodies(w, small), 0.0625, 1e-3)
    assert w.fn("Base.nsteps").f["contents"] == nref > 30
    assert w.calls["nbk_ml_hermite_upload"] == 1 and w.calls["nbk_ml_hermite_body_sum"] == nref
    for r, g in zip(ref, got):
        assert g.f["t"] == r.f["t"] == 0.0625
        for k in ("q", "p", "f", "df"):
            assert rel(g.f[k], r.f[k]) < 1e-11, k
        assert abs(g.f["h"] - r.f["h"]) <= 1e-10 * abs(r.f["h"])


def test_hermite_advance_whole_loop_in_one_call(world_subtrees):
    """n <= Nbk.hermite_resident_max (): upload, ONE nbk_hermite_advance_resident, download."""
    w = world_subtrees
    w.set_eps2(0.0)
    small = bodies("plummer_n16_std_seed20241.txt")
    w.fn("Base.nsteps").f["contents"] = 0
    ref = w.run("Hermite.advance", hermite_bodies(w, small), 0.0625, 1e-3)
    nref = w.fn("Base.nsteps").f["contents"]
    w.fn("Base.nsteps").f["contents"] = 0
    w.calls.clear()
    got = w.run("Hermite_gpu.advance", hermite_bodies(w, small), 0.0625, 1e-3)
    assert w.fn("Base.nsteps").f["contents"] == nref
    assert w.calls["nbk_ml_hermite_advance_resident"] == 1 and "nbk_ml_hermite_body_sum" not in w.calls
    for r, g in zip(ref, got):
        for k in ("t", "h"):
            assert abs(g.f[k] - r.f[k]) <= 1e-10 * abs(r.f[k])
        for k in ("q", "p", "f", "df"):
            assert rel(g.f[k], r.f[k]) < 1e-11, k


def test_omelyan_operators(world):
    w = world
    w.set_eps2(0.2 * 0.2)
    small = bodies("plummer_n16_std_seed20241.txt")
    mk = lambda: from_list([w.run("Omelyan_advancer.OA.make", 0.0, m, q, p) for m, q, p in small])  # noqa: E731
    for op in ("b_operate", "c_operate"):
        ref, got = mk(), mk()
        w.run("Omelyan_advancer.OA." + op, 1.0 / 64, ref)
        w.run("Omelyan_gpu." + op, 1.0 / 64, got)
        for r, g in zip(to_list(ref), to_list(got)):
            assert rel(g.f["p"], r.f["p"]) < 1e-14 and rel(g.f["q_star"], r.f["q_star"]) < 1e-14, op
            assert g.f["q"] == r.f["q"]
    w.set_eps2(0.0)


def test_advancer_interact_bodies_level_1(world):
    w = world
    w.set_eps2(0.0)
    small = bodies("plummer_n16_std_seed20241.txt")

    def prepared():
        bs = [w.run("Advancer.A.make", 0.0, m, q, p) for m, q, p in small]
        w.run("Advancer.A.initialize_before_first_step", bs, 1e-3)
        for b in bs:
            w.run("Advancer.A.predict_q012", b, 1.0 / 64)
            w.run("Advancer.A.clear_before_step", b)
        return bs
    for imin, imax in ((0, 16), (5, 16), (3, 9)):
        ref, got = prepared(), prepared()
        w.run("Advancer.A.interact_bodies", ref, imin, imax, 1.0 / 128, 1.0 / 384)
        w.run("Advancer_gpu.interact_bodies", got, imin, imax, 1.0 / 128, 1.0 / 384)
        for r, g in zip(ref, got):
            assert g.f["q"] == r.f["q"] and g.f["cs"] == r.f["cs"]
            for k in ("dVdq0", "dVdq1", "dVdq2"):
                assert rel(g.f[k], r.f[k]) < 1e-13 or max(map(abs, r.f[k])) == 0.0, (imin, imax, k)


@pytest.mark.parametrize("subtrees", [False, True])
@pytest.mark.parametrize("eps", [0.0, 0.2])
def test_advancer_advance_level_2_is_the_reference_bit_for_bit(world, world_subtrees, eps, subtrees):
    """Advancer_gpu.advance: records resident, the recursion walked on the host around
    nbk_adv_begin_step / interact / finish / permute (adv_range_max = 0), or with every subtree of
    up to 7 rows handed to nbk_adv_advance_range; first call and resumed."""
    w = world_subtrees if subtrees else world
    w.set_eps2(eps * eps)
    small = bodies("plummer_n16_std_seed20241.txt")
    mk = lambda: [w.run("Advancer.A.make", 0.0, m, q, p) for m, q, p in small]  # noqa: E731
    ref = w.run("Advancer.A.advance", mk(), 0.0625, 1e-3)
    ref2 = w.run("Advancer.A.advance", ref, 0.0625, 1e-3)
    w.calls.clear()
    got = w.run("Advancer_gpu.advance", mk(), 0.0625, 1e-3)
    assert w.calls["nbk_ml_adv_upload"] == 1 and w.calls["nbk_ml_adv_init_first"] == 1
    assert w.calls["nbk_ml_adv_interact"] == 3 * w.calls["nbk_ml_adv_begin_step"] > (3 if subtrees else 30)
    assert (w.calls.get("nbk_ml_adv_advance_range", 0) > 0) == subtrees
    got2 = w.run("Advancer_gpu.advance", got, 0.0625, 1e-3)
    for a, b in ((ref, got), (ref2, got2)):
        assert [x.f["m"] for x in a] == [x.f["m"] for x in b]
        for r, g in zip(a, b):
            for k in ("t", "h", "hmax", "t0", "q", "p", "dVdq0", "dVdq1", "dVdq2", "q0", "q1", "q2", "p0", "cs"):
                assert g.f[k] == r.f[k], k
    w.set_eps2(0.0)


@pytest.mark.parametrize("subtrees", [False, True])
def test_leapfrog_advance_resident_is_the_reference_bit_for_bit(world, world_subtrees, subtrees):
    """Leapfrog_gpu.advance: rows resident, advance_subsystem walked on the host around
    nbk_lf_drift / lf_kick_block / lf_sort_sub (lf_subsystem_max = 0), or with every subtree of up
    to 5 rows handed to nbk_lf_advance_subsystem."""
    w = world_subtrees if subtrees else world
    w.set_eps2(0.0)
    small = bodies("plummer_n16_std_seed20241.txt")
    mk = lambda: [w.run("Leapfrog.make", 0.0, m, q, p) for m, q, p in small]  # noqa: E731
    ref = w.run("Leapfrog.advance", mk(), 1.0 / 64, 1e-2)
    w.calls.clear()
    got = w.run("Leapfrog_gpu.advance", mk(), 1.0 / 64, 1e-2)
    assert w.calls["nbk_ml_lf_upload"] == 1 and w.calls["nbk_ml_lf_download"] == 1
    assert w.calls["nbk_ml_lf_kick_block"] == w.calls["nbk_ml_lf_sort_sub"] > (0 if subtrees else 10)
    assert (w.calls.get("nbk_ml_lf_advance_subsystem", 0) > 0) == subtrees
    assert [b.f["m"] for b in got] == [b.f["m"] for b in ref]
    ids = lambda bs: [b.f["id"] - min(x.f["id"] for x in bs) for b in bs]  # noqa: E731
    assert ids(got) == ids(ref)
    for r, g in zip(ref, got):
        for k in ("t", "dt_max", "q", "v"):
            assert g.f[k] == r.f[k], k


def test_energy_and_analysis_modules(world):
    w = world
    w.set_eps2(0.0)
    big = bodies("plummer_n64_seed20240.txt")
    ab = [w.run("Advancer.A.make", 0.0, m, q, p) for m, q, p in big]
    mods = w.it.load_source("""
module ER = Energy.Make (Advancer.A)
module EG = Energy_gpu.Make (Advancer.A)
module AR = Analysis.Make (Advancer.A)
module AG = Analysis_gpu.Make (Advancer.A)
""", "Holders").mods
    for f in ("total_kinetic_energy", "total_potential_energy", "energy"):
        assert call(mods["EG"].vals[f][0], ab) == call(mods["ER"].vals[f][0], ab), f
    # kinetic_energy / potential_energy of single bodies are inherited through `include`
    assert call(mods["EG"].vals["kinetic_energy"][0], ab[3]) == call(mods["ER"].vals["kinetic_energy"][0], ab[3])
    ident = lambda b: next(i for i, x in enumerate(ab) if x is b)  # noqa: E731
    rb = to_list(call(mods["AR"].vals["binaries"][0], ab))
    gb = to_list(call(mods["AG"].vals["binaries"][0], ab))
    assert len(rb) == len(gb) > 10
    assert [(ident(a), ident(b), e) for (a, b), e in rb] == [(ident(a), ident(b), e) for (a, b), e in gb]
    rt = call(mods["AR"].vals["binaries_threshold"][0], 1e-4, ab)
    gt = call(mods["AG"].vals["binaries_threshold"][0], 1e-4, ab)
    assert 2 < len(rt) < len(rb)
    assert [(ident(a), ident(b), e) for a, b, e in rt] == [(ident(a), ident(b), e) for a, b, e in gt]
    r1, r2 = call(mods["AR"].vals["tightest_binary"][0], ab)
    g1, g2 = call(mods["AG"].vals["tightest_binary"][0], ab)
    assert (ident(r1), ident(r2)) == (ident(g1), ident(g2))
    assert call(mods["AG"].vals["bodies_to_el_phase_space"][0], ab) == \
        call(mods["AR"].vals["bodies_to_el_phase_space"][0], ab)


def test_dual_number_hermite_loops(world):
    """Dfloat_hermite_gpu: the two pair loops of dFloat_hermite.ml for K directions at once, beside
    the reference's own DFloat_hermite.start_bs / advance_body run one direction at a time over
    the restated Diff library (oracle/ref_fixtures/diff_restated)."""
    import random
    w = world
    w.set_eps2(0.0)
    w.fn("DFloat_advancer.DFloatBase.eps2").f["contents"] = Ctor("C", 0.0)
    n, K = 6, 3
    small = bodies("plummer_n16_std_seed20241.txt", n)
    rnd = random.Random(7)
    dq = [[rnd.uniform(-1, 1) for _ in range(3 * n)] for _ in range(K)]
    dp = [[rnd.uniform(-1, 1) for _ in range(3 * n)] for _ in range(K)]

    def D(v, d):
        return Ctor("D", (v, d))

    def val(x):
        return x.arg if x.name == "C" else x.arg[0]

    def tan(x):
        return 0.0 if x.name == "C" else x.arg[1]
    flat = lambda rows: [x for r in rows for x in r]  # noqa: E731
    m = [b[0] for b in small]
    q, p = flat(b[1] for b in small), flat(b[2] for b in small)
    zeros = lambda k: [0.0] * k  # noqa: E731
    state = Record(dict(n=n, k=K, t=zeros(n), m=m, q=q, p=p, f=zeros(3 * n), df=zeros(3 * n),
                        t_tan=zeros(K * n), q_tan=flat(dq), p_tan=flat(dp), f_tan=zeros(3 * n * K),
                        df_tan=zeros(3 * n * K)))
    w.run("Dfloat_hermite_gpu.start_forces", state)
    started = []
    for k in range(K):
        dbs = [Record(dict(id=i + 1, t=Ctor("C", 0.0), m=Ctor("C", m[i]),
                           q=[D(q[3 * i + c], dq[k][3 * i + c]) for c in range(3)],
                           p=[D(p[3 * i + c], dp[k][3 * i + c]) for c in range(3)], h=Ctor("C", -1.0),
                           f=[Ctor("C", 0.0)] * 3, df=[Ctor("C", 0.0)] * 3)) for i in range(n)]
        out = to_list(w.run("DFloat_hermite.start_bs", Ctor("C", 1e-2), from_list(dbs)))
        started.append(out)
        for i, b in enumerate(out):
            for c in range(3):
                fv, ft = state.f["f"][3 * i + c], state.f["f_tan"][3 * n * k + 3 * i + c]
                assert abs(fv - val(b.f["f"][c])) <= 1e-13 * abs(fv)
                assert abs(ft - tan(b.f["f"][c])) <= 1e-12 * max(abs(ft), 1e-3)
                dv, dt_ = state.f["df"][3 * i + c], state.f["df_tan"][3 * n * k + 3 * i + c]
                assert abs(dv - val(b.f["df"][c])) <= 1e-13 * abs(dv)
                assert abs(dt_ - tan(b.f["df"][c])) <= 1e-12 * max(abs(dt_), 1e-3)
    # advance_body of body 2 to a dual time: fp, dfp and their tangents
    nt, nt_tan = 1.0 / 128, [0.3, -0.2, 0.1]
    fp, dfp, fpt, dfpt = w.run("Dfloat_hermite_gpu.force_on_predicted", state, 2, nt, nt_tan)
    for k in range(K):
        bs = started[k]
        # {b with h = nt - t} so that next_time b = (nt, nt_tan): the reference's own advance_body
        b = Record(dict(bs[2].f, h=D(nt, nt_tan[k])))
        nb = w.run("DFloat_hermite.advance_body", Ctor("C", 1e-2), b,
                   from_list([x for j, x in enumerate(bs) if j != 2]))
        for c in range(3):
            assert abs(fp[c] - val(nb.f["f"][c])) <= 1e-13 * abs(fp[c])
            assert abs(dfp[c] - val(nb.f["df"][c])) <= 1e-12 * abs(dfp[c])
            assert abs(fpt[3 * k + c] - tan(nb.f["f"][c])) <= 1e-11 * max(abs(fpt[3 * k + c]), 1e-3)
            assert abs(dfpt[3 * k + c] - tan(nb.f["df"][c])) <= 1e-11 * max(abs(dfpt[3 * k + c]), 1e-3)
