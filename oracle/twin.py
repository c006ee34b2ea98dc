"""twin.py — a SECOND, independent restatement of the reference's pair physics and integrators,
in plain Python, written from the OCaml text (not from the C restatement next to it).

TEST INFRASTRUCTURE ONLY (like everything under oracle/): tests/test_oracle_twin.py runs it
beside the compiled oracle on small systems and demands bit-for-bit agreement.  Python floats
are IEEE binary64, CPython never contracts a*b+c into an FMA, math.sqrt is correctly rounded
and `**` on floats is the C library's pow - the same arithmetic an ocamlopt build of the
reference performs - so a transcription slip in either restatement (an association order, a
misplaced factor, a compiler flag that fuses or reorders) shows up as a differing bit.
It pins nothing against the OCaml binary itself (none can be built here: DESIGN.md §6).

Every function cites the reference lines it transcribes; bodies are dicts / lists of Python
floats, loops are the reference's loops in the reference's order.
"""
import math

INF = float("inf")
NAN = float("nan")
EPS = 2.220446049250313e-16  # epsilon_float


class Dual:
    """First-order forward-mode dual number, the stand-in for Diff.DFloat.diff (dFloat.ml:20; the
    library is not vendored and pinned nowhere): value + one tangent with the
    textbook sum / product / quotient rules, d sqrt x = dx / (2 sqrt x), d x^c = c x^(c-1) dx,
    comparisons on the value.  A float operand is the constant `C x` (zero tangent).  With it the
    Hermite functions below ARE dFloat_hermite.ml, as that file is hermite.ml over `diff`."""
    __slots__ = ("v", "d")

    def __init__(self, v, d=0.0):
        self.v, self.d = v, d

    @staticmethod
    def of(x):
        return x if isinstance(x, Dual) else Dual(x, 0.0)

    def __add__(a, b):
        b = Dual.of(b)
        return Dual(a.v + b.v, a.d + b.d)

    def __radd__(b, a):
        return Dual.of(a) + b

    def __sub__(a, b):
        b = Dual.of(b)
        return Dual(a.v - b.v, a.d - b.d)

    def __rsub__(b, a):
        return Dual.of(a) - b

    def __neg__(a):
        return Dual(-a.v, -a.d)

    def __mul__(a, b):
        b = Dual.of(b)
        return Dual(a.v * b.v, a.d * b.v + a.v * b.d)

    def __rmul__(b, a):
        return Dual.of(a) * b

    def __truediv__(a, b):
        b = Dual.of(b)
        q = a.v / b.v
        return Dual(q, (a.d - q * b.d) / b.v)

    def __rtruediv__(b, a):
        return Dual.of(a) / b

    def __pow__(a, c):
        return Dual(a.v ** c, c * a.v ** (c - 1.0) * a.d)

    def sqrt(a):
        r = math.sqrt(a.v)
        return Dual(r, a.d / (2.0 * r))

    def __lt__(a, b):
        return a.v < Dual.of(b).v

    def __gt__(a, b):
        return a.v > Dual.of(b).v

    def __le__(a, b):
        return a.v <= Dual.of(b).v

    def __ge__(a, b):
        return a.v >= Dual.of(b).v

    def __eq__(a, b):
        return a.v == Dual.of(b).v

    def __hash__(a):
        return hash(a.v)


def gsqrt(x):  # sqrt on a float or a Dual
    return x.sqrt() if isinstance(x, Dual) else math.sqrt(x)


# ---- base.ml -------------------------------------------------------------------------------------
def norm_squared(v):  # base.ml:36-40
    x, y, z = v[0], v[1], v[2]
    return x * x + y * y + z * z


def norm(v):  # base.ml:43
    return gsqrt(norm_squared(v))


def distance_squared(x, y):  # base.ml:46-50
    dx = x[0] - y[0]
    dy = x[1] - y[1]
    dz = x[2] - y[2]
    return dx * dx + dy * dy + dz * dz


def distance(x, y):  # base.ml:53
    return gsqrt(distance_squared(x, y))


def dot(x, y):  # base.ml:56-57
    return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]


def square(x):
    return x * x


def v(m1, q1, m2, q2, eps2):  # base.ml:60-63
    r2 = distance_squared(q1, q2)
    r = math.sqrt(r2 + eps2)
    return -m1 * m2 / r


def grad_v(m1, q1, m2, q2, gv, eps2):  # base.ml:73-80
    r2 = distance_squared(q1, q2)
    re = r2 + eps2
    r3 = re * math.sqrt(re)
    factor = m1 * m2 / r3
    gv[0] = factor * (q1[0] - q2[0])
    gv[1] = factor * (q1[1] - q2[1])
    gv[2] = factor * (q1[2] - q2[2])


def grad_v_dgrad_v(m1, q1, p1, m2, q2, p2, gv, dgv, eps2):  # base.ml:91-110
    gv[0] = q1[0] - q2[0]
    dgv[0] = p1[0] / m1 - p2[0] / m2
    gv[1] = q1[1] - q2[1]
    dgv[1] = p1[1] / m1 - p2[1] / m2
    gv[2] = q1[2] - q2[2]
    dgv[2] = p1[2] / m1 - p2[2] / m2
    r2 = norm_squared(gv)
    rdv = dot(gv, dgv)
    re = r2 + eps2
    r3 = gsqrt(re) * re
    r5 = r3 * re
    factorg = m1 * m2 / r3
    factord = 3.0 * rdv * m1 * m2 / r5
    dgv[0] = factorg * dgv[0] - factord * gv[0]
    gv[0] = factorg * gv[0]
    dgv[1] = factorg * dgv[1] - factord * gv[1]
    gv[1] = factorg * gv[1]
    dgv[2] = factorg * dgv[2] - factord * gv[2]
    gv[2] = factorg * gv[2]


# ---- energy.ml:57-78 -------------------------------------------------------------------------------
def total_kinetic_energy(m, p):
    e = 0.0
    for i in range(len(m)):
        e = e + norm_squared(p[i]) / (2.0 * m[i])
    return e


def total_potential_energy(m, q, eps2):
    pe = 0.0
    n = len(m)
    for i in range(n):
        for j in range(i + 1, n):
            pe = pe + v(m[i], q[i], m[j], q[j], eps2)
    return pe


# ---- hermite.ml -------------------------------------------------------------------------------------
def hermite_start_bs(eta, m, q, p, eps2):  # hermite.ml:148-169 -> f, df, h per body
    n = len(m)
    f = [[0.0] * 3 for _ in range(n)]
    df = [[0.0] * 3 for _ in range(n)]
    gv, dgv = [0.0] * 3, [0.0] * 3
    for a in range(n):
        for b in range(a + 1, n):
            grad_v_dgrad_v(m[a], q[a], p[a], m[b], q[b], p[b], gv, dgv, eps2)
            for i in range(3):
                f[a][i] = f[a][i] - gv[i]
                f[b][i] = f[b][i] + gv[i]
                df[a][i] = df[a][i] - dgv[i]
                df[b][i] = df[b][i] + dgv[i]
    h = [6.0 * eta * norm(f[a]) / (m[a] * norm(df[a])) for a in range(n)]  # :145-146
    return f, df, h


def _predict_position(b, t):  # hermite.ml:53-62
    dt = t - b["t"]
    pc = dt / b["m"]
    fc = pc * dt * 0.5
    dfc = fc * dt / 3.0
    return [b["q"][i] + pc * b["p"][i] + fc * b["f"][i] + dfc * b["df"][i] for i in range(3)]


def _predict_momentum(b, t):  # hermite.ml:64-72
    dt = t - b["t"]
    fc = dt
    dfc = fc * dt * 0.5
    return [b["p"][i] + fc * b["f"][i] + dfc * b["df"][i] for i in range(3)]


def _timescale(eta, m, q, qp, f, h):  # hermite.ml:87-93
    r = distance(q, qp)
    fn = norm(f)
    if r == 0.0:
        return h * (2.0 ** 0.25)
    return h * (eta * square(h) * fn / (m * r)) ** 0.25


def _advance_body(eta, b, others, eps2):  # hermite.ml:95-130
    nt = b["t"] + b["h"]
    qp = _predict_position(b, nt)
    pp = _predict_momentum(b, nt)
    fp, dfp, gv, dgv = [0.0] * 3, [0.0] * 3, [0.0] * 3, [0.0] * 3
    for b2 in others:
        qp2 = _predict_position(b2, nt)
        pp2 = _predict_momentum(b2, nt)
        grad_v_dgrad_v(b["m"], qp, pp, b2["m"], qp2, pp2, gv, dgv, eps2)
        for i in range(3):
            fp[i] = fp[i] - gv[i]
            dfp[i] = dfp[i] - dgv[i]
    h, m, q, p, f, df = b["h"], b["m"], b["q"], b["p"], b["f"], b["df"]
    p_new, q_new = [0.0] * 3, [0.0] * 3
    for i in range(3):
        p_new[i] = p[i] + 0.5 * h * (f[i] + fp[i]) + square(h) / 12.0 * (df[i] - dfp[i])
        q_new[i] = q[i] + 0.5 * h / m * (p[i] + p_new[i]) + square(h) / (12.0 * m) * (f[i] - fp[i])
    nb = dict(b)
    nb.update(t=b["t"] + h, q=q_new, p=p_new, f=fp, df=dfp, h=_timescale(eta, m, q_new, qp, fp, h))
    return nb


def _merge_one(nb, bs):  # List.merge compare_next_time [new_b] bs (stable: new_b first on ties)
    knew = nb["t"] + nb["h"]
    out, placed = [], False
    for b in bs:
        kb = b["t"] + b["h"]
        cmp = -1 if knew < kb else (1 if knew > kb else 0)
        if not placed and cmp <= 0:
            out.append(nb)
            placed = True
        out.append(b)
    if not placed:
        out.append(nb)
    return out


def hermite_advance(m, q, p, dt, sf, eps2, t0=0.0):  # hermite.ml:180-189 on fresh bodies
    """Returns (bodies sorted by id, number of advance_body calls)."""
    n = len(m)
    f, df, h = hermite_start_bs(sf, m, q, p, eps2)
    bs = [dict(id=i + 1, t=t0, m=m[i], q=list(q[i]), p=list(p[i]), f=f[i], df=df[i], h=h[i])
          for i in range(n)]
    # List.fast_sort compare_next_time (stable)
    bs = sorted(bs, key=lambda b: b["t"] + b["h"])
    stop = t0 + dt
    nsteps = 0
    while bs[0]["t"] + bs[0]["h"] <= stop:  # advance_bodies (hermite.ml:171-178)
        b, rest = bs[0], bs[1:]
        nb = _advance_body(sf, b, rest, eps2)
        nsteps += 1
        bs = _merge_one(nb, rest)
    done = []  # finish_bs (hermite.ml:132-143)
    while bs:
        b, bs = bs[0], bs[1:]
        b2 = dict(b)
        b2["h"] = stop - b["t"]
        done = [_advance_body(sf, b2, bs + done, eps2)] + done
        nsteps += 1
    done.sort(key=lambda b: b["id"])
    return done, nsteps


# ---- leapfrog.ml ------------------------------------------------------------------------------------
def _lf_kick(eta, dt, b1, b2):  # leapfrog.ml:70-104
    q1, v1, q2, v2 = b1["q"], b1["v"], b2["q"], b2["v"]
    m1, m2 = b1["m"], b2["m"]
    dx = q2[0] - q1[0]
    dy = q2[1] - q1[1]
    dz = q2[2] - q1[2]
    dvx = v2[0] - v1[0]
    dvy = v2[1] - v1[1]
    dvz = v2[2] - v1[2]
    r2 = dx * dx + dy * dy + dz * dz
    vsq = dvx * dvx + dvy * dvy + dvz * dvz
    vdr = dx * dvx + dy * dvy + dz * dvz
    r = math.sqrt(r2)
    r3 = r2 * r
    c = _div(dt, r3)
    c1 = m2 * c
    c2 = m1 * c
    v1[0] = v1[0] + c1 * dx
    v1[1] = v1[1] + c1 * dy
    v1[2] = v1[2] + c1 * dz
    v2[0] = v2[0] - c2 * dx
    v2[1] = v2[1] - c2 * dy
    v2[2] = v2[2] - c2 * dz
    mtot = m1 + m2
    tff = eta * _sqrt(_div(r3, mtot))
    tc = eta * _sqrt(_div(r2, vsq))
    scaled_vdr = _div(vdr, r2)
    dtff = 1.5 * scaled_vdr * tff
    dtc = scaled_vdr * tc * (1.0 + _div(mtot, vsq * r))
    tff = abs(_div(tff, 1.0 - 0.5 * dtff))
    tc = abs(_div(tc, 1.0 - 0.5 * dtc))
    tscale = tff if tff < tc else tc
    if b1["dt_max"] > tscale:
        b1["dt_max"] = tscale
    if b2["dt_max"] > tscale:
        b2["dt_max"] = tscale


def _div(a, b):  # OCaml's /. follows IEEE: x/0 = inf or nan, no exception
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0.0:
            return NAN
        return math.copysign(INF, a) * math.copysign(1.0, b)


def _sqrt(x):  # OCaml's sqrt: nan for negative arguments, inf for inf
    if x != x or x < 0.0:
        return NAN
    return math.sqrt(x) if x != INF else INF


def _ocaml_compare(x, y):  # polymorphic compare on floats: nan below everything, equal to itself
    if x < y:
        return -1
    if x > y:
        return 1
    if x == y:
        return 0
    xn, yn = x != x, y != y
    if xn and yn:
        return 0
    return -1 if xn else 1


def _stable_sort(bs, cmp):
    import functools
    return sorted(bs, key=functools.cmp_to_key(cmp))


def _lf_drift(dt, b):  # leapfrog.ml:64-68
    b["t"] = b["t"] + dt
    for k in range(3):
        b["q"][k] = b["q"][k] + dt * b["v"][k]


def _lf_advance_subsystem(dt, eta, bs, iend, count):  # leapfrog.ml:119-142
    if iend <= 0:
        return
    ibegin = iend  # find_can_advance_index (:113-117)
    while not (ibegin == 0 or bs[ibegin - 1]["dt_max"] < dt):
        ibegin -= 1
    dt2 = dt / 2.0
    for i in range(ibegin, iend):
        bs[i]["dt_max"] = INF
    _lf_advance_subsystem(dt2, eta, bs, ibegin, count)
    for i in range(ibegin, iend):
        _lf_drift(dt2, bs[i])
    for i in range(ibegin, iend):
        for j in range(i):
            _lf_kick(eta, dt, bs[i], bs[j])
    count[0] += 1  # a diagnostic, counted like the compiled restatement: calls with iend > 0
    for i in range(ibegin, iend):
        _lf_drift(dt2, bs[i])
    _lf_advance_subsystem(dt2, eta, bs, ibegin, count)
    bs[:iend] = _stable_sort(bs[:iend], lambda a, b: _ocaml_compare(a["dt_max"], b["dt_max"]))  # sort_sub


def leapfrog_advance(m, q, p, dt, eta):  # Leapfrog.make (:28-34) then advance (:144-158)
    n = len(m)
    bs = [dict(id=i + 1, t=0.0, dt_max=0.0, m=m[i], q=list(q[i]), v=[x / m[i] for x in p[i]])
          for i in range(n)]
    for b in bs:
        b["dt_max"] = INF
    for i in range(n):
        for j in range(i + 1, n):
            _lf_kick(eta, 0.0, bs[i], bs[j])
    bs = _stable_sort(bs, lambda a, b: _ocaml_compare(a["dt_max"], b["dt_max"]))
    count = [0]
    _lf_advance_subsystem(dt, eta, bs, n, count)
    return bs, count[0]


# ---- omelyan_advancer.ml:44-104 ----------------------------------------------------------------------
def omelyan_advance(m, q, p, t, h, eps2):
    n = len(m)
    q = [list(x) for x in q]
    p = [list(x) for x in p]
    gv = [0.0] * 3

    def b_operate():
        f = h / 6.0
        for a in range(n):
            for b in range(a + 1, n):
                grad_v(m[a], q[a], m[b], q[b], gv, eps2)
                for i in range(3):
                    p[a][i] = p[a][i] - f * gv[i]
                    p[b][i] = p[b][i] + f * gv[i]

    def a_operate():
        for a in range(n):
            f = h / (2.0 * m[a])
            for i in range(3):
                q[a][i] = q[a][i] + p[a][i] * f

    def c_operate():
        pf = 2.0 * h / 3.0
        qs = [list(x) for x in q]
        for a in range(n):
            f = square(h) / (24.0 * m[a])
            for b in range(a + 1, n):
                f2 = square(h) / (24.0 * m[b])
                grad_v(m[a], q[a], m[b], q[b], gv, eps2)
                for i in range(3):
                    qs[a][i] = qs[a][i] - f * gv[i]
                    qs[b][i] = qs[b][i] + f2 * gv[i]
        for a in range(n):
            for b in range(a + 1, n):
                grad_v(m[a], qs[a], m[b], qs[b], gv, eps2)
                for i in range(3):
                    p[a][i] = p[a][i] - pf * gv[i]
                    p[b][i] = p[b][i] + pf * gv[i]

    b_operate()
    a_operate()
    c_operate()
    a_operate()
    b_operate()
    return q, p, t + h


# ---- advancer.ml:143-428 ----------------------------------------------------------------------------
def _adv_make(i, t, m, q, p):  # advancer.ml:98-116
    nan3 = lambda: [NAN, NAN, NAN]  # noqa: E731
    return dict(id=i, t=t, m=m, q=list(q), p=list(p), t0=NAN, h=NAN, hmax=NAN, p0=nan3(), q0=nan3(),
                q1=nan3(), q2=nan3(), dVdq0=nan3(), dVdq1=nan3(), dVdq2=nan3(), cs=nan3())


def _predict_q012(b, hnew):  # advancer.ml:143-164
    m, q, p, h = b["m"], b["q"], b["p"], b["h"]
    dV0, dV1, dV2 = b["dVdq0"], b["dVdq1"], b["dVdq2"]
    hnew2 = hnew * hnew
    h2 = h * h
    hnew3 = hnew2 * hnew
    h3 = h2 * h
    for i in range(3):
        b["p0"][i] = p[i]
        b["q0"][i] = q[i]
        q0 = b["q0"]
        b["q1"][i] = (q0[i] + hnew * p[i] / (2.0 * m)
                      - hnew3 * (h + hnew) / (8.0 * h3 * m) * dV0[i]
                      + hnew3 * (2.0 * h + hnew) / (16.0 * h3 * m) * dV1[i]
                      - hnew2 * (6.0 * h2 + 3.0 * h * hnew + hnew2) / (8.0 * h3 * m) * dV2[i])
        b["q2"][i] = (q0[i] + hnew * p[i] / m
                      - hnew3 * (h + hnew) / (h3 * m) * dV0[i]
                      + hnew3 * (2.0 * h + hnew) / (2.0 * h3 * m) * dV1[i]
                      - hnew2 * (3.0 * h2 + 3.0 * h * hnew + hnew2) / (h3 * m) * dV2[i])
    b["h"] = hnew
    b["t0"] = b["t"]


def _clear_before_step(b):  # advancer.ml:166-171
    for k in ("dVdq0", "dVdq1", "dVdq2", "cs"):
        b[k][0] = b[k][1] = b[k][2] = 0.0
    b["cs"][0] = 1.0


def _predict_position_a(b, t):  # advancer.ml:173-195
    if b["t"] == t:
        return
    dt = t - b["t0"]
    h = b["h"]
    h2 = h * h
    cs = b["cs"]
    cs[0] = (2.0 * dt * dt - 3.0 * dt * h + h2) / h2
    cs[1] = (4.0 * dt * (h - dt)) / h2
    cs[2] = dt * (2.0 * dt - h) / h2
    c0, c1, c2 = cs[0], cs[1], cs[2]
    for i in range(3):
        b["q"][i] = c0 * b["q0"][i] + c1 * b["q1"][i] + c2 * b["q2"][i]
    b["t"] = t


def _interact_bodies(bs, imin, imax, t, vC, eps2, stats):  # advancer.ml:207-236
    n = len(bs)
    gv = [0.0] * 3
    for i in range(imin, n):
        _predict_position_a(bs[i], t)
    for i in range(imin, imax):
        b = bs[i]
        c0, c1, c2 = vC * b["cs"][0], vC * b["cs"][1], vC * b["cs"][2]
        for j in range(i + 1, n):
            b2 = bs[j]
            c02, c12, c22 = vC * b2["cs"][0], vC * b2["cs"][1], vC * b2["cs"][2]
            grad_v(b["m"], b["q"], b2["m"], b2["q"], gv, eps2)
            for k in range(3):
                g = gv[k]
                b["dVdq0"][k] = b["dVdq0"][k] + c0 * g
                b["dVdq1"][k] = b["dVdq1"][k] + c1 * g
                b["dVdq2"][k] = b["dVdq2"][k] + c2 * g
                b2["dVdq0"][k] = b2["dVdq0"][k] - c02 * g
                b2["dVdq1"][k] = b2["dVdq1"][k] - c12 * g
                b2["dVdq2"][k] = b2["dVdq2"][k] - c22 * g
    stats[0] += 1
    na, nn = imax - imin, n - imin
    stats[1] += na * (nn - 1) - na * (na - 1) // 2


def _finish_body(b, t):  # advancer.ml:238-252
    cp = b["h"] / b["m"]
    cV0 = cp
    cV1 = 0.5 * cp
    for i in range(3):
        v0i, v1i, v2i, pi = b["dVdq0"][i], b["dVdq1"][i], b["dVdq2"][i], b["p0"][i]
        b["q"][i] = b["q0"][i] + cp * pi - cV0 * v0i - cV1 * v1i
        b["p"][i] = pi - v0i - v1i - v2i
    b["t"] = t


def _assign_timestep(b, dt):  # advancer.ml:257-266 (the warning is output only)
    if dt < 100.0 * EPS * abs(b["t"]) + 100.0 * EPS:
        b["hmax"] = 100.0 * EPS * b["t"] + 100.0 * EPS
    else:
        b["hmax"] = dt


def _h_adaptive(sf, b):  # advancer.ml:272-281
    h = b["h"]
    dq = distance(b["q"], b["q2"])
    if dq == 0.0:
        _assign_timestep(b, h * 2.01)
    else:
        a = norm(b["dVdq2"]) * 6.0 / h / b["m"]
        a_dq = 0.5 * h * h * a
        ratio = a_dq / dq
        _assign_timestep(b, h * (sf * ratio) ** 0.3333333)


def _hmax_lt(b1, b2):  # advancer.ml:283-289
    h1, h2 = b1["hmax"], b2["hmax"]
    return -1 if h1 < h2 else (1 if h1 > h2 else 0)


def _advance_range(bs, imax, dt, sf, eps2, stats):  # advancer.ml:316-350, extpot = None
    if bs[imax - 1]["hmax"] < dt:
        _advance_range(bs, imax, dt / 2.0, sf, eps2, stats)
        _advance_range(bs, imax, dt / 2.0, sf, eps2, stats)
        return
    i = imax - 1
    while i >= 0 and dt < bs[i]["hmax"]:
        i -= 1
    imin = i + 1 if i >= 0 else 0
    for i in range(imin, imax):
        _predict_q012(bs[i], dt)
        _clear_before_step(bs[i])
    t0 = bs[imin]["t"]
    _interact_bodies(bs, imin, imax, t0, dt / 6.0, eps2, stats)
    if imin > 0:
        _advance_range(bs, imin, dt / 2.0, sf, eps2, stats)
    _interact_bodies(bs, imin, imax, t0 + 0.5 * dt, 2.0 * dt / 3.0, eps2, stats)
    if imin > 0:
        _advance_range(bs, imin, dt / 2.0, sf, eps2, stats)
    _interact_bodies(bs, imin, imax, t0 + dt, dt / 6.0, eps2, stats)
    for i in range(imin, imax):
        _finish_body(bs[i], t0 + dt)
        _h_adaptive(sf, bs[i])
    bs[:imax] = _stable_sort(bs[:imax], _hmax_lt)  # sort_range


def _initialize_before_first_step(bs, sf, eps2):  # advancer.ml:357-406
    n = len(bs)
    gv, dgv, fdot = [0.0] * 3, [0.0] * 3, [0.0] * 3
    h = 1.0
    c0 = h / 6.0
    c1 = 2.0 * h / 3.0
    c2 = c0
    for b in bs:
        _clear_before_step(b)
        b["h"] = h
    for i in range(n):
        bi = bs[i]
        for j in range(i + 1, n):
            bj = bs[j]
            grad_v_dgrad_v(bi["m"], bi["q"], bi["p"], bj["m"], bj["q"], bj["p"], gv, dgv, eps2)
            for k in range(3):
                bi["dVdq0"][k] = bi["dVdq0"][k] + c0 * (gv[k] - h * dgv[k])
                bi["dVdq1"][k] = bi["dVdq1"][k] + c1 * (gv[k] - 0.5 * h * dgv[k])
                bi["dVdq2"][k] = bi["dVdq2"][k] + c2 * gv[k]
                bj["dVdq0"][k] = bj["dVdq0"][k] - c0 * (gv[k] - h * dgv[k])
                bj["dVdq1"][k] = bj["dVdq1"][k] - c1 * (gv[k] - 0.5 * h * dgv[k])
                bj["dVdq2"][k] = bj["dVdq2"][k] - c2 * gv[k]
    for b in bs:
        for j in range(3):
            fdot[j] = b["dVdq0"][j] - b["dVdq1"][j] + 3.0 * b["dVdq2"][j]
        _assign_timestep(b, 1e-2 * sf ** 0.33333 * b["h"] * norm(b["dVdq2"]) / norm(fdot))


def advancer_advance(m, q, p, dt, sf, eps2, t=0.0):  # A.make then A.advance (advancer.ml:422-428)
    """Returns (bodies in the reference's output order, [interact_bodies calls, pair interactions])."""
    bs = [_adv_make(i, t, m[i], q[i], p[i]) for i in range(len(m))]
    stats = [0, 0]
    # first_step: classify_float h <> FP_normal (advancer.ml:417-420) - fresh bodies carry nan
    if any(not (b["h"] == b["h"] and abs(b["h"]) != INF and abs(b["h"]) >= 2.2250738585072014e-308)
           for b in bs):
        _initialize_before_first_step(bs, sf, eps2)
    bs = _stable_sort(bs, _hmax_lt)
    _advance_range(bs, len(bs), dt, sf, eps2, stats)
    return bs, stats


# ---- analysis.ml ------------------------------------------------------------------------------------
def binary_energy(m1, q1, p1, m2, q2, p2, eps2):  # analysis.ml:809-822
    vrel = [0.0] * 3
    for i in range(3):
        vrel[i] = p2[i] / m2 - p1[i] / m1
    mu = m1 * m2 / (m1 + m2)
    ke = 0.5 * mu * dot(vrel, vrel)
    pe = v(m1, q1, m2, q2, eps2)
    return ke + pe


def binaries_threshold(m, q, p, eps2, eg):  # analysis.ml:849-861, pairs in loop order
    out = []
    n = len(m)
    for i in range(n):
        for j in range(i + 1, n):
            e = binary_energy(m[i], q[i], p[i], m[j], q[j], p[j], eps2)
            if e < -abs(eg):
                out.append((i, j, e))
    return out


def bodies_to_el_phase_space(m, q, p, eps2):  # analysis.ml:904-919 (+ :539-543, energy.ml:57-58)
    n = len(m)
    out = []
    for i in range(n):
        qi, pi = q[i], p[i]
        lvec = [qi[1] * pi[2] - qi[2] * pi[1], qi[2] * pi[0] - qi[0] * pi[2], qi[0] * pi[1] - qi[1] * pi[0]]
        lb = norm(lvec)
        ke = norm_squared(pi) / (2.0 * m[i])
        pe = 0.0
        for j in range(i):
            pe = pe + v(m[i], qi, m[j], q[j], eps2)
        for j in range(i + 1, n):
            pe = pe + v(m[i], qi, m[j], q[j], eps2)
        etot = ke + pe
        out.append([etot, etot / m[i], lb, lb / m[i]])
    return out


# ---- ics.ml -------------------------------------------------------------------------------------------
def adjust_frame(m, q, p):  # ics.ml:161-175 over analysis.ml:508-537
    n = len(m)
    p_tot = [0.0] * 3
    for b in range(n):
        for i in range(3):
            p_tot[i] = p_tot[i] + p[b][i]
    m_tot = 0.0
    for b in range(n):
        m_tot = m_tot + m[b]
    com = [0.0] * 3
    for b in range(n):
        for i in range(3):
            com[i] = com[i] + q[b][i] * (m[b] / m_tot)
    qn = [[q[b][i] - com[i] for i in range(3)] for b in range(n)]
    pn = [[p[b][i] - p_tot[i] * m[b] / m_tot for i in range(3)] for b in range(n)]
    return list(m), qn, pn


def rescale_to_standard_units(m, q, p, eps2):  # ics.ml:304-331
    n = len(m)
    mtot = 0.0
    for b in range(n):
        mtot = mtot + m[b]
    mn = [m[b] / mtot for b in range(n)]
    pn = [[p[b][i] * mn[b] / m[b] for i in range(3)] for b in range(n)]
    e = total_kinetic_energy(mn, pn) + total_potential_energy(mn, q, eps2)
    assert e < 0.0
    efactor = (-0.25) / e
    q2 = [[q[b][i] / efactor for i in range(3)] for b in range(n)]
    p2 = [[pn[b][i] * math.sqrt(efactor) for i in range(3)] for b in range(n)]
    return mn, q2, p2


def leapfrog_advance_const_dt(m, q, p, dt, nsteps):  # Leapfrog.make then advance_const_dt (:160-165)
    bs = [dict(id=i + 1, t=0.0, dt_max=0.0, m=m[i], q=list(q[i]), v=[x / m[i] for x in p[i]])
          for i in range(len(m))]
    count = [0]
    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 6000))  # dt_max = 0 halves dt down to the last subnormal: ~1100 levels
    try:
        for _ in range(nsteps):
            _lf_advance_subsystem(dt, INF, bs, len(bs), count)
    finally:
        sys.setrecursionlimit(old)
    return bs


# ---- ics.ml recipes over the fixed generator of SURVEY.md §8d ------------------------------------------
class Rng:
    """splitmix64; Random.float x = (next >> 11) * 2^-53 * x; one draw per Random.float call, draws
    in the textual order of each recipe (the reference's own stream is self-seeded and cannot be
    reproduced)."""
    M = (1 << 64) - 1

    def __init__(self, seed):
        self.s = seed & self.M

    def float(self, x):
        self.s = (self.s + 0x9E3779B97F4A7C15) & self.M
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.M
        z = z ^ (z >> 31)
        return float(z >> 11) * 2.0 ** -53 * x


def _random_between(rng, a, b):  # ics.ml:177-179
    dx = b - a
    return a + rng.float(dx)


def _random_from_dist(rng, dist, xmin=0.0, xmax=1.0, ymin=0.0, ymax=1.0):  # ics.ml:181-187
    x = _random_between(rng, xmin, xmax)
    y = _random_between(rng, ymin, ymax)
    while not (y < dist(x)):
        x = _random_between(rng, xmin, xmax)
        y = _random_between(rng, ymin, ymax)
    return x


def _random_vector(rng, r):  # ics.ml:189-197
    pi = 4.0 * math.atan(1.0)
    cos_theta = _random_between(rng, -1.0, 1.0)
    phi = _random_between(rng, 0.0, 2.0 * pi)
    theta = math.acos(cos_theta)
    sin_theta = math.sin(theta)
    return [r * math.cos(phi) * sin_theta, r * math.sin(phi) * sin_theta, r * cos_theta]


def make_plummer(n, seed):  # ics.ml:211-227
    rng = Rng(seed)
    m = 1.0 / float(n)
    pi = 4.0 * math.atan(1.0)
    sf = 16.0 / (3.0 * pi)
    ms, qs, ps = [], [], []
    for _ in range(n):
        r = 1.0 / math.sqrt(_random_between(rng, 0.0, 1.0) ** (-2.0 / 3.0) - 1.0)
        x = _random_from_dist(rng, lambda x: square(x) * (1.0 - square(x)) ** 3.5, ymax=0.1)
        q = _random_vector(rng, r / sf)
        p = _random_vector(rng, math.sqrt(sf) * m * x * math.sqrt(2.0) * (1.0 + square(r)) ** (-0.25))
        ms.append(m)
        qs.append(q)
        ps.append(p)
    return adjust_frame(ms, qs, ps)


def make_hot_spherical(n, seed):  # ics.ml:229-239
    rng = Rng(seed)
    v0 = math.sqrt(21.0 / 10.0)
    m = 1.0 / float(n)
    ms, qs, ps = [], [], []
    for _ in range(n):
        r = _random_from_dist(rng, square)
        vv = _random_between(rng, 0.0, v0)
        qs.append(_random_vector(rng, r))
        ps.append(_random_vector(rng, m * vv))
        ms.append(m)
    return adjust_frame(ms, qs, ps)


# ---- dFloat_pushforward.ml:39-76 over dFloat_hermite.ml ----------------------------------------------
def hermite_jacobian_advance(m, q, p, dt, sf, eps2):
    """jacobian_advance dt sf bs on fresh bodies: one forward pass of Hermite.advance over Dual per
    input direction; state packed as [q(3n); p(3n)].  Returns (jac[6n][6n] with
    jac[r][c] = d out[r] / d in[c], the advanced packed state)."""
    n = len(m)
    d6 = 6 * n
    jac = [[0.0] * d6 for _ in range(d6)]
    out = [0.0] * d6
    for c in range(d6):
        qd = [[Dual(q[i][k], 1.0 if c == 3 * i + k else 0.0) for k in range(3)] for i in range(n)]
        pd = [[Dual(p[i][k], 1.0 if c == 3 * n + 3 * i + k else 0.0) for k in range(3)] for i in range(n)]
        md = [Dual(x) for x in m]
        bs, _ = hermite_advance(md, qd, pd, Dual(dt), Dual(sf), Dual(eps2), t0=Dual(0.0))
        for i in range(n):
            for k in range(3):
                jac[3 * i + k][c] = bs[i]["q"][k].d
                jac[3 * n + 3 * i + k][c] = bs[i]["p"][k].d
                if c == 0:
                    out[3 * i + k] = bs[i]["q"][k].v
                    out[3 * n + 3 * i + k] = bs[i]["p"][k].v
    return jac, out
