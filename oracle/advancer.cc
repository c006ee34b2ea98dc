// oracle/advancer.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/base.h header).
//
// Restatement of Advancer.A, the block-time-step Simpson-rule integrator bin/nbody.ml and
// bin/el_mass_segregation.ml actually run:
//   advancer.ml:143-164  predict_q012          advancer.ml:166-171  clear_before_step
//   advancer.ml:173-195  predict_position      advancer.ml:197-205  predict_all
//   advancer.ml:207-236  interact_bodies       advancer.ml:238-252  finish_body
//   advancer.ml:257-281  assign_timestep / h_adaptive
//   advancer.ml:283-297  hmax_lt / sort_range  advancer.ml:299-314  do_external_potential
//   advancer.ml:316-350  advance_range         advancer.ml:357-406  initialize_before_first_step
//   advancer.ml:408-428  first_step / advance
// Each expression keeps the reference's association order.  The follow_flag/print branch
// (:343-347) is output only and omitted.
#include "oracle.h"
extern "C" {
#include "base.h"
}

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

extern long g_oracle_step_cap;  // hermite.cc: test-run guard, not part of the algorithm

namespace {

struct ABody {
  int id;
  double t, m, q[3], p[3], h, hmax, t0;
  double dVdq0[3], dVdq1[3], dVdq2[3], p0[3], q0[3], q1[3], q2[3], cs[3];
};

struct Ctx {
  double eps2;
  oracle_extpot_fn extpot;  // gradV of an external potential, or NULL
  void *extpot_arg;
  long n_interact;          // number of interact_bodies calls
  long n_pairs;             // number of Base.grad_v evaluations in them
};

void predict_q012(ABody &b, double hnew) {
  double h = b.h, m = b.m;
  double hnew2 = hnew * hnew, h2 = h * h;
  double hnew3 = hnew2 * hnew, h3 = h2 * h;
  for (int i = 0; i < 3; ++i) {
    b.p0[i] = b.p[i];
    b.q0[i] = b.q[i];
    b.q1[i] = b.q0[i] + hnew * b.p[i] / (2.0 * m) -
              hnew3 * (h + hnew) / (8.0 * h3 * m) * b.dVdq0[i] +
              hnew3 * (2.0 * h + hnew) / (16.0 * h3 * m) * b.dVdq1[i] -
              hnew2 * (6.0 * h2 + 3.0 * h * hnew + hnew2) / (8.0 * h3 * m) * b.dVdq2[i];
    b.q2[i] = b.q0[i] + hnew * b.p[i] / m - hnew3 * (h + hnew) / (h3 * m) * b.dVdq0[i] +
              hnew3 * (2.0 * h + hnew) / (2.0 * h3 * m) * b.dVdq1[i] -
              hnew2 * (3.0 * h2 + 3.0 * h * hnew + hnew2) / (h3 * m) * b.dVdq2[i];
  }
  b.h = hnew;
  b.t0 = b.t;
}

void clear_before_step(ABody &b) {
  for (int i = 0; i < 3; ++i) b.dVdq0[i] = b.dVdq1[i] = b.dVdq2[i] = b.cs[i] = 0.0;
  b.cs[0] = 1.0;
}

void predict_position(ABody &b, double t) {
  if (b.t == t) return;
  double dt = t - b.t0, h = b.h;
  double h2 = h * h;
  b.cs[0] = (2.0 * dt * dt - 3.0 * dt * h + h2) / h2;
  b.cs[1] = (4.0 * dt * (h - dt)) / h2;
  b.cs[2] = dt * (2.0 * dt - h) / h2;
  double c0 = b.cs[0], c1 = b.cs[1], c2 = b.cs[2];
  for (int i = 0; i < 3; ++i) b.q[i] = c0 * b.q0[i] + c1 * b.q1[i] + c2 * b.q2[i];
  b.t = t;
}

void predict_all(ABody &b, double t) {
  predict_position(b, t);
  double dt = t - b.t0, h = b.h;
  double c0 = (4.0 * dt - 3.0 * h) / (h * h), c1 = 4.0 * (h - 2.0 * dt) / (h * h),
         c2 = (4.0 * dt - h) / (h * h);
  for (int i = 0; i < 3; ++i) b.p[i] = b.m * (c0 * b.q0[i] + c1 * b.q1[i] + c2 * b.q2[i]);
}

void interact_bodies(Ctx &cx, std::vector<ABody> &bs, int imin, int imax, double t, double vC) {
  int n = (int)bs.size();
  double gv[3];
  for (int i = imin; i < n; ++i) predict_position(bs[i], t);
  for (int i = imin; i < imax; ++i) {
    ABody &b = bs[i];
    double c0 = vC * b.cs[0], c1 = vC * b.cs[1], c2 = vC * b.cs[2];
    for (int j = i + 1; j < n; ++j) {
      ABody &b2 = bs[j];
      double c02 = vC * b2.cs[0], c12 = vC * b2.cs[1], c22 = vC * b2.cs[2];
      ob_grad_v(cx.eps2, b.m, b.q, b2.m, b2.q, gv);
      for (int k = 0; k < 3; ++k) {
        double gvi = gv[k];
        b.dVdq0[k] = b.dVdq0[k] + c0 * gvi;
        b.dVdq1[k] = b.dVdq1[k] + c1 * gvi;
        b.dVdq2[k] = b.dVdq2[k] + c2 * gvi;
        b2.dVdq0[k] = b2.dVdq0[k] - c02 * gvi;
        b2.dVdq1[k] = b2.dVdq1[k] - c12 * gvi;
        b2.dVdq2[k] = b2.dVdq2[k] - c22 * gvi;
      }
    }
    cx.n_pairs += n - 1 - i;
  }
  cx.n_interact += 1;
}

void finish_body(ABody &b, double t) {
  double cp = b.h / b.m;
  double cV0 = cp, cV1 = 0.5 * cp;
  for (int i = 0; i < 3; ++i) {
    double v0i = b.dVdq0[i], v1i = b.dVdq1[i], v2i = b.dVdq2[i], pi = b.p0[i];
    b.q[i] = b.q0[i] + cp * pi - cV0 * v0i - cV1 * v1i;
    b.p[i] = pi - v0i - v1i - v2i;
  }
  b.t = t;
}

void assign_timestep(ABody &b, double dt) {
  if (dt < 100.0 * DBL_EPSILON * fabs(b.t) + 100.0 * DBL_EPSILON)
    b.hmax = 100.0 * DBL_EPSILON * b.t + 100.0 * DBL_EPSILON;
  else
    b.hmax = dt;
}

void h_adaptive(double sf, ABody &b) {
  double h = b.h;
  double dq = ob_distance(b.q, b.q2);
  if (dq == 0.0) {
    assign_timestep(b, h * 2.01);
  } else {
    double a = ob_norm(b.dVdq2) * 6.0 / h / b.m;
    double a_dq = 0.5 * h * h * a;
    double ratio = a_dq / dq;
    assign_timestep(b, h * pow(sf * ratio, 0.3333333));
  }
}

bool hmax_less(const ABody &a, const ABody &b) { return a.hmax < b.hmax; }

void do_external_potential(Ctx &cx, ABody &b) {
  if (!cx.extpot) return;
  double h = b.h, gv0[3], gv1[3], gv2[3];
  predict_all(b, b.t0);
  cx.extpot(cx.extpot_arg, b.m, b.q, b.p, gv0);
  predict_all(b, b.t0 + 0.5 * b.h);
  cx.extpot(cx.extpot_arg, b.m, b.q, b.p, gv1);
  predict_all(b, b.t0 + b.h);
  cx.extpot(cx.extpot_arg, b.m, b.q, b.p, gv2);
  for (int i = 0; i < 3; ++i) {
    b.dVdq0[i] = b.dVdq0[i] + h * gv0[i] / 6.0;
    b.dVdq1[i] = b.dVdq1[i] + 2.0 * h * gv1[i] / 3.0;
    b.dVdq2[i] = b.dVdq2[i] + h * gv2[i] / 6.0;
  }
}

void advance_range(Ctx &cx, std::vector<ABody> &bs, int imax, double dt, double sf) {
  if (cx.n_interact > g_oracle_step_cap) return;
  if (bs[imax - 1].hmax < dt) {
    advance_range(cx, bs, imax, dt / 2.0, sf);
    advance_range(cx, bs, imax, dt / 2.0, sf);
    return;
  }
  int imin = imax - 1;
  while (imin >= 0 && dt < bs[imin].hmax) --imin;
  imin = imin + 1;  // loop returns 0 when i < 0, else i+1
  for (int i = imin; i < imax; ++i) {
    predict_q012(bs[i], dt);
    clear_before_step(bs[i]);
  }
  double t0 = bs[imin].t;
  interact_bodies(cx, bs, imin, imax, t0, dt / 6.0);
  if (imin > 0) advance_range(cx, bs, imin, dt / 2.0, sf);
  interact_bodies(cx, bs, imin, imax, t0 + 0.5 * dt, 2.0 * dt / 3.0);
  if (imin > 0) advance_range(cx, bs, imin, dt / 2.0, sf);
  interact_bodies(cx, bs, imin, imax, t0 + dt, dt / 6.0);
  for (int i = imin; i < imax; ++i) {
    do_external_potential(cx, bs[i]);
    finish_body(bs[i], t0 + dt);
    h_adaptive(sf, bs[i]);
  }
  std::stable_sort(bs.begin(), bs.begin() + imax, hmax_less);
}

void initialize_before_first_step(Ctx &cx, std::vector<ABody> &bs, double sf) {
  int n = (int)bs.size();
  double gv[3], dgv[3], fdot[3], h = 1.0;
  double c0 = h / 6.0, c1 = 2.0 * h / 3.0;
  double c2 = c0;
  for (int i = 0; i < n; ++i) {
    clear_before_step(bs[i]);
    bs[i].h = h;
  }
  for (int i = 0; i < n; ++i) {
    ABody &b = bs[i];
    for (int j = i + 1; j < n; ++j) {
      ABody &bj = bs[j];
      ob_grad_v_dgrad_v(cx.eps2, b.m, b.q, b.p, bj.m, bj.q, bj.p, gv, dgv);
      for (int k = 0; k < 3; ++k) {
        b.dVdq0[k] = b.dVdq0[k] + c0 * (gv[k] - h * dgv[k]);
        b.dVdq1[k] = b.dVdq1[k] + c1 * (gv[k] - 0.5 * h * dgv[k]);
        b.dVdq2[k] = b.dVdq2[k] + c2 * gv[k];
        bj.dVdq0[k] = bj.dVdq0[k] - c0 * (gv[k] - h * dgv[k]);
        bj.dVdq1[k] = bj.dVdq1[k] - c1 * (gv[k] - 0.5 * h * dgv[k]);
        bj.dVdq2[k] = bj.dVdq2[k] - c2 * gv[k];
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    ABody &b = bs[i];
    for (int j = 0; j < 3; ++j) fdot[j] = b.dVdq0[j] - b.dVdq1[j] + 3.0 * b.dVdq2[j];
    assign_timestep(b, 1e-2 * pow(sf, 0.33333) * b.h * ob_norm(b.dVdq2) / ob_norm(fdot));
  }
}

bool first_step(const ABody &b) { return std::fpclassify(b.h) != FP_NORMAL; }

}  // namespace

extern "C" {

// Advancer.A.advance ?extpot bs dt sf (advancer.ml:422-428) on the full record state.
// state[i*ORACLE_ADV_STRIDE ..] = {t, m, q[3], p[3], h, hmax, t0, dVdq0[3], dVdq1[3],
// dVdq2[3], p0[3], q0[3], q1[3], q2[3], cs[3]} (35 doubles); fresh bodies (A.make,
// advancer.ml:98-116) carry NaN in everything after p.  Output order is the reference's
// (sorted by hmax); id[] is permuted alongside.  stats (may be NULL) receives
// {interact_bodies calls, grad_v evaluations}.
void oracle_advancer_advance(int n, int *id, double *state, double dt, double sf, double eps2,
                             oracle_extpot_fn extpot, void *extpot_arg, long *stats) {
  std::vector<ABody> bs(n);
  for (int i = 0; i < n; ++i) {
    const double *s = state + (size_t)i * ORACLE_ADV_STRIDE;
    ABody &b = bs[i];
    b.id = id[i];
    b.t = s[0];
    b.m = s[1];
    for (int k = 0; k < 3; ++k) {
      b.q[k] = s[2 + k];
      b.p[k] = s[5 + k];
      b.dVdq0[k] = s[11 + k];
      b.dVdq1[k] = s[14 + k];
      b.dVdq2[k] = s[17 + k];
      b.p0[k] = s[20 + k];
      b.q0[k] = s[23 + k];
      b.q1[k] = s[26 + k];
      b.q2[k] = s[29 + k];
      b.cs[k] = s[32 + k];
    }
    b.h = s[8];
    b.hmax = s[9];
    b.t0 = s[10];
  }
  Ctx cx = {eps2, extpot, extpot_arg, 0, 0};
  bool any_first = false;
  for (auto &b : bs) any_first = any_first || first_step(b);
  if (any_first) initialize_before_first_step(cx, bs, sf);
  std::stable_sort(bs.begin(), bs.end(), hmax_less);
  advance_range(cx, bs, n, dt, sf);
  for (int i = 0; i < n; ++i) {
    double *s = state + (size_t)i * ORACLE_ADV_STRIDE;
    const ABody &b = bs[i];
    id[i] = b.id;
    s[0] = b.t;
    s[1] = b.m;
    for (int k = 0; k < 3; ++k) {
      s[2 + k] = b.q[k];
      s[5 + k] = b.p[k];
      s[11 + k] = b.dVdq0[k];
      s[14 + k] = b.dVdq1[k];
      s[17 + k] = b.dVdq2[k];
      s[20 + k] = b.p0[k];
      s[23 + k] = b.q0[k];
      s[26 + k] = b.q1[k];
      s[29 + k] = b.q2[k];
      s[32 + k] = b.cs[k];
    }
    s[8] = b.h;
    s[9] = b.hmax;
    s[10] = b.t0;
  }
  if (stats) {
    stats[0] = cx.n_interact;
    stats[1] = cx.n_pairs;
  }
}

// The external field of test/advance_test.ml:39-46: Plummer potential, M = 1, eps = 0.1:
// gv = m x / (r^2 + 0.01)^(3/2);  v = -m / sqrt(r^2 + 0.01).
void oracle_extpot_plummer_test(void *arg, double m, const double *q, const double *p,
                                double *gv) {
  (void)arg;
  (void)p;
  double r = ob_norm(q);
  double r2 = r * r + 0.01;
  double r3 = r2 * sqrt(r2);
  for (int k = 0; k < 3; ++k) gv[k] = m * q[k] / r3;
}

double oracle_extpot_plummer_test_energy(int n, const double *m, const double *q) {
  double e = 0.0;
  for (int i = 0; i < n; ++i) {
    double r = ob_norm(q + 3 * i);
    double r2 = r * r + 0.01;
    e = e + (-m[i]) / sqrt(r2);
  }
  return e;
}

}  // extern "C"
