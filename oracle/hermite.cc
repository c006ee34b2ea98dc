// oracle/hermite.cc — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/base.h header).
//
// Restatement of the 4th-order individual-time-step Hermite integrator:
//   hermite.ml:53-72   predict_position_into / predict_momentum_into
//   hermite.ml:74-85   next_time / compare_next_time / compare_id
//   hermite.ml:87-93   timescale
//   hermite.ml:95-130  advance_body         (1 x (N-1) acc+jerk with on-the-fly prediction)
//   hermite.ml:132-143 finish_bs
//   hermite.ml:145-169 start_timescale / start_bs   (all-pairs acc+jerk)
//   hermite.ml:171-189 advance_bodies / advance
// and, through the same template over the Dual number type, of dFloat_hermite.ml:18-209 and
// dFloat_pushforward.ml:39-76 (advance_wrapper / jacobian_advance).
//
// The reference keeps the bodies in an OCaml list ordered by next time; `others` below is
// that list, and every sum runs in list order, so T = double reproduces the reference's
// rounding exactly (given the same inputs and libm pow).
#include "oracle.h"
#include "num.hh"

#include <algorithm>
#include <cstring>
#include <vector>

// Guard for test runs: pathological (eta, state) pairs shrink h below the time's ulp and the
// reference's loop (hermite.ml:171-178) then never terminates.  Past the cap the step loops
// stop early (callers see t < stop time).  Not part of the reference's algorithm.
long g_oracle_step_cap = 200000000L;
extern "C" void oracle_set_step_cap(long cap) { g_oracle_step_cap = cap; }

namespace {

template <class T> struct HBody {
  int id;
  T t, m, q[3], p[3], h, f[3], df[3];
};

template <class T> T *predict_position_into(T qp[3], const HBody<T> &b, T t) {
  T dt = t - b.t;
  T pc = dt / b.m;
  T fc = pc * dt * T(0.5);
  T dfc = fc * dt / T(3.0);
  for (int i = 0; i < 3; ++i) qp[i] = b.q[i] + pc * b.p[i] + fc * b.f[i] + dfc * b.df[i];
  return qp;
}

template <class T> T *predict_momentum_into(T pp[3], const HBody<T> &b, T t) {
  T dt = t - b.t;
  T fc = dt;
  T dfc = fc * dt * T(0.5);
  for (int i = 0; i < 3; ++i) pp[i] = b.p[i] + fc * b.f[i] + dfc * b.df[i];
  return pp;
}

template <class T> T next_time(const HBody<T> &b) { return b.t + b.h; }

template <class T> int compare_next_time(const HBody<T> &b1, const HBody<T> &b2) {
  T nt1 = next_time(b1), nt2 = next_time(b2);
  return nt1 < nt2 ? -1 : (nt1 > nt2 ? 1 : 0);
}

template <class T> T timescale(T eta, T m, const T q[3], const T qp[3], const T f[3], T h) {
  T r = onum::distance(q, qp);
  T fn = onum::norm(f);
  if (r == T(0.0)) return h * pow(T(2.0), 0.25);
  return h * pow(eta * onum::square(h) * fn / (m * r), 0.25);
}

// hermite.ml:95-130; `others` in list order.
template <class T>
HBody<T> advance_body(T eps2, T eta, const HBody<T> &b, const HBody<T> *const *others,
                      size_t n_others, long *nsteps) {
  T nt = next_time(b);
  T qp[3], pp[3], fp[3] = {T(0.0), T(0.0), T(0.0)}, dfp[3] = {T(0.0), T(0.0), T(0.0)};
  T gv[3], dgv[3], qp2[3], pp2[3];
  predict_position_into(qp, b, nt);
  predict_momentum_into(pp, b, nt);
  for (size_t k = 0; k < n_others; ++k) {
    const HBody<T> &b2 = *others[k];
    predict_position_into(qp2, b2, nt);
    predict_momentum_into(pp2, b2, nt);
    onum::grad_v_dgrad_v(eps2, b.m, qp, pp, b2.m, qp2, pp2, gv, dgv);
    for (int i = 0; i < 3; ++i) {
      fp[i] = fp[i] - gv[i];
      dfp[i] = dfp[i] - dgv[i];
    }
  }
  HBody<T> nb = b;
  T h = b.h, m = b.m;
  for (int i = 0; i < 3; ++i) {
    nb.p[i] = b.p[i] + T(0.5) * h * (b.f[i] + fp[i]) +
              onum::square(h) / T(12.0) * (b.df[i] - dfp[i]);
    nb.q[i] = b.q[i] + T(0.5) * h / m * (b.p[i] + nb.p[i]) +
              onum::square(h) / (T(12.0) * m) * (b.f[i] - fp[i]);
  }
  ++*nsteps;
  nb.t = b.t + h;
  for (int i = 0; i < 3; ++i) {
    nb.f[i] = fp[i];
    nb.df[i] = dfp[i];
  }
  nb.h = timescale(eta, m, nb.q, qp, fp, h);
  return nb;
}

// hermite.ml:145-169
template <class T> void start_bs(T eps2, T eta, std::vector<HBody<T>> &bs) {
  T gv[3], dgv[3];
  for (auto &b : bs)
    for (int i = 0; i < 3; ++i) b.f[i] = b.df[i] = T(0.0);
  for (size_t a = 0; a < bs.size(); ++a) {
    HBody<T> &b = bs[a];
    for (size_t c = a + 1; c < bs.size(); ++c) {
      HBody<T> &b2 = bs[c];
      onum::grad_v_dgrad_v(eps2, b.m, b.q, b.p, b2.m, b2.q, b2.p, gv, dgv);
      for (int i = 0; i < 3; ++i) {
        b.f[i] = b.f[i] - gv[i];
        b2.f[i] = b2.f[i] + gv[i];
        b.df[i] = b.df[i] - dgv[i];
        b2.df[i] = b2.df[i] + dgv[i];
      }
    }
  }
  for (auto &b : bs) b.h = T(6.0) * eta * onum::norm(b.f) / (b.m * onum::norm(b.df));
}

// hermite.ml:132-143 and 171-189
template <class T>
std::vector<HBody<T>> advance(T eps2, std::vector<HBody<T>> bs, T dt, T sf, long *nsteps) {
  if (bs[0].h < T(0.0)) start_bs(eps2, sf, bs);
  std::stable_sort(bs.begin(), bs.end(), [](const HBody<T> &a, const HBody<T> &b) {
    return compare_next_time(a, b) < 0;
  });
  T stop_time = bs[0].t + dt;
  const size_t n = bs.size();
  // `order` is the OCaml list: pointers into a stable pool.
  std::vector<HBody<T>> pool(bs);
  std::vector<HBody<T> *> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = &pool[i];
  // advance_bodies
  while (!(next_time(*order[0]) > stop_time)) {
    if (*nsteps > g_oracle_step_cap) break;
    HBody<T> *head = order[0];
    HBody<T> nb = advance_body(eps2, sf, *head, order.data() + 1, n - 1, nsteps);
    *head = nb;
    // List.merge compare_next_time [new_b] tl: new_b goes before the first element e with
    // cmp(new_b, e) <= 0.
    size_t k = 1;
    while (k < n && compare_next_time(*head, *order[k]) > 0) ++k;
    std::memmove(order.data(), order.data() + 1, (k - 1) * sizeof(HBody<T> *));
    order[k - 1] = head;
  }
  // finish_bs: for b :: rest -> advance_body {b with h = stop_time - b.t} (rest @ done)
  std::vector<HBody<T>> done;  // most recently finished first (OCaml cons)
  done.reserve(n);
  std::vector<const HBody<T> *> others(n - 1);
  for (size_t a = 0; a < n; ++a) {
    HBody<T> b = *order[a];
    b.h = stop_time - b.t;
    size_t c = 0;
    for (size_t r = a + 1; r < n; ++r) others[c++] = order[r];
    for (size_t r = done.size(); r-- > 0;) others[c++] = &done[r];
    done.push_back(advance_body(eps2, sf, b, others.data(), n - 1, nsteps));
  }
  // done_bs as an OCaml list is newest first; Array.fast_sort compare_id is stable.
  std::reverse(done.begin(), done.end());
  std::stable_sort(done.begin(), done.end(),
                   [](const HBody<T> &a, const HBody<T> &b) { return a.id < b.id; });
  return done;
}

std::vector<HBody<double>> load(int n, const int *id, const double *t, const double *m,
                                const double *q, const double *p, const double *h,
                                const double *f, const double *df) {
  std::vector<HBody<double>> bs(n);
  for (int i = 0; i < n; ++i) {
    bs[i].id = id[i];
    bs[i].t = t[i];
    bs[i].m = m[i];
    bs[i].h = h[i];
    for (int k = 0; k < 3; ++k) {
      bs[i].q[k] = q[3 * i + k];
      bs[i].p[k] = p[3 * i + k];
      bs[i].f[k] = f[3 * i + k];
      bs[i].df[k] = df[3 * i + k];
    }
  }
  return bs;
}

}  // namespace

extern "C" {

// Hermite.advance bs dt sf on flat arrays (all in/out, length n or 3n); returns bodies in id
// order like hermite.ml:136.  *nsteps is bumped like Base.nsteps (hermite.ml:124).
void oracle_hermite_advance(int n, int *id, double *t, double *m, double *q, double *p,
                            double *h, double *f, double *df, double dt, double sf,
                            double eps2, long *nsteps) {
  auto out = advance<double>(eps2, load(n, id, t, m, q, p, h, f, df), dt, sf, nsteps);
  for (int i = 0; i < n; ++i) {
    id[i] = out[i].id;
    t[i] = out[i].t;
    m[i] = out[i].m;
    h[i] = out[i].h;
    for (int k = 0; k < 3; ++k) {
      q[3 * i + k] = out[i].q[k];
      p[3 * i + k] = out[i].p[k];
      f[3 * i + k] = out[i].f[k];
      df[3 * i + k] = out[i].df[k];
    }
  }
}

// Hermite.advance_body eta b bs for body `ib` against all others in array order
// (hermite.ml:95-130): writes the new body's t,q,p,f,df,h into out[0..13] =
// {t, h, q[3], p[3], f[3], df[3]}.
void oracle_hermite_advance_body(int n, const double *t, const double *m, const double *q,
                                 const double *p, const double *h, const double *f,
                                 const double *df, int ib, double eta, double eps2,
                                 double *out) {
  std::vector<int> id(n);
  for (int i = 0; i < n; ++i) id[i] = i;
  auto bs = load(n, id.data(), t, m, q, p, h, f, df);
  std::vector<const HBody<double> *> others;
  for (int i = 0; i < n; ++i)
    if (i != ib) others.push_back(&bs[i]);
  long ns = 0;
  HBody<double> nb = advance_body<double>(eps2, eta, bs[ib], others.data(), others.size(), &ns);
  out[0] = nb.t;
  out[1] = nb.h;
  for (int k = 0; k < 3; ++k) {
    out[2 + k] = nb.q[k];
    out[5 + k] = nb.p[k];
    out[8 + k] = nb.f[k];
    out[11 + k] = nb.df[k];
  }
}

// DFloat_hermite.advance_body eta b bs (dFloat_hermite.ml:115-150) for body `ib` against all
// others in array order, ONE tangent direction: every field of every body is a dual (value
// arrays as in oracle_hermite_advance_body, tangent arrays *_d of the same shapes); masses and
// eta are constants.  out / out_d[0..13] = {t, h, q[3], p[3], f[3], df[3]} of the new body:
// values and tangents.
void oracle_dfloat_hermite_advance_body(int n, const double *t, const double *m, const double *q,
                                        const double *p, const double *h, const double *f,
                                        const double *df, const double *t_d, const double *q_d,
                                        const double *p_d, const double *h_d, const double *f_d,
                                        const double *df_d, int ib, double eta, double eps2,
                                        double *out, double *out_d) {
  std::vector<HBody<Dual>> bs(n);
  for (int i = 0; i < n; ++i) {
    bs[i].id = i;
    bs[i].t = Dual(t[i], t_d[i]);
    bs[i].m = Dual(m[i]);
    bs[i].h = Dual(h[i], h_d[i]);
    for (int k = 0; k < 3; ++k) {
      bs[i].q[k] = Dual(q[3 * i + k], q_d[3 * i + k]);
      bs[i].p[k] = Dual(p[3 * i + k], p_d[3 * i + k]);
      bs[i].f[k] = Dual(f[3 * i + k], f_d[3 * i + k]);
      bs[i].df[k] = Dual(df[3 * i + k], df_d[3 * i + k]);
    }
  }
  std::vector<const HBody<Dual> *> others;
  for (int i = 0; i < n; ++i)
    if (i != ib) others.push_back(&bs[i]);
  long ns = 0;
  HBody<Dual> nb =
      advance_body<Dual>(Dual(eps2), Dual(eta), bs[ib], others.data(), others.size(), &ns);
  out[0] = nb.t.v, out_d[0] = nb.t.d;
  out[1] = nb.h.v, out_d[1] = nb.h.d;
  for (int k = 0; k < 3; ++k) {
    out[2 + k] = nb.q[k].v, out_d[2 + k] = nb.q[k].d;
    out[5 + k] = nb.p[k].v, out_d[5 + k] = nb.p[k].d;
    out[8 + k] = nb.f[k].v, out_d[8 + k] = nb.f[k].d;
    out[11 + k] = nb.df[k].v, out_d[11 + k] = nb.df[k].d;
  }
}

// Tangent of the rectangular acc+jerk sum (DFloatBase.grad_v_dgrad_v, dFloat_advancer.ml:66-81,
// accumulated like dFloat_hermite.ml:177-186) for ONE tangent direction (dq, dp); masses and
// eps2 are constants (dFloat_hermite.ml:36-46).  Outputs the value sums and their tangents.
void oracle_grad_v_dgrad_v_sum_tangent(int n, const double *m, const double *q, const double *p,
                                       const double *dq, const double *dp, double eps2, int t0,
                                       int t1, int s0, int s1, double *gsum, double *dgsum,
                                       double *gsum_tan, double *dgsum_tan) {
  (void)n;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = t0; i < t1; ++i) {
    Dual g[3], dg[3], gv[3], dgv[3], q1[3], p1[3], q2[3], p2[3];
    for (int k = 0; k < 3; ++k) {
      q1[k] = Dual(q[3 * i + k], dq[3 * i + k]);
      p1[k] = Dual(p[3 * i + k], dp[3 * i + k]);
    }
    for (int j = s0; j < s1; ++j) {
      if (j == i) continue;
      for (int k = 0; k < 3; ++k) {
        q2[k] = Dual(q[3 * j + k], dq[3 * j + k]);
        p2[k] = Dual(p[3 * j + k], dp[3 * j + k]);
      }
      onum::grad_v_dgrad_v(Dual(eps2), Dual(m[i]), q1, p1, Dual(m[j]), q2, p2, gv, dgv);
      for (int k = 0; k < 3; ++k) {
        g[k] = g[k] + gv[k];
        dg[k] = dg[k] + dgv[k];
      }
    }
    for (int k = 0; k < 3; ++k) {
      size_t o = 3 * (size_t)(i - t0) + k;
      gsum[o] = g[k].v;
      dgsum[o] = dg[k].v;
      gsum_tan[o] = g[k].d;
      dgsum_tan[o] = dg[k].d;
    }
  }
}

// Tangent of the plain gradient sum (DFloatBase.grad_v, dFloat_advancer.ml:57-64).
void oracle_grad_v_sum_tangent(int n, const double *m, const double *q, const double *dq,
                               double eps2, int t0, int t1, int s0, int s1, double *gsum,
                               double *gsum_tan) {
  (void)n;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = t0; i < t1; ++i) {
    Dual g[3], gv[3], q1[3], q2[3];
    for (int k = 0; k < 3; ++k) q1[k] = Dual(q[3 * i + k], dq[3 * i + k]);
    for (int j = s0; j < s1; ++j) {
      if (j == i) continue;
      for (int k = 0; k < 3; ++k) q2[k] = Dual(q[3 * j + k], dq[3 * j + k]);
      onum::grad_v(Dual(eps2), Dual(m[i]), q1, Dual(m[j]), q2, gv);
      for (int k = 0; k < 3; ++k) g[k] = g[k] + gv[k];
    }
    for (int k = 0; k < 3; ++k) {
      size_t o = 3 * (size_t)(i - t0) + k;
      gsum[o] = g[k].v;
      gsum_tan[o] = g[k].d;
    }
  }
}

// DFloat_pushforward.Make(DFloat_hermite).jacobian_advance dt sf bs
// (dFloat_pushforward.ml:39-76): bodies are fresh Hermite bodies (h = -1, f = df = 0,
// t = 0, ids 1..n); the (q,p) state is packed as [q(3n); p(3n)]; jac is 6n x 6n row-major,
// jac[r*6n + c] = d out[r] / d in[c], one forward pass per column.  out6n (may be NULL)
// receives the advanced packed state.
void oracle_hermite_jacobian_advance(int n, const double *m, const double *q, const double *p,
                                     double dt, double sf, double eps2, double *jac,
                                     double *out6n) {
  const int d6 = 6 * n;
  for (int c = 0; c < d6; ++c) {
    std::vector<HBody<Dual>> bs(n);
    for (int i = 0; i < n; ++i) {
      bs[i].id = i + 1;
      bs[i].t = Dual(0.0);
      bs[i].m = Dual(m[i]);
      bs[i].h = Dual(-1.0);
      for (int k = 0; k < 3; ++k) {
        bs[i].q[k] = Dual(q[3 * i + k], (c == 3 * i + k) ? 1.0 : 0.0);
        bs[i].p[k] = Dual(p[3 * i + k], (c == 3 * n + 3 * i + k) ? 1.0 : 0.0);
        bs[i].f[k] = bs[i].df[k] = Dual(0.0);
      }
    }
    long ns = 0;
    auto out = advance<Dual>(Dual(eps2), bs, Dual(dt), Dual(sf), &ns);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) {
        jac[(size_t)(3 * i + k) * d6 + c] = out[i].q[k].d;
        jac[(size_t)(3 * n + 3 * i + k) * d6 + c] = out[i].p[k].d;
        if (out6n && c == 0) {
          out6n[3 * i + k] = out[i].q[k].v;
          out6n[3 * n + 3 * i + k] = out[i].p[k].v;
        }
      }
  }
}

// dFloat_pushforward.ml:141-150: mean of (J^T Omega J)[i+3n][i] over i < 3n, Omega from
// `symplectic` (:132-139).  1.0 for a symplectic map.
double oracle_omega_pushforward(int n, const double *jac) {
  const int d6 = 6 * n, no2 = 3 * n;
  // S[i][no2+i] = -1, S[no2+i][i] = 1.
  double sum = 0.0;
  for (int i = 0; i < no2; ++i) {
    int a = i + no2, b = i;
    // (J^T *|| S) first (dFloat_pushforward.ml:147, *|| associates to the left): its row a is
    // J[no2+j][a] for j < no2 and -J[j-no2][a] beyond, exactly (one non-zero term per sum); then
    // the product with J accumulates over j = 0 .. 6n-1 in order (:81-95)
    double e = 0.0;
    for (int j = 0; j < no2; ++j) e += jac[(size_t)(no2 + j) * d6 + a] * jac[(size_t)j * d6 + b];
    for (int j = no2; j < d6; ++j) e += -jac[(size_t)(j - no2) * d6 + a] * jac[(size_t)j * d6 + b];
    sum += e;
  }
  return sum / (double)no2;
}

}  // extern "C"
