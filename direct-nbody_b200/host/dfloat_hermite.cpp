// dfloat_hermite.cpp — host mirror of DFloat_hermite (dFloat_hermite.ml:18-209) and of
// DFloat_pushforward.Make(DFloat_hermite) (dFloat_pushforward.ml:35-76, 141-150).
//
// The reference writes Hermite a second time over Diff.DFloat.diff: every field of a body
// (t, q, p, h, f, df) is a dual number; masses, eta (sf), dt and eps are constants (`C x`,
// dFloat_hermite.ml:36-46, dFloat_pushforward.ml:51).  Here a dual carries K tangents at once
// (`jacobian` needs the 6n unit directions of [q; p]; they do not interact, so one pass with
// K = 6n equals 6n passes with one), the scheduler (list ordered by next time, List.merge) and
// the corrector / time-step formulas are host logic as in the reference - comparisons look at
// values - and both pair loops run on the GPU:
//   start_bs' all-pairs loop (:168-189)     -> nbk_grad_v_dgrad_v_sum_tangent
//   advance_body's 1 x (N-1) loop (:126-135) -> nbk_grad_v_dgrad_v_sum_predicted_tangent
//     (prediction of every body, tangents of t, q, p, f, df included, happens on the device).
// Diff.DFloat itself is third-party and not in the reference tree: the
// tangent rules here are the textbook ones (sum, product, quotient, sqrt, x^c).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.hpp"

namespace {

using namespace nbh;

// value + K tangents; arithmetic by the forward-mode rules
struct VD {
  double v = 0.0;
  std::vector<double> d;
  VD() {}
  VD(double v_, int K) : v(v_), d((size_t)K, 0.0) {}
};
int KK(const VD &a, const VD &b) { return (int)std::max(a.d.size(), b.d.size()); }
double dk(const VD &a, int k) { return (size_t)k < a.d.size() ? a.d[k] : 0.0; }
VD cst(double v, int K) { return VD(v, K); }
VD operator+(const VD &a, const VD &b) {
  VD r(a.v + b.v, KK(a, b));
  for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = dk(a, (int)k) + dk(b, (int)k);
  return r;
}
VD operator-(const VD &a, const VD &b) {
  VD r(a.v - b.v, KK(a, b));
  for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = dk(a, (int)k) - dk(b, (int)k);
  return r;
}
VD operator*(const VD &a, const VD &b) {
  VD r(a.v * b.v, KK(a, b));
  for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = dk(a, (int)k) * b.v + a.v * dk(b, (int)k);
  return r;
}
VD operator/(const VD &a, const VD &b) {
  const double q = a.v / b.v;
  VD r(q, KK(a, b));
  for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = (dk(a, (int)k) - q * dk(b, (int)k)) / b.v;
  return r;
}
VD vsqrt(const VD &a) {
  const double s = std::sqrt(a.v);
  VD r(s, (int)a.d.size());
  for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = a.d[k] / (2.0 * s);
  return r;
}
VD vpow(const VD &a, double c) {
  VD r(std::pow(a.v, c), (int)a.d.size());
  const double g = c * std::pow(a.v, c - 1.0);
  for (size_t k = 0; k < r.d.size(); ++k) r.d[k] = g * a.d[k];
  return r;
}
VD vsquare(const VD &a) { return a * a; }
VD vnorm(const VD *x) { return vsqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]); }
VD vdistance(const VD *x, const VD *y) {
  VD dx = x[0] - y[0], dy = x[1] - y[1], dz = x[2] - y[2];
  return vsqrt(dx * dx + dy * dy + dz * dz);
}

struct DBody {  // dFloat_hermite.ml:22-31
  int id, slot;
  double m;
  VD t, q[3], p[3], h, f[3], df[3];
};

VD next_time(const DBody &b) { return b.t + b.h; }
int compare_next_time(const DBody &a, const DBody &b) {  // dFloat_hermite.ml:86-90, on values
  const double n1 = a.t.v + a.h.v, n2 = b.t.v + b.h.v;
  return n1 < n2 ? -1 : (n1 > n2 ? 1 : 0);
}

// dFloat_hermite.ml:60-70 for the stepping body itself (the corrector needs qp)
void predict_position_into(VD *qp, const DBody &b, const VD &t, int K) {
  VD dt = t - b.t;
  VD pc = dt / cst(b.m, K);
  VD fc = pc * dt * cst(0.5, K);
  VD dfc = fc * dt / cst(3.0, K);
  for (int i = 0; i < 3; ++i) qp[i] = b.q[i] + pc * b.p[i] + fc * b.f[i] + dfc * b.df[i];
}

VD timescale(double eta, double m, const VD *q, const VD *qp, const VD *f, const VD &h, int K) {
  // dFloat_hermite.ml:96-103
  VD r = vdistance(q, qp);
  VD fn = vnorm(f);
  if (r.v == 0.0) return h * vpow(cst(2.0, K), 0.25);
  return h * vpow(cst(eta, K) * vsquare(h) * fn / (cst(m, K) * r), 0.25);
}

long g_step_cap_d = 1L << 40;

// Flat device-facing image of the dual state: values [n] / [3n], tangents [K][n] / [K][3n].
struct Flat {
  int n, K;
  std::vector<double> t, m, q, p, f, df, t_tan, q_tan, p_tan, f_tan, df_tan;
  Flat(int n_, int K_)
      : n(n_), K(K_), t(n_), m(n_), q(3 * (size_t)n_), p(3 * (size_t)n_), f(3 * (size_t)n_),
        df(3 * (size_t)n_), t_tan((size_t)K_ * n_), q_tan((size_t)K_ * 3 * n_),
        p_tan((size_t)K_ * 3 * n_), f_tan((size_t)K_ * 3 * n_), df_tan((size_t)K_ * 3 * n_) {}
  void put(const DBody &b) {
    const int i = b.slot;
    t[i] = b.t.v;
    m[i] = b.m;
    for (int k = 0; k < K; ++k) t_tan[(size_t)k * n + i] = dk(b.t, k);
    for (int c = 0; c < 3; ++c) {
      q[3 * i + c] = b.q[c].v;
      p[3 * i + c] = b.p[c].v;
      f[3 * i + c] = b.f[c].v;
      df[3 * i + c] = b.df[c].v;
      for (int k = 0; k < K; ++k) {
        const size_t o = (size_t)k * 3 * n + 3 * (size_t)i + c;
        q_tan[o] = dk(b.q[c], k);
        p_tan[o] = dk(b.p[c], k);
        f_tan[o] = dk(b.f[c], k);
        df_tan[o] = dk(b.df[c], k);
      }
    }
  }
};

struct DHerm {
  nbk_ctx *ctx;
  double eps2;
  int K;
  Flat *flat;
  std::vector<double> nt_tan, gt, dgt;

  // dFloat_hermite.ml:115-150; the List.iter over the others (:126-135) is the GPU call
  int advance_body(double eta, DBody &b, long *nsteps) {
    const int n = flat->n;
    const VD nt = next_time(b);
    VD qp[3];
    predict_position_into(qp, b, nt, K);
    nt_tan.assign((size_t)K, 0.0);
    for (int k = 0; k < K; ++k) nt_tan[k] = dk(nt, k);
    gt.resize((size_t)K * 3);
    dgt.resize((size_t)K * 3);
    double g[3], dg[3];
    NBH_TRY(ck(ctx, nbk_grad_v_dgrad_v_sum_predicted_tangent(
                        ctx, n, flat->t.data(), flat->m.data(), flat->q.data(), flat->p.data(),
                        flat->f.data(), flat->df.data(), K, flat->t_tan.data(),
                        flat->q_tan.data(), flat->p_tan.data(), flat->f_tan.data(),
                        flat->df_tan.data(), nt.v, nt_tan.data(), eps2, b.slot, b.slot + 1, 0, n,
                        g, dg, gt.data(), dgt.data())));
    VD fp[3], dfp[3];
    for (int i = 0; i < 3; ++i) {
      fp[i] = VD(-g[i], K);
      dfp[i] = VD(-dg[i], K);
      for (int k = 0; k < K; ++k) {
        fp[i].d[k] = -gt[(size_t)k * 3 + i];
        dfp[i].d[k] = -dgt[(size_t)k * 3 + i];
      }
    }
    const VD h = b.h;
    const VD m = cst(b.m, K);
    VD p_new[3], q_new[3];
    for (int i = 0; i < 3; ++i) {
      p_new[i] = b.p[i] + cst(0.5, K) * h * (b.f[i] + fp[i]) +
                 vsquare(h) / cst(12.0, K) * (b.df[i] - dfp[i]);
      q_new[i] = b.q[i] + cst(0.5, K) * h / m * (b.p[i] + p_new[i]) +
                 vsquare(h) / (cst(12.0, K) * m) * (b.f[i] - fp[i]);
    }
    ++*nsteps;
    b.t = b.t + h;
    for (int i = 0; i < 3; ++i) {
      b.q[i] = q_new[i];
      b.p[i] = p_new[i];
      b.f[i] = fp[i];
      b.df[i] = dfp[i];
    }
    b.h = timescale(eta, b.m, q_new, qp, fp, h, K);
    flat->put(b);
    return 0;
  }
};

// DFloat_hermite.advance bs dt sf (dFloat_hermite.ml:199-209) on a pool of dual bodies
int dfloat_advance(nbk_ctx *ctx, std::vector<DBody> &pool, int K, double dt, double sf,
                   double eps2, long *nsteps) {
  const int n = (int)pool.size();
  Flat flat(n, K);
  for (auto &b : pool) flat.put(b);
  // start_bs (dFloat_hermite.ml:168-189) when the first body has h < 0
  if (pool[0].h.v < 0.0) {
    std::vector<double> g(3 * (size_t)n), dg(3 * (size_t)n), gt((size_t)K * 3 * n),
        dgt((size_t)K * 3 * n);
    NBH_TRY(ck(ctx, nbk_grad_v_dgrad_v_sum_tangent(ctx, n, flat.m.data(), flat.q.data(),
                                                   flat.p.data(), K, flat.q_tan.data(),
                                                   flat.p_tan.data(), eps2, 0, n, 0, n, g.data(),
                                                   dg.data(), gt.data(), dgt.data())));
    for (int i = 0; i < n; ++i) {
      DBody &b = pool[i];
      for (int c = 0; c < 3; ++c) {
        b.f[c] = VD(-g[3 * i + c], K);
        b.df[c] = VD(-dg[3 * i + c], K);
        for (int k = 0; k < K; ++k) {
          b.f[c].d[k] = -gt[(size_t)k * 3 * n + 3 * (size_t)i + c];
          b.df[c].d[k] = -dgt[(size_t)k * 3 * n + 3 * (size_t)i + c];
        }
      }
      // start_timescale (dFloat_hermite.ml:164-166)
      b.h = cst(6.0, K) * cst(sf, K) * vnorm(b.f) / (cst(b.m, K) * vnorm(b.df));
      flat.put(b);
    }
  }
  std::vector<DBody *> order(n);
  for (int i = 0; i < n; ++i) order[i] = &pool[i];
  std::stable_sort(order.begin(), order.end(),
                   [](const DBody *a, const DBody *b) { return compare_next_time(*a, *b) < 0; });
  const VD stop_time = order[0]->t + cst(dt, K);
  DHerm H;
  H.ctx = ctx;
  H.eps2 = eps2;
  H.K = K;
  H.flat = &flat;
  // advance_bodies (dFloat_hermite.ml:191-197)
  while (!(next_time(*order[0]).v > stop_time.v)) {
    if (*nsteps > g_step_cap_d)
      return fail(NBK_ERR_STATE, "DFloat_hermite.advance: step cap reached");
    DBody *head = order[0];
    NBH_TRY(H.advance_body(sf, *head, nsteps));
    size_t k = 1;
    while (k < (size_t)n && compare_next_time(*head, *order[k]) > 0) ++k;
    std::memmove(order.data(), order.data() + 1, (k - 1) * sizeof(DBody *));
    order[k - 1] = head;
  }
  // finish_bs (dFloat_hermite.ml:152-162)
  for (int a = 0; a < n; ++a) {
    DBody *b = order[a];
    b->h = stop_time - b->t;
    NBH_TRY(H.advance_body(sf, *b, nsteps));
  }
  return 0;
}

}  // namespace

extern "C" {

void nbh_set_dfloat_step_cap(long cap) { g_step_cap_d = cap; }

// DFloat_hermite.advance over K tangent directions of (q, p): bodies are fresh Hermite bodies
// promoted with promote_body (t = 0 .. as given, h = -1 => start_bs runs first) unless h[0] >= 0.
int nbh_dfloat_hermite_advance(nbk_ctx *ctx, int n, int K, const int *id, double *t, const double *m,
                               double *q, double *p, double *h, double *f, double *df,
                               double *t_tan, double *q_tan, double *p_tan, double *h_tan,
                               double *f_tan, double *df_tan, double dt, double sf, double eps2,
                               long *nsteps) {
  if (n <= 0 || K < 0) return fail(NBK_ERR_ARG, "DFloat_hermite.advance: empty body array");
  std::vector<DBody> pool(n);
  for (int i = 0; i < n; ++i) {
    DBody &b = pool[i];
    b.id = id[i];
    b.slot = i;
    b.m = m[i];
    b.t = VD(t[i], K);
    b.h = VD(h[i], K);
    for (int k = 0; k < K; ++k) {
      if (t_tan) b.t.d[k] = t_tan[(size_t)k * n + i];
      if (h_tan) b.h.d[k] = h_tan[(size_t)k * n + i];
    }
    for (int c = 0; c < 3; ++c) {
      b.q[c] = VD(q[3 * i + c], K);
      b.p[c] = VD(p[3 * i + c], K);
      b.f[c] = VD(f[3 * i + c], K);
      b.df[c] = VD(df[3 * i + c], K);
      for (int k = 0; k < K; ++k) {
        const size_t o = (size_t)k * 3 * n + 3 * (size_t)i + c;
        if (q_tan) b.q[c].d[k] = q_tan[o];
        if (p_tan) b.p[c].d[k] = p_tan[o];
        if (f_tan) b.f[c].d[k] = f_tan[o];
        if (df_tan) b.df[c].d[k] = df_tan[o];
      }
    }
  }
  NBH_TRY(dfloat_advance(ctx, pool, K, dt, sf, eps2, nsteps));
  // Array.fast_sort compare_id (dFloat_hermite.ml:155-157)
  std::vector<DBody *> done(n);
  for (int i = 0; i < n; ++i) done[i] = &pool[i];
  std::stable_sort(done.begin(), done.end(),
                   [](const DBody *a, const DBody *b) { return a->id < b->id; });
  // outputs are written in id order; the caller's id[] is const and, for bodies made in order,
  // already sorted - rows follow the sorted order
  for (int i = 0; i < n; ++i) {
    const DBody &b = *done[i];
    t[i] = b.t.v;
    h[i] = b.h.v;
    for (int k = 0; k < K; ++k) {
      if (t_tan) t_tan[(size_t)k * n + i] = dk(b.t, k);
      if (h_tan) h_tan[(size_t)k * n + i] = dk(b.h, k);
    }
    for (int c = 0; c < 3; ++c) {
      q[3 * i + c] = b.q[c].v;
      p[3 * i + c] = b.p[c].v;
      f[3 * i + c] = b.f[c].v;
      df[3 * i + c] = b.df[c].v;
      for (int k = 0; k < K; ++k) {
        const size_t o = (size_t)k * 3 * n + 3 * (size_t)i + c;
        if (q_tan) q_tan[o] = dk(b.q[c], k);
        if (p_tan) p_tan[o] = dk(b.p[c], k);
        if (f_tan) f_tan[o] = dk(b.f[c], k);
        if (df_tan) df_tan[o] = dk(b.df[c], k);
      }
    }
  }
  return 0;
}

// DFloat_pushforward.Make(DFloat_hermite).jacobian_advance dt sf bs (dFloat_pushforward.ml:39-76):
// fresh Hermite bodies (t = 0, h = -1, f = df = 0, ids 1..n); the state is packed as
// [q(3n); p(3n)]; jac is 6n x 6n row-major, jac[r*6n + c] = d out[r] / d in[c]; out6n (may be
// NULL) receives the advanced packed state.  One pass with the 6n unit directions.
int nbh_dfloat_jacobian_advance(nbk_ctx *ctx, int n, const double *m, const double *q,
                                const double *p, double dt, double sf, double eps2, double *jac,
                                double *out6n, long *nsteps) {
  if (n <= 0 || !jac) return fail(NBK_ERR_ARG, "jacobian_advance: bad arguments");
  const int K = 6 * n;
  std::vector<DBody> pool(n);
  for (int i = 0; i < n; ++i) {
    DBody &b = pool[i];
    b.id = i + 1;
    b.slot = i;
    b.m = m[i];
    b.t = VD(0.0, K);
    b.h = VD(-1.0, K);
    for (int c = 0; c < 3; ++c) {
      b.q[c] = VD(q[3 * i + c], K);
      b.p[c] = VD(p[3 * i + c], K);
      b.q[c].d[3 * i + c] = 1.0;
      b.p[c].d[3 * n + 3 * i + c] = 1.0;
      b.f[c] = VD(0.0, K);
      b.df[c] = VD(0.0, K);
    }
  }
  long ns = 0;
  NBH_TRY(dfloat_advance(ctx, pool, K, dt, sf, eps2, &ns));
  if (nsteps) *nsteps += ns;
  for (int i = 0; i < n; ++i) {  // ids are 1..n in slot order: already the compare_id order
    const DBody &b = pool[i];
    for (int c = 0; c < 3; ++c) {
      for (int k = 0; k < K; ++k) {
        jac[(size_t)(3 * i + c) * K + k] = dk(b.q[c], k);
        jac[(size_t)(3 * n + 3 * i + c) * K + k] = dk(b.p[c], k);
      }
      if (out6n) {
        out6n[3 * i + c] = b.q[c].v;
        out6n[3 * n + 3 * i + c] = b.p[c].v;
      }
    }
  }
  return 0;
}

// omega_pushforward (dFloat_pushforward.ml:141-150): mean over i < 3n of (J^T Omega J)[i+3n][i]
// with Omega = `symplectic` (:132-139); 1 for a symplectic map.  Plain host algebra on the
// Jacobian (the dense 6n x 6n products are out of scope for the GPU path).
double nbh_omega_pushforward(int n, const double *jac) {
  const int d6 = 6 * n, no2 = 3 * n;
  double sum = 0.0;
  for (int i = 0; i < no2; ++i) {
    const int a = i + no2, b = i;
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
