// hermite.cpp — host mirror of Hermite (hermite.ml:53-189).  The scheduler (a list ordered by
// next time, List.merge re-insertion, hermite.ml:171-178), the corrector (:116-123) and the
// time-step formulas (:87-93, :145-146) are host logic exactly as in the reference; the two
// pair loops run on the GPU: start_bs' all-pairs acc+jerk (:154-167) through
// nbk_grad_v_dgrad_v_sum, and advance_body's 1 x (N-1) loop with on-the-fly prediction
// (:106-115) through nbk_hermite_body_sum on device-resident state - one kernel launch and a
// 48-byte read-back per step; the stepped body's new state rides along with the next launch.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.hpp"

namespace {

using namespace nbh;

struct HBody {
  int id, slot;  // slot = row in the device-resident state
  double t, m, q[3], p[3], h, f[3], df[3];
};

double next_time(const HBody &b) { return b.t + b.h; }

int compare_next_time(const HBody &a, const HBody &b) {  // hermite.ml:77-82
  double n1 = next_time(a), n2 = next_time(b);
  return n1 < n2 ? -1 : (n1 > n2 ? 1 : 0);
}

void predict_position_into(double *qp, const HBody &b, double t) {  // hermite.ml:53-62
  double dt = t - b.t;
  double pc = dt / b.m;
  double fc = pc * dt * 0.5;
  double dfc = fc * dt / 3.0;
  for (int i = 0; i < 3; ++i) qp[i] = b.q[i] + pc * b.p[i] + fc * b.f[i] + dfc * b.df[i];
}

double timescale(double eta, double m, const double *q, const double *qp, const double *f,
                 double h) {  // hermite.ml:87-93
  double r = distance(q, qp);
  double fn = norm(f);
  if (r == 0.0) return h * std::pow(2.0, 0.25);
  return h * std::pow(eta * square(h) * fn / (m * r), 0.25);
}

long g_step_cap = 1L << 40;  // guard against the reference's runaway step-size collapse
int g_hermite_mode = 0;  // 0 auto, 1 host scheduler (launch per step), 2 device-resident loop

struct Herm {
  nbk_ctx *ctx;
  double eps2;
  int pending = -1;       // slot whose new state has not reached the device yet
  double pending_state[13];

  void stage(const HBody &b) {
    pending = b.slot;
    pending_state[0] = b.t;
    for (int k = 0; k < 3; ++k) {
      pending_state[1 + k] = b.q[k];
      pending_state[4 + k] = b.p[k];
      pending_state[7 + k] = b.f[k];
      pending_state[10 + k] = b.df[k];
    }
  }

  // hermite.ml:95-130; the List.iter over the others is the GPU call.
  int advance_body(double eta, HBody &b, long *nsteps) {
    const double nt = next_time(b);
    double qp[3], g[3], dg[3], fp[3], dfp[3];
    predict_position_into(qp, b, nt);
    NBH_TRY(ck(ctx, nbk_hermite_body_sum(ctx, b.slot, nt, eps2, pending, pending_state, g, dg)));
    pending = -1;
    for (int i = 0; i < 3; ++i) {
      fp[i] = -g[i];
      dfp[i] = -dg[i];
    }
    const double h = b.h, m = b.m;
    double p_new[3], q_new[3];
    for (int i = 0; i < 3; ++i) {
      p_new[i] = b.p[i] + 0.5 * h * (b.f[i] + fp[i]) + square(h) / 12.0 * (b.df[i] - dfp[i]);
      q_new[i] = b.q[i] + 0.5 * h / m * (b.p[i] + p_new[i]) +
                 square(h) / (12.0 * m) * (b.f[i] - fp[i]);
    }
    ++*nsteps;
    b.t = b.t + h;
    for (int i = 0; i < 3; ++i) {
      b.q[i] = q_new[i];
      b.p[i] = p_new[i];
      b.f[i] = fp[i];
      b.df[i] = dfp[i];
    }
    b.h = timescale(eta, m, q_new, qp, fp, h);
    stage(b);
    return 0;
  }
};

}  // namespace

extern "C" {

void nbh_set_hermite_mode(int mode) { g_hermite_mode = mode; }
void nbh_set_step_cap(long cap) { g_step_cap = cap; }

int nbh_hermite_advance(nbk_ctx *ctx, int n, int *id, double *t, double *m, double *q, double *p,
                        double *h, double *f, double *df, double dt, double sf, double eps2,
                        long *nsteps) {
  if (n <= 0) return fail(NBK_ERR_ARG, "Hermite.advance: empty body array");  // List.hd []
  std::vector<HBody> pool(n);
  for (int i = 0; i < n; ++i) {
    HBody &b = pool[i];
    b.id = id[i];
    b.slot = i;
    b.t = t[i];
    b.m = m[i];
    b.h = h[i];
    for (int k = 0; k < 3; ++k) {
      b.q[k] = q[3 * i + k];
      b.p[k] = p[3 * i + k];
      b.f[k] = f[3 * i + k];
      b.df[k] = df[3 * i + k];
    }
  }
  // start_bs (hermite.ml:148-169) when the first body has h < 0
  if (pool[0].h < 0.0) {
    std::vector<double> g(3 * (size_t)n), dg(3 * (size_t)n);
    NBH_TRY(ck(ctx, nbk_grad_v_dgrad_v_sum(ctx, n, m, q, p, eps2, 0, n, 0, n, g.data(), dg.data())));
    for (int i = 0; i < n; ++i) {
      HBody &b = pool[i];
      for (int k = 0; k < 3; ++k) {
        b.f[k] = -g[3 * i + k];
        b.df[k] = -dg[3 * i + k];
      }
      b.h = 6.0 * sf * norm(b.f) / (b.m * norm(b.df));  // start_timescale, hermite.ml:145-146
    }
  }
  {  // device-resident image of the Hermite.b array
    std::vector<double> tt(n), mm(n), qq(3 * (size_t)n), pp(3 * (size_t)n), ff(3 * (size_t)n),
        dd(3 * (size_t)n);
    for (int i = 0; i < n; ++i) {
      tt[i] = pool[i].t;
      mm[i] = pool[i].m;
      for (int k = 0; k < 3; ++k) {
        qq[3 * i + k] = pool[i].q[k];
        pp[3 * i + k] = pool[i].p[k];
        ff[3 * i + k] = pool[i].f[k];
        dd[3 * i + k] = pool[i].df[k];
      }
    }
    std::vector<double> hh(n);
    for (int i = 0; i < n; ++i) hh[i] = pool[i].h;
    NBH_TRY(ck(ctx, nbk_hermite_upload(ctx, n, tt.data(), mm.data(), qq.data(), pp.data(),
                                       hh.data(), ff.data(), dd.data())));
  }
  // Small systems: the whole advance_bodies / finish_bs loop runs on the device (one
  // persistent kernel, nbk_hermite_advance_resident); larger ones step from the host scheduler
  // below with one launch per advance_body.
  if (g_hermite_mode != 1 && (g_hermite_mode == 2 || n <= nbk_hermite_resident_max())) {
    double tmin = pool[0].t + pool[0].h;
    int imin = 0;
    for (int i = 1; i < n; ++i)
      if (pool[i].t + pool[i].h < tmin) {
        tmin = pool[i].t + pool[i].h;
        imin = i;
      }
    const double stop = pool[imin].t + dt;  // (List.hd sorted).t +. dt, hermite.ml:188
    long ns = 0;
    NBH_TRY(ck(ctx, nbk_hermite_advance_resident(ctx, stop, sf, eps2, g_step_cap, &ns)));
    *nsteps += ns;
    NBH_TRY(ck(ctx, nbk_hermite_download(ctx, t, q, p, h, f, df)));
    // Array.fast_sort compare_id (hermite.ml:136): device rows are in input order
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return id[a] < id[b]; });
    bool sorted = true;
    for (int i = 0; i < n; ++i) sorted = sorted && perm[i] == i;
    if (!sorted) {
      auto permute1 = [&](double *x) {
        std::vector<double> c(x, x + n);
        for (int i = 0; i < n; ++i) x[i] = c[perm[i]];
      };
      auto permute3 = [&](double *x) {
        std::vector<double> c(x, x + 3 * (size_t)n);
        for (int i = 0; i < n; ++i)
          for (int k = 0; k < 3; ++k) x[3 * i + k] = c[3 * perm[i] + k];
      };
      std::vector<int> ic(id, id + n);
      for (int i = 0; i < n; ++i) id[i] = ic[perm[i]];
      permute1(t);
      permute1(m);
      permute1(h);
      permute3(q);
      permute3(p);
      permute3(f);
      permute3(df);
    }
    return 0;
  }
  // List.fast_sort compare_next_time (hermite.ml:187), stable
  std::vector<HBody *> order(n);
  for (int i = 0; i < n; ++i) order[i] = &pool[i];
  std::stable_sort(order.begin(), order.end(),
                   [](const HBody *a, const HBody *b) { return compare_next_time(*a, *b) < 0; });
  const double stop_time = order[0]->t + dt;
  Herm H;
  H.ctx = ctx;
  H.eps2 = eps2;
  // advance_bodies (hermite.ml:171-178)
  while (!(next_time(*order[0]) > stop_time)) {
    if (*nsteps > g_step_cap) return fail(NBK_ERR_STATE, "Hermite.advance: step cap reached");
    HBody *head = order[0];
    NBH_TRY(H.advance_body(sf, *head, nsteps));
    // List.merge compare_next_time [new_b] tl: before the first e with cmp(new_b, e) <= 0
    size_t k = 1;
    while (k < (size_t)n && compare_next_time(*head, *order[k]) > 0) ++k;
    std::memmove(order.data(), order.data() + 1, (k - 1) * sizeof(HBody *));
    order[k - 1] = head;
  }
  // finish_bs (hermite.ml:132-143): step every body to stop_time in list order
  for (int a = 0; a < n; ++a) {
    HBody *b = order[a];
    b->h = stop_time - b->t;
    NBH_TRY(H.advance_body(sf, *b, nsteps));
  }
  // Array.fast_sort compare_id (hermite.ml:136); done_bs is built newest-first
  std::vector<HBody *> done(order.rbegin(), order.rend());
  std::reverse(done.begin(), done.end());
  std::stable_sort(done.begin(), done.end(),
                   [](const HBody *a, const HBody *b) { return a->id < b->id; });
  for (int i = 0; i < n; ++i) {
    const HBody &b = *done[i];
    id[i] = b.id;
    t[i] = b.t;
    m[i] = b.m;
    h[i] = b.h;
    for (int k = 0; k < 3; ++k) {
      q[3 * i + k] = b.q[k];
      p[3 * i + k] = b.p[k];
      f[3 * i + k] = b.f[k];
      df[3 * i + k] = b.df[k];
    }
  }
  return 0;
}

}  // extern "C"
