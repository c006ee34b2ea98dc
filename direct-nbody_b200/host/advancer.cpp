// advancer.cpp — host mirror of Advancer.A (advancer.ml:143-428) and of
// Integrator.Make(B)(A).constant_sf_integrate (integrator.ml:58-84) over it.
// Predictors, the Simpson-rule corrector, time-step control, block recursion and sorts are host
// logic with the reference's association order; the pair loops run on the GPU:
//   initialize_before_first_step (:369-393)  -> nbk_grad_v_dgrad_v_sum, then
//       dVdq0 = c0 (g - h dg), dVdq1 = c1 (g - h/2 dg), dVdq2 = c2 g
//   interact_bodies (:207-236)               -> nbk_interact_sum = two rectangular sweeps:
//       active targets [imin,imax) x sources [imin,n); passive targets [imax,n) x [imin,imax);
//       dVdqK_x += (vC cs_x[K]) S_x    (grad_v is antisymmetric bit for bit)
// Only bodies [imin,n) are shipped per call.  The follow_flag print branch (:343-347) is
// output only and omitted.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include "common.hpp"

namespace {

using namespace nbh;

struct ABody {
  int id;
  double t, m, q[3], p[3], h, hmax, t0;
  double dVdq0[3], dVdq1[3], dVdq2[3], p0[3], q0[3], q1[3], q2[3], cs[3];
};

struct Adv {
  nbk_ctx *ctx;
  double eps2;
  nbh_extpot_fn extpot;
  void *extpot_arg;
  long n_interact = 0, n_pairs = 0;
  std::vector<ABody> bs;
  std::vector<double> m, q, S;

  static void predict_q012(ABody &b, double hnew) {  // advancer.ml:143-164
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

  static void clear_before_step(ABody &b) {  // advancer.ml:166-171
    for (int i = 0; i < 3; ++i) b.dVdq0[i] = b.dVdq1[i] = b.dVdq2[i] = b.cs[i] = 0.0;
    b.cs[0] = 1.0;
  }

  static void predict_position(ABody &b, double t) {  // advancer.ml:173-195
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

  static void predict_all(ABody &b, double t) {  // advancer.ml:197-205
    predict_position(b, t);
    double dt = t - b.t0, h = b.h;
    double c0 = (4.0 * dt - 3.0 * h) / (h * h), c1 = 4.0 * (h - 2.0 * dt) / (h * h),
           c2 = (4.0 * dt - h) / (h * h);
    for (int i = 0; i < 3; ++i) b.p[i] = b.m * (c0 * b.q0[i] + c1 * b.q1[i] + c2 * b.q2[i]);
  }

  int interact_bodies(int imin, int imax, double t, double vC) {  // advancer.ml:207-236
    const int n = (int)bs.size();
    for (int i = imin; i < n; ++i) predict_position(bs[i], t);
    const int nn = n - imin, na = imax - imin;  // shipped bodies, active bodies
    m.resize(nn);
    q.resize(3 * (size_t)nn);
    S.resize(3 * (size_t)nn);
    for (int i = 0; i < nn; ++i) {
      m[i] = bs[imin + i].m;
      for (int k = 0; k < 3; ++k) q[3 * i + k] = bs[imin + i].q[k];
    }
    NBH_TRY(ck(ctx, nbk_interact_sum(ctx, nn, na, m.data(), q.data(), eps2, S.data())));
    for (int i = 0; i < nn; ++i) {
      ABody &b = bs[imin + i];
      const double c0 = vC * b.cs[0], c1 = vC * b.cs[1], c2 = vC * b.cs[2];
      for (int k = 0; k < 3; ++k) {
        const double s = S[3 * (size_t)i + k];
        b.dVdq0[k] = b.dVdq0[k] + c0 * s;
        b.dVdq1[k] = b.dVdq1[k] + c1 * s;
        b.dVdq2[k] = b.dVdq2[k] + c2 * s;
      }
    }
    n_interact += 1;
    n_pairs += (long)na * (nn - 1) - (long)na * (na - 1) / 2;
    return 0;
  }

  static void finish_body(ABody &b, double t) {  // advancer.ml:238-252
    double cp = b.h / b.m;
    double cV0 = cp, cV1 = 0.5 * cp;
    for (int i = 0; i < 3; ++i) {
      double v0i = b.dVdq0[i], v1i = b.dVdq1[i], v2i = b.dVdq2[i], pi = b.p0[i];
      b.q[i] = b.q0[i] + cp * pi - cV0 * v0i - cV1 * v1i;
      b.p[i] = pi - v0i - v1i - v2i;
    }
    b.t = t;
  }

  static void assign_timestep(ABody &b, double dt) {  // advancer.ml:257-266 (warning omitted)
    if (dt < 100.0 * DBL_EPSILON * std::fabs(b.t) + 100.0 * DBL_EPSILON)
      b.hmax = 100.0 * DBL_EPSILON * b.t + 100.0 * DBL_EPSILON;
    else
      b.hmax = dt;
  }

  static void h_adaptive(double sf, ABody &b) {  // advancer.ml:272-281
    double h = b.h;
    double dq = distance(b.q, b.q2);
    if (dq == 0.0) {
      assign_timestep(b, h * 2.01);
    } else {
      double a = norm(b.dVdq2) * 6.0 / h / b.m;
      double a_dq = 0.5 * h * h * a;
      double ratio = a_dq / dq;
      assign_timestep(b, h * std::pow(sf * ratio, 0.3333333));
    }
  }

  static bool hmax_less(const ABody &a, const ABody &b) { return a.hmax < b.hmax; }

  void do_external_potential(ABody &b) {  // advancer.ml:299-314
    if (!extpot) return;
    double h = b.h, gv0[3], gv1[3], gv2[3];
    predict_all(b, b.t0);
    extpot(extpot_arg, b.m, b.q, b.p, gv0);
    predict_all(b, b.t0 + 0.5 * b.h);
    extpot(extpot_arg, b.m, b.q, b.p, gv1);
    predict_all(b, b.t0 + b.h);
    extpot(extpot_arg, b.m, b.q, b.p, gv2);
    for (int i = 0; i < 3; ++i) {
      b.dVdq0[i] = b.dVdq0[i] + h * gv0[i] / 6.0;
      b.dVdq1[i] = b.dVdq1[i] + 2.0 * h * gv1[i] / 3.0;
      b.dVdq2[i] = b.dVdq2[i] + h * gv2[i] / 6.0;
    }
  }

  int advance_range(int imax, double dt, double sf) {  // advancer.ml:316-350
    if (bs[imax - 1].hmax < dt) {
      NBH_TRY(advance_range(imax, dt / 2.0, sf));
      NBH_TRY(advance_range(imax, dt / 2.0, sf));
      return 0;
    }
    int imin = imax - 1;
    while (imin >= 0 && dt < bs[imin].hmax) --imin;
    imin = imin + 1;
    for (int i = imin; i < imax; ++i) {
      predict_q012(bs[i], dt);
      clear_before_step(bs[i]);
    }
    const double t0 = bs[imin].t;
    NBH_TRY(interact_bodies(imin, imax, t0, dt / 6.0));
    if (imin > 0) NBH_TRY(advance_range(imin, dt / 2.0, sf));
    NBH_TRY(interact_bodies(imin, imax, t0 + 0.5 * dt, 2.0 * dt / 3.0));
    if (imin > 0) NBH_TRY(advance_range(imin, dt / 2.0, sf));
    NBH_TRY(interact_bodies(imin, imax, t0 + dt, dt / 6.0));
    for (int i = imin; i < imax; ++i) {
      do_external_potential(bs[i]);
      finish_body(bs[i], t0 + dt);
      h_adaptive(sf, bs[i]);
    }
    std::stable_sort(bs.begin(), bs.begin() + imax, hmax_less);  // sort_range, :290-297
    return 0;
  }

  int initialize_before_first_step(double sf) {  // advancer.ml:357-406
    const int n = (int)bs.size();
    const double h = 1.0;
    const double c0 = h / 6.0, c1 = 2.0 * h / 3.0;
    const double c2 = c0;
    std::vector<double> mm(n), qq(3 * (size_t)n), pp(3 * (size_t)n), g(3 * (size_t)n),
        dg(3 * (size_t)n);
    for (int i = 0; i < n; ++i) {
      clear_before_step(bs[i]);
      bs[i].h = h;
      mm[i] = bs[i].m;
      for (int k = 0; k < 3; ++k) {
        qq[3 * i + k] = bs[i].q[k];
        pp[3 * i + k] = bs[i].p[k];
      }
    }
    NBH_TRY(ck(ctx, nbk_grad_v_dgrad_v_sum(ctx, n, mm.data(), qq.data(), pp.data(), eps2, 0, n, 0,
                                           n, g.data(), dg.data())));
    for (int i = 0; i < n; ++i) {
      ABody &b = bs[i];
      double fdot[3];
      for (int k = 0; k < 3; ++k) {
        const double gv = g[3 * i + k], dgv = dg[3 * i + k];
        b.dVdq0[k] = c0 * (gv - h * dgv);
        b.dVdq1[k] = c1 * (gv - 0.5 * h * dgv);
        b.dVdq2[k] = c2 * gv;
      }
      for (int j = 0; j < 3; ++j) fdot[j] = b.dVdq0[j] - b.dVdq1[j] + 3.0 * b.dVdq2[j];
      assign_timestep(b, 1e-2 * std::pow(sf, 0.33333) * b.h * norm(b.dVdq2) / norm(fdot));
    }
    return 0;
  }

  int advance(double dt, double sf) {  // advancer.ml:422-428
    bool any_first = false;
    for (auto &b : bs) any_first = any_first || (std::fpclassify(b.h) != FP_NORMAL);
    if (any_first) NBH_TRY(initialize_before_first_step(sf));
    std::stable_sort(bs.begin(), bs.end(), hmax_less);
    return advance_range((int)bs.size(), dt, sf);
  }

  void load(int n, const int *id, const double *state) {
    bs.resize(n);
    for (int i = 0; i < n; ++i) {
      const double *s = state + (size_t)i * NBH_ADV_STRIDE;
      ABody &b = bs[i];
      b.id = id[i];
      b.t = s[0];
      b.m = s[1];
      b.h = s[8];
      b.hmax = s[9];
      b.t0 = s[10];
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
    }
  }

  void store(int *id, double *state) const {
    for (size_t i = 0; i < bs.size(); ++i) {
      double *s = state + i * NBH_ADV_STRIDE;
      const ABody &b = bs[i];
      id[i] = b.id;
      s[0] = b.t;
      s[1] = b.m;
      s[8] = b.h;
      s[9] = b.hmax;
      s[10] = b.t0;
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
    }
  }

  int energy(double *e) {  // Energy.energy over the bodies' (m, q, p)
    const int n = (int)bs.size();
    std::vector<double> mm(n), qq(3 * (size_t)n), pp(3 * (size_t)n);
    for (int i = 0; i < n; ++i) {
      mm[i] = bs[i].m;
      for (int k = 0; k < 3; ++k) {
        qq[3 * i + k] = bs[i].q[k];
        pp[3 * i + k] = bs[i].p[k];
      }
    }
    double ke = 0.0, pe = 0.0;
    NBH_TRY(ck(ctx, nbk_energy(ctx, n, mm.data(), qq.data(), pp.data(), eps2, &ke, &pe)));
    *e = ke + pe;
    return 0;
  }
};

// The same control flow with the body records resident on the device (nbk_adv_*): per
// interact_bodies nothing crosses PCIe; per block step the host reads one time stamp and the
// finished rows of the active block (for h_adaptive, whose libm pow must stay on the host to
// keep the reference's time-step sequence), and sends a permutation when the sort moves rows.
// Bit-identical to the host-buffer path above (same kernels on the same packed rows, same
// operation sequences), which tests/ assert.
struct AdvResident {
  nbk_ctx *ctx;
  double eps2;
  long n_interact = 0, n_pairs = 0;
  int n = 0;
  int tree_max = 0;  // prefixes up to this many rows recurse on the device (0: host recursion)
  std::vector<double> hmax, recs;
  std::vector<int> id, perm;

  static void from_record(const double *s, ABody &b) {
    b.t = s[0];
    b.m = s[1];
    b.h = s[8];
    b.hmax = s[9];
    b.t0 = s[10];
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
  }

  int sort_prefix(int imax) {  // sort_range bs hmax_lt 0 imax (stable), rows follow on the device
    perm.resize(imax);
    for (int i = 0; i < imax; ++i) perm[i] = i;
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return hmax[a] < hmax[b]; });
    bool identity = true;
    for (int i = 0; i < imax; ++i) identity = identity && perm[i] == i;
    if (identity) return 0;
    NBH_TRY(ck(ctx, nbk_adv_permute(ctx, imax, perm.data())));
    std::vector<double> h2(imax);
    std::vector<int> i2(imax);
    for (int i = 0; i < imax; ++i) {
      h2[i] = hmax[perm[i]];
      i2[i] = id[perm[i]];
    }
    std::copy(h2.begin(), h2.end(), hmax.begin());
    std::copy(i2.begin(), i2.end(), id.begin());
    return 0;
  }

  int interact(int imin, int imax, double t, double vC) {
    NBH_TRY(ck(ctx, nbk_adv_interact(ctx, imin, imax, t, vC, eps2)));
    const long na = imax - imin, nn = n - imin;
    n_interact += 1;
    n_pairs += na * (nn - 1) - na * (na - 1) / 2;
    return 0;
  }

  // advance_range for a prefix that fits the device-side recursion (nbk_adv_advance_range): one
  // launch runs every block step of the subtree; hmax and ids follow the returned permutation.
  int advance_range_on_device(int imax, double dt, double sf) {
    perm.resize(imax);
    long long st[2] = {0, 0};
    NBH_TRY(ck(ctx, nbk_adv_advance_range(ctx, imax, dt, sf, eps2, hmax.data(), perm.data(), st)));
    std::vector<int> i2(imax);
    for (int i = 0; i < imax; ++i) i2[i] = id[perm[i]];
    std::copy(i2.begin(), i2.end(), id.begin());
    n_interact += st[0];
    n_pairs += st[1];
    return 0;
  }

  int advance_range(int imax, double dt, double sf) {  // advancer.ml:316-350
    if (tree_max > 0 && imax <= tree_max) return advance_range_on_device(imax, dt, sf);
    if (hmax[imax - 1] < dt) {
      NBH_TRY(advance_range(imax, dt / 2.0, sf));
      NBH_TRY(advance_range(imax, dt / 2.0, sf));
      return 0;
    }
    int imin = imax - 1;
    while (imin >= 0 && dt < hmax[imin]) --imin;
    imin = imin + 1;
    NBH_TRY(ck(ctx, nbk_adv_begin_step(ctx, imin, imax, dt)));
    double t0 = 0.0;
    NBH_TRY(ck(ctx, nbk_adv_get_t(ctx, imin, &t0)));  // let t0 = bs.(imin).t
    NBH_TRY(interact(imin, imax, t0, dt / 6.0));
    if (imin > 0) NBH_TRY(advance_range(imin, dt / 2.0, sf));
    NBH_TRY(interact(imin, imax, t0 + 0.5 * dt, 2.0 * dt / 3.0));
    if (imin > 0) NBH_TRY(advance_range(imin, dt / 2.0, sf));
    NBH_TRY(interact(imin, imax, t0 + dt, dt / 6.0));
    const int na = imax - imin;
    recs.resize((size_t)na * NBH_ADV_STRIDE);
    NBH_TRY(ck(ctx, nbk_adv_finish(ctx, imin, imax, t0 + dt, recs.data())));
    for (int k = 0; k < na; ++k) {
      ABody b;
      from_record(recs.data() + (size_t)k * NBH_ADV_STRIDE, b);
      Adv::h_adaptive(sf, b);
      hmax[imin + k] = b.hmax;
    }
    return sort_prefix(imax);
  }

  // energy_out (may be NULL): Energy.energy of the advanced bodies, computed from the resident
  // records before they are read back (nbk_adv_energy: nothing but two sums leaves the device)
  int run(int n_, int *id_io, double *state, double dt, double sf, double *energy_out = nullptr) {
    n = n_;
    id.assign(id_io, id_io + n);
    hmax.resize(n);
    bool any_first = false;
    for (int i = 0; i < n; ++i) {
      const double *s = state + (size_t)i * NBH_ADV_STRIDE;
      hmax[i] = s[9];
      any_first = any_first || (std::fpclassify(s[8]) != FP_NORMAL);
    }
    NBH_TRY(ck(ctx, nbk_adv_upload(ctx, n, state)));
    if (any_first) {  // initialize_before_first_step (advancer.ml:357-406)
      NBH_TRY(ck(ctx, nbk_adv_init_first(ctx, eps2)));
      recs.resize((size_t)n * NBH_ADV_STRIDE);
      NBH_TRY(ck(ctx, nbk_adv_download(ctx, 0, n, recs.data())));
      for (int i = 0; i < n; ++i) {
        ABody b;
        from_record(recs.data() + (size_t)i * NBH_ADV_STRIDE, b);
        double fdot[3];
        for (int j = 0; j < 3; ++j) fdot[j] = b.dVdq0[j] - b.dVdq1[j] + 3.0 * b.dVdq2[j];
        Adv::assign_timestep(b, 1e-2 * std::pow(sf, 0.33333) * b.h * norm(b.dVdq2) / norm(fdot));
        hmax[i] = b.hmax;
      }
    }
    NBH_TRY(sort_prefix(n));  // Array.fast_sort hmax_lt bs
    NBH_TRY(advance_range(n, dt, sf));
    if (energy_out) {
      double ke = 0.0, pe = 0.0;
      NBH_TRY(ck(ctx, nbk_adv_energy(ctx, eps2, &ke, &pe)));
      *energy_out = ke + pe;
    }
    recs.resize((size_t)n * NBH_ADV_STRIDE);
    NBH_TRY(ck(ctx, nbk_adv_download(ctx, 0, n, recs.data())));
    for (int i = 0; i < n; ++i) {
      double *s = state + (size_t)i * NBH_ADV_STRIDE;
      std::copy(recs.begin() + (size_t)i * NBH_ADV_STRIDE,
                recs.begin() + (size_t)(i + 1) * NBH_ADV_STRIDE, s);
      s[9] = hmax[i];
      id_io[i] = id[i];
    }
    return 0;
  }
};

// 0: host-buffer path (nbk_interact_sum per interact_bodies); 1: records resident, block recursion
// on the host (bit-identical to 0); 2: records resident and every subtree of the recursion whose
// prefix fits runs on the device (nbk_adv_advance_range; agrees with 1 to rounding)
int g_advancer_resident = 2;
int g_advancer_tree_max = 0;  // 0: automatic

int tree_max_for(int n) {
  if (g_advancer_resident < 2) return 0;
  const int cap = nbk_adv_range_max();
  if (g_advancer_tree_max > 0) return std::min(g_advancer_tree_max, cap);
  return n <= 8192 ? cap : cap / 2;  // large systems: big blocks are faster through the tiled sweeps
}

}  // namespace

extern "C" {

void nbh_set_advancer_resident(int mode) { g_advancer_resident = mode; }
void nbh_set_advancer_tree_max(int rows) { g_advancer_tree_max = rows; }

int nbh_advancer_advance(nbk_ctx *ctx, int n, int *id, double *state, double dt, double sf,
                         double eps2, nbh_extpot_fn extpot, void *extpot_arg, long *stats) {
  if (n <= 0) return fail(NBK_ERR_ARG, "Advancer.A.advance: empty body array");
  if (g_advancer_resident && !extpot) {  // an external field is a host callback: host path
    AdvResident R;
    R.ctx = ctx;
    R.eps2 = eps2;
    R.tree_max = tree_max_for(n);
    NBH_TRY(R.run(n, id, state, dt, sf));
    if (stats) {
      stats[0] = R.n_interact;
      stats[1] = R.n_pairs;
    }
    return 0;
  }
  Adv A;
  A.ctx = ctx;
  A.eps2 = eps2;
  A.extpot = extpot;
  A.extpot_arg = extpot_arg;
  A.load(n, id, state);  // Array.map body_copy (advancer.ml:423)
  NBH_TRY(A.advance(dt, sf));
  A.store(id, state);
  if (stats) {
    stats[0] = A.n_interact;
    stats[1] = A.n_pairs;
  }
  return 0;
}

void nbh_extpot_plummer_test(void *, double m, const double *q, const double *, double *gv) {
  double r = norm(q);
  double r2 = r * r + 0.01;
  double r3 = r2 * std::sqrt(r2);
  for (int k = 0; k < 3; ++k) gv[k] = m * q[k] / r3;
}

double nbh_extpot_plummer_test_energy(int n, const double *m, const double *q) {
  double e = 0.0;
  for (int i = 0; i < n; ++i) {
    double r = norm(q + 3 * i);
    double r2 = r * r + 0.01;
    e = e + (-m[i]) / std::sqrt(r2);
  }
  return e;
}

// integrator.ml:58-84 with filter = identity and max_dt = infinity (the defaults).
int nbh_constant_sf_integrate(nbk_ctx *ctx, int n, int *id, double *state, double dt, double ci,
                              double sf, double max_eg_err, double eps2, double *energy_errors,
                              int cap, int *nerr) {
  if (n <= 0) return fail(NBK_ERR_ARG, "constant_sf_integrate: empty body array");
  if (g_advancer_resident) {
    // Records resident on the device during each A.advance; the checkpoint energy is taken from
    // them before they come back (same body order, same kernels and summation trees as the
    // host-buffer path below, hence the same bits).
    Adv A0;
    A0.ctx = ctx;
    A0.eps2 = eps2;
    A0.extpot = nullptr;
    A0.extpot_arg = nullptr;
    A0.load(n, id, state);
    double e = 0.0;
    NBH_TRY(A0.energy(&e));
    const double stop_t = dt + state[0];
    std::vector<int> good_id(n);
    std::vector<double> good((size_t)n * NBH_ADV_STRIDE);
    int count = 0;
    for (;;) {
      std::copy(id, id + n, good_id.begin());
      std::copy(state, state + (size_t)n * NBH_ADV_STRIDE, good.begin());
      AdvResident R;
      R.ctx = ctx;
      R.eps2 = eps2;
      R.tree_max = tree_max_for(n);
      double new_e = 0.0;
      NBH_TRY(R.run(n, id, state, std::fmin(INFINITY, ci), sf, &new_e));
      const double new_t = state[0];
      const double eerr = std::fabs((new_e - e) / e);
      if (energy_errors && count < cap) energy_errors[count] = eerr;
      ++count;
      if (eerr > max_eg_err) {  // raise (Eg_err bs): hand back the last good state
        std::copy(good_id.begin(), good_id.end(), id);
        std::copy(good.begin(), good.end(), state);
        if (nerr) *nerr = count;
        return fail(NBH_ERR_EG, "Integrator.Eg_err: energy error bound exceeded");
      }
      if (new_t >= stop_t) break;
    }
    if (nerr) *nerr = count;
    return 0;
  }
  Adv A;
  A.ctx = ctx;
  A.eps2 = eps2;
  A.extpot = nullptr;
  A.extpot_arg = nullptr;
  A.load(n, id, state);
  double e = 0.0;
  NBH_TRY(A.energy(&e));
  const double stop_t = dt + A.bs[0].t;
  int count = 0;
  for (;;) {
    std::vector<ABody> good = A.bs;  // A.advance is functional: keep the last good state
    NBH_TRY(A.advance(std::fmin(INFINITY, ci), sf));
    double new_e = 0.0;
    NBH_TRY(A.energy(&new_e));
    const double new_t = A.bs[0].t;
    const double eerr = std::fabs((new_e - e) / e);
    if (energy_errors && count < cap) energy_errors[count] = eerr;
    ++count;
    if (eerr > max_eg_err) {  // raise (Eg_err bs)
      A.bs = good;
      A.store(id, state);
      if (nerr) *nerr = count;
      return fail(NBH_ERR_EG, "Integrator.Eg_err: energy error bound exceeded");
    }
    if (new_t >= stop_t) break;
  }
  A.store(id, state);
  if (nerr) *nerr = count;
  return 0;
}

}  // extern "C"
