// energy_omelyan.cpp — host mirror of Energy.Make(B) (energy.ml:55-82),
// Analysis.bodies_to_el_phase_space (analysis.ml:904-919) and Omelyan_advancer.OA
// (omelyan_advancer.ml:39-104) with every pair loop routed through libnbk.
#include <cmath>
#include <vector>

#include "common.hpp"

namespace nbh {
double norm(const double *v) { return std::sqrt(norm_squared(v)); }
double distance(const double *x, const double *y) {
  double dx = x[0] - y[0], dy = x[1] - y[1], dz = x[2] - y[2];
  return std::sqrt(dx * dx + dy * dy + dz * dz);
}
}  // namespace nbh

using namespace nbh;

extern "C" {

int nbh_energy(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p,
               double eps2, double *ke, double *pe) {
  return ck(ctx, nbk_energy(ctx, n, m, q, p, eps2, ke, pe));
}

int nbh_bodies_to_el_phase_space(nbk_ctx *ctx, int n, const double *m, const double *q,
                                 const double *p, double eps2, double *el) {
  std::vector<double> phi(n);
  NBH_TRY(ck(ctx, nbk_v_sum(ctx, n, m, q, eps2, 0, n, 0, n, phi.data(), nullptr)));
  for (int i = 0; i < n; ++i) {
    const double *qi = q + 3 * i, *pi = p + 3 * i;
    double l[3] = {qi[1] * pi[2] - qi[2] * pi[1], qi[2] * pi[0] - qi[0] * pi[2],
                   qi[0] * pi[1] - qi[1] * pi[0]};
    double lb = norm(l);
    double ke = norm_squared(pi) / (2.0 * m[i]);
    double etot = ke + phi[i];
    el[4 * i + 0] = etot;
    el[4 * i + 1] = etot / m[i];
    el[4 * i + 2] = lb;
    el[4 * i + 3] = lb / m[i];
  }
  return 0;
}

// OA.advance: B(h) A(h) C(h) A(h) B(h); grad_v is antisymmetric bit for bit, so the pairwise
// p_i -= f gv, p_j += f gv of b_operate (:44-56) is p_i -= f gsum_i, and likewise for c_operate.
int nbh_omelyan_advance(nbk_ctx *ctx, int n, const double *m, double *q, double *p, double *t,
                        double h, double eps2) {
  std::vector<double> g(3 * (size_t)n), qs(3 * (size_t)n);
  auto b_operate = [&]() -> int {
    const double f = h / 6.0;
    NBH_TRY(ck(ctx, nbk_grad_v_sum(ctx, n, m, q, eps2, 0, n, 0, n, g.data())));
    for (size_t a = 0; a < g.size(); ++a) p[a] = p[a] - f * g[a];
    return 0;
  };
  auto a_operate = [&]() {
    for (int i = 0; i < n; ++i) {
      const double f = h / (2.0 * m[i]);
      for (int k = 0; k < 3; ++k) q[3 * i + k] = q[3 * i + k] + p[3 * i + k] * f;
    }
  };
  auto c_operate = [&]() -> int {
    const double pf = 2.0 * h / 3.0;
    NBH_TRY(ck(ctx, nbk_grad_v_sum(ctx, n, m, q, eps2, 0, n, 0, n, g.data())));
    for (int i = 0; i < n; ++i) {
      const double f = square(h) / (24.0 * m[i]);
      for (int k = 0; k < 3; ++k) qs[3 * i + k] = q[3 * i + k] - f * g[3 * i + k];
    }
    NBH_TRY(ck(ctx, nbk_grad_v_sum(ctx, n, m, qs.data(), eps2, 0, n, 0, n, g.data())));
    for (size_t a = 0; a < g.size(); ++a) p[a] = p[a] - pf * g[a];
    return 0;
  };
  NBH_TRY(b_operate());
  a_operate();
  NBH_TRY(c_operate());
  a_operate();
  NBH_TRY(b_operate());
  *t = *t + h;
  return 0;
}

// nsteps of OA.advance h with the bodies resident on the device (nbk_oa_*): one upload, all the
// B A C A B operators and their four sweeps on the GPU, one download.  energies (may be NULL)
// receives Energy.energy after every step - the per-step diagnostics of
// bin/el_mass_segregation.ml's loop - and el (may be NULL) bodies_to_el_phase_space of the final
// state, both from one potential sweep per step.  A multi-device context steps through the
// host-buffer calls instead (its sweeps shard over the devices).
int nbh_omelyan_advance_steps(nbk_ctx *ctx, int n, const double *m, double *q, double *p,
                              double *t, double h, double eps2, int nsteps, double *energies,
                              double *el) {
  if (nsteps < 0) return fail(NBK_ERR_ARG, "nbh_omelyan_advance_steps: nsteps < 0");
  if (nbk_device_count(ctx) > 1 || n == 0) {
    for (int s = 0; s < nsteps; ++s) {
      NBH_TRY(nbh_omelyan_advance(ctx, n, m, q, p, t, h, eps2));
      if (energies) {
        double ke = 0.0, pe = 0.0;
        NBH_TRY(nbh_energy(ctx, n, m, q, p, eps2, &ke, &pe));
        energies[s] = ke + pe;
      }
    }
    if (el) NBH_TRY(nbh_bodies_to_el_phase_space(ctx, n, m, q, p, eps2, el));
    return 0;
  }
  NBH_TRY(ck(ctx, nbk_oa_upload(ctx, n, m, q, p)));
  for (int s = 0; s < nsteps; ++s) {
    NBH_TRY(ck(ctx, nbk_oa_steps(ctx, h, eps2, 1)));
    *t = *t + h;
    if (energies || (el && s == nsteps - 1)) {
      double ke = 0.0, pe = 0.0;
      NBH_TRY(ck(ctx, nbk_oa_diagnostics(ctx, eps2, &ke, &pe, s == nsteps - 1 ? el : nullptr)));
      if (energies) energies[s] = ke + pe;
    }
  }
  if (el && nsteps == 0) NBH_TRY(ck(ctx, nbk_oa_diagnostics(ctx, eps2, nullptr, nullptr, el)));
  return ck(ctx, nbk_oa_download(ctx, q, p));
}

}  // extern "C"
