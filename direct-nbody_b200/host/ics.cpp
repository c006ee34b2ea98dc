// ics.cpp — host mirror of the reference's Ics recipes used by the BASELINE configs
// (ics.ml:161-239, 304-343; bin/el_mass_segregation.ml:150-154), seeded as SURVEY.md §8d
// fixes: splitmix64, u = (x >> 11) * 2^-53, one draw per Random.float call, draws in the
// textual order of each recipe.  O(N) host work; the one O(N^2) piece
// (rescale_to_standard_units' Energy.energy, ics.ml:317) runs on the GPU through nbk_energy.
#include "nbh.h"

#include <cmath>
#include <string>

namespace nbh {
thread_local std::string g_err;
int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
}  // namespace nbh

namespace {

struct Rng {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  }
  double uniform(double x) { return (double)(next() >> 11) * 0x1.0p-53 * x; }  // Random.float x
  double between(double a, double b) {  // ics.ml:177-179
    double dx = b - a;
    return a + uniform(dx);
  }
  template <class F> double from_dist(double xmin, double xmax, double ymin, double ymax, F dist) {
    // ics.ml:181-187
    double x = between(xmin, xmax), y = between(ymin, ymax);
    while (!(y < dist(x))) {
      x = between(xmin, xmax);
      y = between(ymin, ymax);
    }
    return x;
  }
  void vector(double len, double out[3]) {  // ics.ml:189-197
    const double pi = 4.0 * std::atan(1.0);
    double cos_theta = between(-1.0, 1.0);
    double phi = between(0.0, 2.0 * pi);
    double theta = std::acos(cos_theta);
    double sin_theta = std::sin(theta);
    out[0] = len * std::cos(phi) * sin_theta;
    out[1] = len * std::sin(phi) * sin_theta;
    out[2] = len * cos_theta;
  }
};

inline double sq(double x) { return x * x; }

}  // namespace

extern "C" {

const char *nbh_last_error(void) { return nbh::g_err.c_str(); }

int nbh_adjust_frame(int n, const double *m, double *q, double *p) {
  // Analysis.total_momentum / total_mass / center_of_mass (analysis.ml:508-537)
  double p_tot[3] = {0, 0, 0}, com[3] = {0, 0, 0}, m_tot = 0.0;
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) p_tot[k] = p_tot[k] + p[3 * i + k];
  for (int i = 0; i < n; ++i) m_tot = m_tot + m[i];
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) com[k] = com[k] + q[3 * i + k] * (m[i] / m_tot);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < 3; ++k) {
      q[3 * i + k] = q[3 * i + k] - com[k];
      p[3 * i + k] = p[3 * i + k] - p_tot[k] * m[i] / m_tot;
    }
  return 0;
}

int nbh_make_plummer(int n, uint64_t seed, double *m, double *q, double *p) {
  Rng rng{seed};
  const double mass = 1.0 / (double)n, pi = 4.0 * std::atan(1.0);
  const double sf = 16.0 / (3.0 * pi);
  for (int i = 0; i < n; ++i) {
    double r = 1.0 / std::sqrt(std::pow(rng.between(0.0, 1.0), -2.0 / 3.0) - 1.0);
    double x = rng.from_dist(0.0, 1.0, 0.0, 0.1,
                             [](double x) { return sq(x) * std::pow(1.0 - sq(x), 3.5); });
    rng.vector(r / sf, q + 3 * i);
    rng.vector(std::sqrt(sf) * mass * x * std::sqrt(2.0) * std::pow(1.0 + sq(r), -0.25),
               p + 3 * i);
    m[i] = mass;
  }
  return nbh_adjust_frame(n, m, q, p);
}

int nbh_make_hot_spherical(int n, uint64_t seed, double *m, double *q, double *p) {
  Rng rng{seed};
  const double v0 = std::sqrt(21.0 / 10.0), mass = 1.0 / (double)n;
  for (int i = 0; i < n; ++i) {
    double r = rng.from_dist(0.0, 1.0, 0.0, 1.0, sq);
    double v = rng.between(0.0, v0);
    rng.vector(r, q + 3 * i);
    rng.vector(mass * v, p + 3 * i);
    m[i] = mass;
  }
  return nbh_adjust_frame(n, m, q, p);
}

int nbh_make_cold_spherical(int n, uint64_t seed, double *m, double *q, double *p) {
  Rng rng{seed};
  const double mass = 1.0 / (double)n;
  for (int i = 0; i < n; ++i) {
    double r = rng.from_dist(0.0, 12.0 / 5.0, 0.0, 144.0 / 25.0, sq);
    rng.vector(r, q + 3 * i);
    p[3 * i] = p[3 * i + 1] = p[3 * i + 2] = 0.0;
    m[i] = mass;
  }
  return nbh_adjust_frame(n, m, q, p);
}

int nbh_make_uniform_box(int n, uint64_t seed, double *m, double *q, double *p) {
  Rng rng{seed};
  for (int i = 0; i < n; ++i) {
    m[i] = 1.0;
    for (int k = 0; k < 3; ++k) q[3 * i + k] = rng.uniform(1.0);
    for (int k = 0; k < 3; ++k) p[3 * i + k] = rng.uniform(0.01);
  }
  return nbh_adjust_frame(n, m, q, p);
}

int nbh_add_mass_disc(int n, uint64_t seed, double heavyfrac, double heavyfac, double *m,
                      double *p) {
  Rng rng{seed};
  for (int i = 0; i < n; ++i) {
    double mnew = (rng.uniform(1.0) < heavyfrac) ? heavyfac : 1.0;
    for (int k = 0; k < 3; ++k) p[3 * i + k] = p[3 * i + k] * mnew / m[i];  // rescale_mass
    m[i] = mnew;
  }
  return 0;
}

int nbh_rescale_to_standard_units(nbk_ctx *ctx, int n, double *m, double *q, double *p,
                                  double eps2) {
  double mtot = 0.0;
  for (int i = 0; i < n; ++i) mtot = mtot + m[i];
  for (int i = 0; i < n; ++i) {
    double mnew = m[i] / mtot;
    for (int k = 0; k < 3; ++k) p[3 * i + k] = p[3 * i + k] * mnew / m[i];
    m[i] = mnew;
  }
  double ke = 0.0, pe = 0.0;
  int rc = nbk_energy(ctx, n, m, q, p, eps2, &ke, &pe);
  if (rc != 0) return nbh::fail(rc, nbk_last_error(ctx));
  double e = ke + pe;
  if (!(e < 0.0)) return nbh::fail(NBH_ERR_ASSERT, "rescale_to_standard_units: assert(e < 0.0)");
  double efactor = (-0.25) / e;
  for (int i = 0; i < 3 * n; ++i) {
    q[i] = q[i] / efactor;
    p[i] = p[i] * std::sqrt(efactor);
  }
  return 0;
}

}  // extern "C"
