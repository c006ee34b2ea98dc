/* oracle/ics.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/base.h header).
 *
 * Seeded restatement of the initial-condition recipes the BASELINE configs are built from:
 *
 *   random_between / random_from_dist / random_vector   ics.ml:177-197
 *   make_cold_spherical   ics.ml:199-209        make_plummer        ics.ml:211-227
 *   make_hot_spherical    ics.ml:229-239        adjust_frame        ics.ml:161-175
 *   rescale_to_standard_units   ics.ml:304-331  rescale_mass / add_mass_spectrum  ics.ml:333-343
 *   draw_mass_disc        bin/el_mass_segregation.ml:150-154
 *   Analysis.total_momentum / total_mass / center_of_mass    analysis.ml:508-537
 *
 * The reference draws from OCaml's self-seeded Random (test/run_tests.ml:10,
 * bin/el_mass_segregation.ml:68-72), whose stream cannot be reproduced; SURVEY.md §8d
 * fixes instead: splitmix64, u = (x >> 11) * 2^-53 in [0,1), one draw per Random.float
 * call, draws taken in the textual order of each recipe.  The product-side generator
 * (direct-nbody_b200/host/ics.cpp) follows the same convention and tests check both agree
 * bit for bit.
 */
#include "oracle.h"
#include "base.h"

#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t s;
} rng_t;

static inline uint64_t rng_next(rng_t *r) {
  uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
/* Random.float x */
static inline double rng_float(rng_t *r, double x) {
  return (double)(rng_next(r) >> 11) * 0x1.0p-53 * x;
}
/* ics.ml:177-179 */
static inline double random_between(rng_t *r, double a, double b) {
  double dx = b - a;
  return a + rng_float(r, dx);
}

typedef double (*dist_fn)(double);

/* ics.ml:181-187 (x drawn before y) */
static double random_from_dist(rng_t *r, double xmin, double xmax, double ymin, double ymax,
                               dist_fn dist) {
  double x = random_between(r, xmin, xmax);
  double y = random_between(r, ymin, ymax);
  while (!(y < dist(x))) {
    x = random_between(r, xmin, xmax);
    y = random_between(r, ymin, ymax);
  }
  return x;
}

/* ics.ml:189-197 */
static void random_vector(rng_t *r, double len, double out[3]) {
  double pi = 4.0 * atan(1.0);
  double cos_theta = random_between(r, -1.0, 1.0);
  double phi = random_between(r, 0.0, 2.0 * pi);
  double theta = acos(cos_theta);
  double sin_theta = sin(theta);
  out[0] = len * cos(phi) * sin_theta;
  out[1] = len * sin(phi) * sin_theta;
  out[2] = len * cos_theta;
}

static double square(double x) { return x * x; }
static double plummer_g(double x) { return square(x) * pow(1.0 - square(x), 3.5); }

/* analysis.ml:508-537 + ics.ml:161-175 */
void oracle_adjust_frame(int n, const double *m, double *q, double *p) {
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
}

/* ics.ml:211-227 */
void oracle_make_plummer(int n, uint64_t seed, double *m, double *q, double *p) {
  rng_t rng = {seed};
  double mass = 1.0 / (double)n, pi = 4.0 * atan(1.0);
  double sf = 16.0 / (3.0 * pi);
  for (int i = 0; i < n; ++i) {
    double r = 1.0 / sqrt(pow(random_between(&rng, 0.0, 1.0), -2.0 / 3.0) - 1.0);
    double x = random_from_dist(&rng, 0.0, 1.0, 0.0, 0.1, plummer_g);
    random_vector(&rng, r / sf, q + 3 * i);
    random_vector(&rng, sqrt(sf) * mass * x * sqrt(2.0) * pow(1.0 + square(r), -0.25),
                  p + 3 * i);
    m[i] = mass;
  }
  oracle_adjust_frame(n, m, q, p);
}

/* ics.ml:229-239 */
void oracle_make_hot_spherical(int n, uint64_t seed, double *m, double *q, double *p) {
  rng_t rng = {seed};
  double v0 = sqrt(21.0 / 10.0), mass = 1.0 / (double)n;
  for (int i = 0; i < n; ++i) {
    double r = random_from_dist(&rng, 0.0, 1.0, 0.0, 1.0, square);
    double v = random_between(&rng, 0.0, v0);
    random_vector(&rng, r, q + 3 * i);
    random_vector(&rng, mass * v, p + 3 * i);
    m[i] = mass;
  }
  oracle_adjust_frame(n, m, q, p);
}

/* ics.ml:199-209 */
void oracle_make_cold_spherical(int n, uint64_t seed, double *m, double *q, double *p) {
  rng_t rng = {seed};
  double mass = 1.0 / (double)n;
  for (int i = 0; i < n; ++i) {
    double r = random_from_dist(&rng, 0.0, 12.0 / 5.0, 0.0, 144.0 / 25.0, square);
    random_vector(&rng, r, q + 3 * i);
    p[3 * i] = p[3 * i + 1] = p[3 * i + 2] = 0.0;
    m[i] = mass;
  }
  oracle_adjust_frame(n, m, q, p);
}

/* test/ic_test.ml:47-53: unit masses, q ~ U[0,1)^3, p ~ U[0,0.01)^3, then adjust_frame. */
void oracle_make_uniform_box(int n, uint64_t seed, double *m, double *q, double *p) {
  rng_t rng = {seed};
  for (int i = 0; i < n; ++i) {
    m[i] = 1.0;
    for (int k = 0; k < 3; ++k) q[3 * i + k] = rng_float(&rng, 1.0);
    for (int k = 0; k < 3; ++k) p[3 * i + k] = rng_float(&rng, 0.01);
  }
  oracle_adjust_frame(n, m, q, p);
}

/* ics.ml:333-343 with bin/el_mass_segregation.ml:150-154 as the mass generator. */
void oracle_add_mass_disc(int n, uint64_t seed, double heavyfrac, double heavyfac, double *m,
                          double *p) {
  rng_t rng = {seed};
  for (int i = 0; i < n; ++i) {
    double mnew = (rng_float(&rng, 1.0) < heavyfrac) ? heavyfac : 1.0;
    for (int k = 0; k < 3; ++k) p[3 * i + k] = p[3 * i + k] * mnew / m[i];
    m[i] = mnew;
  }
}

/* ics.ml:304-331; the energy is Energy.energy with the caller's eps2 (global in the
 * reference).  Returns 0, or -1 where the reference's assert(e < 0.0) would fail. */
int oracle_rescale_to_standard_units(int n, double *m, double *q, double *p, double eps2) {
  double mtot = 0.0;
  for (int i = 0; i < n; ++i) mtot = mtot + m[i];
  for (int i = 0; i < n; ++i) {
    double mnew = m[i] / mtot;
    for (int k = 0; k < 3; ++k) p[3 * i + k] = p[3 * i + k] * mnew / m[i];
    m[i] = mnew;
  }
  double e = oracle_energy(n, m, q, p, eps2);
  if (!(e < 0.0)) return -1;
  double efactor = (-0.25) / e;
  for (int i = 0; i < 3 * n; ++i) {
    q[i] = q[i] / efactor;
    p[i] = p[i] * sqrt(efactor);
  }
  return 0;
}
