/* oracle/base.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE binary64, compiled with -ffp-contract=off so that no
 * multiply-add is fused, as OCaml/x86-64 SSE2 code does) of the reference's pair physics:
 *
 *   Base.v               /root/reference/base.ml:60-63
 *   Base.grad_v          /root/reference/base.ml:73-80
 *   Base.grad_v_dgrad_v  /root/reference/base.ml:91-110
 *   Base.norm_squared / norm / distance_squared / distance / dot   base.ml:36-58
 *   Leapfrog.kick        /root/reference/leapfrog.ml:70-104
 *
 * Every expression keeps the reference's association order (OCaml evaluates `a*.b*.c` as
 * `(a*.b)*.c` and `a*.b/.c` as `(a*.b)/.c`).  The softening is the caller's `eps2`
 * (the reference reads the global `!Base.eps2`, base.ml:23, at call time).
 *
 * PARITY STATUS: the reference repository holds no golden vector, fixed seed or fixture for
 * this path (test/run_tests.ml:10 self-seeds; the only fixture is listed in
 * .MISSING_LARGE_BLOBS) and no OCaml toolchain exists in the build image.  The restatement is
 * pinned against the reference's OWN SOURCE TEXT instead: oracle/miniml interprets the unmodified
 * .ml files under /root/reference, oracle/ref_fixtures/gen*.ml drive them over committed seeded
 * inputs, and this oracle reproduces all 8448 values they write BIT FOR BIT
 * (oracle/ref_fixtures/ref_miniml/, tests/test_ref_fixtures.py).  That is a pin through an
 * interpreter written for this repository (oracle/miniml/README.md says what it guarantees),
 * not through an ocamlopt binary: oracle/ref_fixtures/Makefile prepares that run for the first
 * machine with ocamlfind.  Also pinned by (i) the reference's own test invariants restated in
 * tests/test_oracle_invariants.py (test/ic_test.ml:9-23,47-61, test/advance_test.ml:9-52,
 * test/leapfrog_test.ml:7-25), (ii) analytic two-body values, (iii) a __float128 evaluation of
 * the same sums, (iv) an independent Python twin (oracle/twin.py).  Third-party Diff.DFloat
 * arithmetic stays unpinned (oracle/num.hh).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may use anything under oracle/.
 */
#ifndef NB_ORACLE_BASE_H
#define NB_ORACLE_BASE_H

#include <math.h>

/* base.ml:36-41 */
static inline double ob_norm_squared(const double v[3]) {
  double x = v[0], y = v[1], z = v[2];
  return x * x + y * y + z * z;
}
/* base.ml:43 */
static inline double ob_norm(const double v[3]) { return sqrt(ob_norm_squared(v)); }

/* base.ml:46-51 */
static inline double ob_distance_squared(const double x[3], const double y[3]) {
  double dx = x[0] - y[0], dy = x[1] - y[1], dz = x[2] - y[2];
  return dx * dx + dy * dy + dz * dz;
}
/* base.ml:53 */
static inline double ob_distance(const double x[3], const double y[3]) {
  return sqrt(ob_distance_squared(x, y));
}
/* base.ml:56-58 */
static inline double ob_dot(const double x[3], const double y[3]) {
  return x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
}

/* base.ml:60-63   v m1 q1 m2 q2 = ~-.m1*.m2/.r  == ((-m1)*m2)/r */
static inline double ob_v(double eps2, double m1, const double q1[3], double m2,
                          const double q2[3]) {
  double r2 = ob_distance_squared(q1, q2);
  double r = sqrt(r2 + eps2);
  return (-m1) * m2 / r;
}

/* base.ml:73-80 */
static inline void ob_grad_v(double eps2, double m1, const double q1[3], double m2,
                             const double q2[3], double gv[3]) {
  double r2 = ob_distance_squared(q1, q2);
  double re = r2 + eps2;
  double r3 = re * sqrt(re);
  double factor = m1 * m2 / r3;
  gv[0] = factor * (q1[0] - q2[0]);
  gv[1] = factor * (q1[1] - q2[1]);
  gv[2] = factor * (q1[2] - q2[2]);
}

/* base.ml:91-110 */
static inline void ob_grad_v_dgrad_v(double eps2, double m1, const double q1[3],
                                     const double p1[3], double m2, const double q2[3],
                                     const double p2[3], double gv[3], double dgv[3]) {
  gv[0] = q1[0] - q2[0];
  dgv[0] = p1[0] / m1 - p2[0] / m2;
  gv[1] = q1[1] - q2[1];
  dgv[1] = p1[1] / m1 - p2[1] / m2;
  gv[2] = q1[2] - q2[2];
  dgv[2] = p1[2] / m1 - p2[2] / m2;
  double r2 = ob_norm_squared(gv);
  double rdv = ob_dot(gv, dgv);
  double re = r2 + eps2;
  double r3 = sqrt(re) * re;
  double r5 = r3 * re;
  double factorg = m1 * m2 / r3;
  double factord = 3.0 * rdv * m1 * m2 / r5;
  dgv[0] = factorg * dgv[0] - factord * gv[0];
  gv[0] = factorg * gv[0];
  dgv[1] = factorg * dgv[1] - factord * gv[1];
  gv[1] = factorg * gv[1];
  dgv[2] = factorg * dgv[2] - factord * gv[2];
  gv[2] = factorg * gv[2];
}

/* leapfrog.ml:70-104.  Mutates v1, v2, *dtmax1, *dtmax2 exactly as the reference's record
 * fields.  `if a > b then a <- b` keeps a when b is NaN. */
static inline void ob_kick(double eta, double dt, double m1, const double q1[3], double v1[3],
                           double *dtmax1, double m2, const double q2[3], double v2[3],
                           double *dtmax2) {
  double dx = q2[0] - q1[0], dy = q2[1] - q1[1], dz = q2[2] - q1[2];
  double dvx = v2[0] - v1[0], dvy = v2[1] - v1[1], dvz = v2[2] - v1[2];
  double r2 = dx * dx + dy * dy + dz * dz;
  double vsq = dvx * dvx + dvy * dvy + dvz * dvz;
  double vdr = dx * dvx + dy * dvy + dz * dvz;
  double r = sqrt(r2);
  double r3 = r2 * r;
  double c = dt / r3;
  double c1 = m2 * c, c2 = m1 * c;
  v1[0] = v1[0] + c1 * dx;
  v1[1] = v1[1] + c1 * dy;
  v1[2] = v1[2] + c1 * dz;
  v2[0] = v2[0] - c2 * dx;
  v2[1] = v2[1] - c2 * dy;
  v2[2] = v2[2] - c2 * dz;
  double mtot = m1 + m2;
  double tff = eta * sqrt(r3 / mtot);
  double tc = eta * sqrt(r2 / vsq);
  double scaled_vdr = vdr / r2;
  double dtff = 1.5 * scaled_vdr * tff;
  double dtc = scaled_vdr * tc * (1.0 + mtot / (vsq * r));
  tff = fabs(tff / (1.0 - 0.5 * dtff));
  tc = fabs(tc / (1.0 - 0.5 * dtc));
  double tscale = (tff < tc) ? tff : tc;
  if (*dtmax1 > tscale) *dtmax1 = tscale;
  if (*dtmax2 > tscale) *dtmax2 = tscale;
}

/* The time-scale half of leapfrog.ml:94-104 alone, from given (pre-kick) velocities: the
 * "Jacobi order" variant SURVEY.md §3.3 defines kernel parity against. */
static inline double ob_kick_tscale(double eta, double m1, const double q1[3],
                                    const double v1[3], double m2, const double q2[3],
                                    const double v2[3]) {
  double dx = q2[0] - q1[0], dy = q2[1] - q1[1], dz = q2[2] - q1[2];
  double dvx = v2[0] - v1[0], dvy = v2[1] - v1[1], dvz = v2[2] - v1[2];
  double r2 = dx * dx + dy * dy + dz * dz;
  double vsq = dvx * dvx + dvy * dvy + dvz * dvz;
  double vdr = dx * dvx + dy * dvy + dz * dvz;
  double r = sqrt(r2);
  double r3 = r2 * r;
  double mtot = m1 + m2;
  double tff = eta * sqrt(r3 / mtot);
  double tc = eta * sqrt(r2 / vsq);
  double scaled_vdr = vdr / r2;
  double dtff = 1.5 * scaled_vdr * tff;
  double dtc = scaled_vdr * tc * (1.0 + mtot / (vsq * r));
  tff = fabs(tff / (1.0 - 0.5 * dtff));
  tc = fabs(tc / (1.0 - 0.5 * dtc));
  return (tff < tc) ? tff : tc;
}

#endif
