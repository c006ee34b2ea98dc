/* oracle/quad.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/base.h header).
 *
 * "Truth" evaluation of the same per-target sums in IEEE binary128 (__float128, libquadmath),
 * formulas of base.ml:60-110 and leapfrog.ml:83-90.  Used to show that both the reference
 * summation order and the GPU's summation order sit inside the 1e-12 parity tolerance
 * (BASELINE.md §4).
 */
#include "oracle.h"

#include <quadmath.h>

typedef __float128 q128;

void oracle_grad_v_dgrad_v_sum_quad(int n, const double *m, const double *q, const double *p,
                                    double eps2, int t0, int t1, int s0, int s1, double *gsum,
                                    double *dgsum) {
  (void)n;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = t0; i < t1; ++i) {
    q128 g[3] = {0, 0, 0}, dg[3] = {0, 0, 0};
    q128 mi = m[i];
    for (int j = s0; j < s1; ++j) {
      if (j == i) continue;
      q128 mj = m[j];
      q128 r[3], w[3];
      for (int k = 0; k < 3; ++k) {
        r[k] = (q128)q[3 * i + k] - (q128)q[3 * j + k];
        w[k] = (q128)p[3 * i + k] / mi - (q128)p[3 * j + k] / mj;
      }
      q128 r2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      q128 rdv = r[0] * w[0] + r[1] * w[1] + r[2] * w[2];
      q128 re = r2 + (q128)eps2;
      q128 r3 = sqrtq(re) * re;
      q128 r5 = r3 * re;
      q128 fg = mi * mj / r3;
      q128 fd = 3 * rdv * mi * mj / r5;
      for (int k = 0; k < 3; ++k) {
        g[k] += fg * r[k];
        dg[k] += fg * w[k] - fd * r[k];
      }
    }
    for (int k = 0; k < 3; ++k) {
      gsum[3 * (size_t)(i - t0) + k] = (double)g[k];
      if (dgsum) dgsum[3 * (size_t)(i - t0) + k] = (double)dg[k];
    }
  }
}

void oracle_v_sum_quad(int n, const double *m, const double *q, double eps2, int t0, int t1,
                       int s0, int s1, double *phi) {
  (void)n;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = t0; i < t1; ++i) {
    q128 pe = 0;
    for (int j = s0; j < s1; ++j) {
      if (j == i) continue;
      q128 r2 = 0;
      for (int k = 0; k < 3; ++k) {
        q128 d = (q128)q[3 * i + k] - (q128)q[3 * j + k];
        r2 += d * d;
      }
      pe += -(q128)m[i] * (q128)m[j] / sqrtq(r2 + (q128)eps2);
    }
    phi[i - t0] = (double)pe;
  }
}

double oracle_total_potential_energy_quad(int n, const double *m, const double *q, double eps2) {
  q128 pe = 0;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      q128 r2 = 0;
      for (int k = 0; k < 3; ++k) {
        q128 d = (q128)q[3 * i + k] - (q128)q[3 * j + k];
        r2 += d * d;
      }
      pe += -(q128)m[i] * (q128)m[j] / sqrtq(r2 + (q128)eps2);
    }
  return (double)pe;
}
