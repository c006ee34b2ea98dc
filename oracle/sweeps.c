/* oracle/sweeps.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/base.h header).
 *
 * The reference's O(N^2) pair loops restated over flat arrays (mass m[n], position q[3n],
 * momentum p[3n] or velocity v[3n]) — the same packing the C-ABI boundary uses — in the
 * reference's own evaluation order:
 *
 *   oracle_start_bs_sweep          Hermite.start_bs pair loop      hermite.ml:154-167
 *   oracle_total_potential_energy  Energy.total_potential_energy   energy.ml:66-73
 *   oracle_el_potential            Analysis.bodies_to_el_phase_space pe loops  analysis.ml:909-914
 *   oracle_kick_all_pairs          Leapfrog.advance's dt=0 sweep   leapfrog.ml:151-155
 *   oracle_kick_block              Leapfrog.advance_subsystem kick loop  leapfrog.ml:132-136
 *
 * and the rectangular "sum over sources" forms the CUDA kernels are checked against
 * (targets [t0,t1) x sources [s0,s1), self excluded by index, ascending j):
 *
 *   oracle_grad_v_sum, oracle_grad_v_dgrad_v_sum, oracle_v_sum, oracle_kick_sum
 *
 * because Base.grad_v / grad_v_dgrad_v are antisymmetric bit-for-bit and the
 * symmetric loops visit partners in ascending index, a per-target ascending-j sum equals the
 * reference's per-body accumulation bit-for-bit (tests/test_oracle.py asserts it).
 */
#include "oracle.h"
#include "base.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Thread count of the OpenMP loops, set explicitly (bench.py's CPU arm must not inherit the
 * OMP_NUM_THREADS=1 that torchrun exports); n <= 0 leaves it alone. */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* hermite.ml:148-167: bodies get zeroed f, df; for each head b, for each later b2:
 *   f -= gv; f2 += gv; df -= dgv; df2 += dgv. */
void oracle_start_bs_sweep(int n, const double *m, const double *q, const double *p,
                           double eps2, double *f, double *df) {
  double gv[3], dgv[3];
  memset(f, 0, sizeof(double) * 3 * (size_t)n);
  memset(df, 0, sizeof(double) * 3 * (size_t)n);
  for (int i = 0; i < n; ++i) {
    for (int j = i + 1; j < n; ++j) {
      ob_grad_v_dgrad_v(eps2, m[i], q + 3 * i, p + 3 * i, m[j], q + 3 * j, p + 3 * j, gv, dgv);
      for (int k = 0; k < 3; ++k) {
        f[3 * i + k] = f[3 * i + k] - gv[k];
        f[3 * j + k] = f[3 * j + k] + gv[k];
        df[3 * i + k] = df[3 * i + k] - dgv[k];
        df[3 * j + k] = df[3 * j + k] + dgv[k];
      }
    }
  }
}

/* Same symmetric pair loop, all host threads: rows are dealt to threads dynamically, each
 * thread accumulates into private f/df, which are then added in thread order.  Used only as
 * the timed multi-core CPU arm (bench.py --impl reference); per-body sums differ from the
 * single-thread order by rounding. */
void oracle_start_bs_sweep_omp(int n, const double *m, const double *q, const double *p,
                               double eps2, double *f, double *df) {
  int nt = oracle_num_threads();
  size_t len = 3 * (size_t)n;
  double *pf = (double *)calloc((size_t)nt * len * 2, sizeof(double));
#pragma omp parallel
  {
#ifdef _OPENMP
    int tid = omp_get_thread_num();
#else
    int tid = 0;
#endif
    double *tf = pf + (size_t)tid * len * 2, *tdf = tf + len;
    double gv[3], dgv[3];
#pragma omp for schedule(dynamic, 16)
    for (int i = 0; i < n; ++i) {
      double fi[3] = {0, 0, 0}, dfi[3] = {0, 0, 0};
      for (int j = i + 1; j < n; ++j) {
        ob_grad_v_dgrad_v(eps2, m[i], q + 3 * i, p + 3 * i, m[j], q + 3 * j, p + 3 * j, gv, dgv);
        for (int k = 0; k < 3; ++k) {
          fi[k] = fi[k] - gv[k];
          tf[3 * j + k] = tf[3 * j + k] + gv[k];
          dfi[k] = dfi[k] - dgv[k];
          tdf[3 * j + k] = tdf[3 * j + k] + dgv[k];
        }
      }
      for (int k = 0; k < 3; ++k) {
        tf[3 * i + k] += fi[k];
        tdf[3 * i + k] += dfi[k];
      }
    }
  }
  memset(f, 0, sizeof(double) * len);
  memset(df, 0, sizeof(double) * len);
  for (int t = 0; t < nt; ++t) {
    const double *tf = pf + (size_t)t * len * 2, *tdf = tf + len;
    for (size_t a = 0; a < len; ++a) {
      f[a] += tf[a];
      df[a] += tdf[a];
    }
  }
  free(pf);
}

/* gsum_i = sum_{j in [s0,s1), j != i} Base.grad_v(i, j), i in [t0,t1); out[3*(i-t0)+k]. */
void oracle_grad_v_sum(int n, const double *m, const double *q, double eps2, int t0, int t1,
                       int s0, int s1, double *gsum) {
  (void)n;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = t0; i < t1; ++i) {
    double g[3] = {0, 0, 0}, gv[3];
    for (int j = s0; j < s1; ++j) {
      if (j == i) continue;
      ob_grad_v(eps2, m[i], q + 3 * i, m[j], q + 3 * j, gv);
      g[0] = g[0] + gv[0];
      g[1] = g[1] + gv[1];
      g[2] = g[2] + gv[2];
    }
    memcpy(gsum + 3 * (size_t)(i - t0), g, sizeof g);
  }
}

/* gsum_i, dgsum_i = sum_{j != i} Base.grad_v_dgrad_v(i, j). */
void oracle_grad_v_dgrad_v_sum(int n, const double *m, const double *q, const double *p,
                               double eps2, int t0, int t1, int s0, int s1, double *gsum,
                               double *dgsum) {
  (void)n;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = t0; i < t1; ++i) {
    double g[3] = {0, 0, 0}, dg[3] = {0, 0, 0}, gv[3], dgv[3];
    for (int j = s0; j < s1; ++j) {
      if (j == i) continue;
      ob_grad_v_dgrad_v(eps2, m[i], q + 3 * i, p + 3 * i, m[j], q + 3 * j, p + 3 * j, gv, dgv);
      for (int k = 0; k < 3; ++k) {
        g[k] = g[k] + gv[k];
        dg[k] = dg[k] + dgv[k];
      }
    }
    memcpy(gsum + 3 * (size_t)(i - t0), g, sizeof g);
    memcpy(dgsum + 3 * (size_t)(i - t0), dg, sizeof dg);
  }
}

/* energy.ml:66-73: one scalar accumulator over i<j, in loop order. */
double oracle_total_potential_energy(int n, const double *m, const double *q, double eps2) {
  double pe = 0.0;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) pe = pe + ob_v(eps2, m[i], q + 3 * i, m[j], q + 3 * j);
  return pe;
}

/* energy.ml:57-58,63-64: fold of |p|^2 / (2 m). */
double oracle_total_kinetic_energy(int n, const double *m, const double *p) {
  double e = 0.0;
  for (int i = 0; i < n; ++i) e = e + ob_norm_squared(p + 3 * i) / (2.0 * m[i]);
  return e;
}

/* energy.ml:75-78 */
double oracle_energy(int n, const double *m, const double *q, const double *p, double eps2) {
  double ke = oracle_total_kinetic_energy(n, m, p);
  double pe = oracle_total_potential_energy(n, m, q, eps2);
  return ke + pe;
}

/* phi_i = sum_{j in [s0,s1), j != i} Base.v(i, j): the pe accumulator of
 * analysis.ml:909-914 when [s0,s1) = [0,n). */
void oracle_v_sum(int n, const double *m, const double *q, double eps2, int t0, int t1, int s0,
                  int s1, double *phi) {
  (void)n;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = t0; i < t1; ++i) {
    double pe = 0.0;
    for (int j = s0; j < s1; ++j) {
      if (j == i) continue;
      pe = pe + ob_v(eps2, m[i], q + 3 * i, m[j], q + 3 * j);
    }
    phi[i - t0] = pe;
  }
}

/* analysis.ml:904-919: el[4*i..] = {E_i, E_i/m, |L_i|, |L_i|/m}. */
void oracle_bodies_to_el_phase_space(int n, const double *m, const double *q, const double *p,
                                     double eps2, double *el) {
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = 0; i < n; ++i) {
    const double *qi = q + 3 * i, *pi = p + 3 * i;
    double l[3] = {qi[1] * pi[2] - qi[2] * pi[1], qi[2] * pi[0] - qi[0] * pi[2],
                   qi[0] * pi[1] - qi[1] * pi[0]};
    double lb = ob_norm(l);
    double ke = ob_norm_squared(pi) / (2.0 * m[i]);
    double pe = 0.0;
    for (int j = 0; j < i; ++j) pe = pe + ob_v(eps2, m[i], qi, m[j], q + 3 * j);
    for (int j = i + 1; j < n; ++j) pe = pe + ob_v(eps2, m[i], qi, m[j], q + 3 * j);
    double etot = ke + pe;
    el[4 * i + 0] = etot;
    el[4 * i + 1] = etot / m[i];
    el[4 * i + 2] = lb;
    el[4 * i + 3] = lb / m[i];
  }
}

/* leapfrog.ml:151-155 (all pairs i<j) — literal sequential order, mutating v and dt_max. */
void oracle_kick_all_pairs(int n, const double *m, const double *q, double *v, double *dt_max,
                           double eta, double dt) {
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j)
      ob_kick(eta, dt, m[i], q + 3 * i, v + 3 * i, dt_max + i, m[j], q + 3 * j, v + 3 * j,
              dt_max + j);
}

/* leapfrog.ml:132-136: for i in [ibegin,iend) for j in [0,i): kick bs.(i) bs.(j). */
void oracle_kick_block(int n, const double *m, const double *q, double *v, double *dt_max,
                       double eta, double dt, int ibegin, int iend) {
  (void)n;
  for (int i = ibegin; i < iend; ++i)
    for (int j = 0; j < i; ++j)
      ob_kick(eta, dt, m[i], q + 3 * i, v + 3 * i, dt_max + i, m[j], q + 3 * j, v + 3 * j,
              dt_max + j);
}

/* "Jacobi order" form of the same block kick: every pair evaluated from the
 * velocities at sweep start.  For target i in [t0,t1), over sources j in [s0,s1), j != i:
 *   dv_i   = sum_j  (m_j * (dt / r^3)) * (q_j - q_i)       (leapfrog.ml:83-90, role of b1)
 *   tmin_i = min_j  tscale(i, j)   (NaN tscale ignored, start value +inf)  (leapfrog.ml:94-104)
 * The pair time-scale is symmetric bit-for-bit in (i,j) (dx,dv both flip sign), so one
 * function serves both roles. */
void oracle_kick_sum(int n, const double *m, const double *q, const double *v, double eta,
                     double dt, int t0, int t1, int s0, int s1, double *dv, double *tmin) {
  (void)n;
#pragma omp parallel for schedule(dynamic, 16)
  for (int i = t0; i < t1; ++i) {
    const double *q1 = q + 3 * i;
    double a[3] = {0, 0, 0}, tm = INFINITY;
    for (int j = s0; j < s1; ++j) {
      if (j == i) continue;
      const double *q2 = q + 3 * j;
      double dx = q2[0] - q1[0], dy = q2[1] - q1[1], dz = q2[2] - q1[2];
      double r2 = dx * dx + dy * dy + dz * dz;
      double r = sqrt(r2);
      double r3 = r2 * r;
      double c = dt / r3;
      double c1 = m[j] * c;
      a[0] = a[0] + c1 * dx;
      a[1] = a[1] + c1 * dy;
      a[2] = a[2] + c1 * dz;
      double ts = ob_kick_tscale(eta, m[i], q1, v + 3 * i, m[j], q2, v + 3 * j);
      if (tm > ts) tm = ts;
    }
    memcpy(dv + 3 * (size_t)(i - t0), a, sizeof a);
    tmin[i - t0] = tm;
  }
}

/* analysis.ml:809-822  binary_energy b1 b2 */
static inline double ob_binary_energy(double eps2, double m1, const double *q1, const double *p1,
                                      double m2, const double *q2, const double *p2) {
  double vrel[3];
  for (int i = 0; i < 3; ++i) vrel[i] = p2[i] / m2 - p1[i] / m1;
  double mu = m1 * m2 / (m1 + m2);
  double ke = 0.5 * mu * ob_dot(vrel, vrel);
  double pe = ob_v(eps2, m1, q1, m2, q2);
  return ke + pe;
}

/* analysis.ml:836-861  binaries (thr = 0) / binaries_threshold (thr = -|eg|): every pair i<j
 * with binary_energy < thr, reported in LOOP order (the reference conses, so its list is this
 * order reversed).  Returns the number found; stores at most max_pairs. */
int oracle_binaries(int n, const double *m, const double *q, const double *p, double eps2,
                    double thr, int max_pairs, int *pi, int *pj, double *e) {
  int k = 0;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      double en = ob_binary_energy(eps2, m[i], q + 3 * i, p + 3 * i, m[j], q + 3 * j, p + 3 * j);
      if (en < thr) {
        if (k < max_pairs) {
          pi[k] = i;
          pj[k] = j;
          e[k] = en;
        }
        ++k;
      }
    }
  return k;
}
