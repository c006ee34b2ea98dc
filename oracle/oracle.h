/* oracle/oracle.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C entry points of the CPU restatement of farr/direct-nbody's pairwise-gravity hot path
 * (see oracle/base.h for the arithmetic, its reference citations and the parity status).
 * Built by oracle/Makefile into oracle/_build/liboracle.so; loaded by oracle/oracle.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may use it.  All arrays are caller-owned, contiguous, double unless noted:
 * mass m[n], position q[3n], momentum p[3n] (velocity v[3n] for Leapfrog).
 */
#ifndef NB_ORACLE_H
#define NB_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int oracle_num_threads(void);
void oracle_set_num_threads(int n);
/* test-run guard: step loops stop early past this many steps (default 2e8) */
void oracle_set_step_cap(long cap);

/* sweeps.c */
void oracle_start_bs_sweep(int n, const double *m, const double *q, const double *p,
                           double eps2, double *f, double *df);
void oracle_start_bs_sweep_omp(int n, const double *m, const double *q, const double *p,
                               double eps2, double *f, double *df);
void oracle_grad_v_sum(int n, const double *m, const double *q, double eps2, int t0, int t1,
                       int s0, int s1, double *gsum);
void oracle_grad_v_dgrad_v_sum(int n, const double *m, const double *q, const double *p,
                               double eps2, int t0, int t1, int s0, int s1, double *gsum,
                               double *dgsum);
double oracle_total_potential_energy(int n, const double *m, const double *q, double eps2);
double oracle_total_kinetic_energy(int n, const double *m, const double *p);
double oracle_energy(int n, const double *m, const double *q, const double *p, double eps2);
void oracle_v_sum(int n, const double *m, const double *q, double eps2, int t0, int t1, int s0,
                  int s1, double *phi);
void oracle_bodies_to_el_phase_space(int n, const double *m, const double *q, const double *p,
                                     double eps2, double *el);
void oracle_kick_all_pairs(int n, const double *m, const double *q, double *v, double *dt_max,
                           double eta, double dt);
void oracle_kick_block(int n, const double *m, const double *q, double *v, double *dt_max,
                       double eta, double dt, int ibegin, int iend);
void oracle_kick_sum(int n, const double *m, const double *q, const double *v, double eta,
                     double dt, int t0, int t1, int s0, int s1, double *dv, double *tmin);

int oracle_binaries(int n, const double *m, const double *q, const double *p, double eps2,
                    double thr, int max_pairs, int *pi, int *pj, double *e);

/* quad.c (binary128 truth) */
void oracle_grad_v_dgrad_v_sum_quad(int n, const double *m, const double *q, const double *p,
                                    double eps2, int t0, int t1, int s0, int s1, double *gsum,
                                    double *dgsum);
void oracle_v_sum_quad(int n, const double *m, const double *q, double eps2, int t0, int t1,
                       int s0, int s1, double *phi);
double oracle_total_potential_energy_quad(int n, const double *m, const double *q, double eps2);

/* ics.c */
void oracle_adjust_frame(int n, const double *m, double *q, double *p);
void oracle_make_plummer(int n, uint64_t seed, double *m, double *q, double *p);
void oracle_make_hot_spherical(int n, uint64_t seed, double *m, double *q, double *p);
void oracle_make_cold_spherical(int n, uint64_t seed, double *m, double *q, double *p);
void oracle_make_uniform_box(int n, uint64_t seed, double *m, double *q, double *p);
void oracle_add_mass_disc(int n, uint64_t seed, double heavyfrac, double heavyfac, double *m,
                          double *p);
int oracle_rescale_to_standard_units(int n, double *m, double *q, double *p, double eps2);

/* hermite.cc */
void oracle_hermite_advance(int n, int *id, double *t, double *m, double *q, double *p,
                            double *h, double *f, double *df, double dt, double sf,
                            double eps2, long *nsteps);
void oracle_hermite_advance_body(int n, const double *t, const double *m, const double *q,
                                 const double *p, const double *h, const double *f,
                                 const double *df, int ib, double eta, double eps2,
                                 double *out14);
void oracle_dfloat_hermite_advance_body(int n, const double *t, const double *m, const double *q,
                                        const double *p, const double *h, const double *f,
                                        const double *df, const double *t_d, const double *q_d,
                                        const double *p_d, const double *h_d, const double *f_d,
                                        const double *df_d, int ib, double eta, double eps2,
                                        double *out14, double *out14_d);
void oracle_grad_v_dgrad_v_sum_tangent(int n, const double *m, const double *q, const double *p,
                                       const double *dq, const double *dp, double eps2, int t0,
                                       int t1, int s0, int s1, double *gsum, double *dgsum,
                                       double *gsum_tan, double *dgsum_tan);
void oracle_grad_v_sum_tangent(int n, const double *m, const double *q, const double *dq,
                               double eps2, int t0, int t1, int s0, int s1, double *gsum,
                               double *gsum_tan);
void oracle_hermite_jacobian_advance(int n, const double *m, const double *q, const double *p,
                                     double dt, double sf, double eps2, double *jac,
                                     double *out6n);
double oracle_omega_pushforward(int n, const double *jac);

/* leapfrog.cc */
void oracle_leapfrog_advance(int n, int *id, double *t, double *dt_max, double *m, double *q,
                             double *v, double dt, double eta, long *nkicks);
void oracle_leapfrog_advance_const_dt(int n, int *id, double *t, double *dt_max, double *m,
                                      double *q, double *v, double dt, int nsteps);
void oracle_omelyan_advance(int n, const double *m, double *q, double *p, double *t, double h,
                            double eps2);

/* advancer.cc */
#define ORACLE_ADV_STRIDE 35
typedef void (*oracle_extpot_fn)(void *arg, double m, const double *q, const double *p,
                                 double *gv);
void oracle_advancer_advance(int n, int *id, double *state, double dt, double sf, double eps2,
                             oracle_extpot_fn extpot, void *extpot_arg, long *stats);
void oracle_extpot_plummer_test(void *arg, double m, const double *q, const double *p,
                                double *gv);
double oracle_extpot_plummer_test_energy(int n, const double *m, const double *q);

#ifdef __cplusplus
}
#endif
#endif
