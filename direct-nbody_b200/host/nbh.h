/* nbh.h — C entry points of libnbh, the host-side mirror of the reference's OCaml modules
 * (Hermite, Leapfrog, Omelyan_advancer.OA, Advancer.A, Energy, Integrator, Ics) written in
 * C++ above libnbk's C ABI.  The OCaml toolchain is absent from the build image, so this is
 * the stand-in for "the reference's integrators with their pair loops replaced by
 * nbk_* calls" (INTEGRATION.md shows the OCaml edit it mirrors).  Same names, argument
 * meaning and error behaviour as the OCaml functions; every O(N^2) loop goes through
 * include/nbk.h — there is no CPU pair loop in this library.
 *
 * Flat-array calling convention for the Python tests / bench: m[n], q[3n], p[3n] as in nbk.h.
 * Functions return 0 or a negative nbk error code; nbh_last_error() gives the text
 * (an OCaml caller would see Failure / Invalid_argument / Integrator.Eg_err).
 */
#ifndef NBH_H
#define NBH_H

#include <stdint.h>

#include "../../include/nbk.h"

#ifdef __cplusplus
extern "C" {
#endif

#define NBH_ERR_ASSERT (-10) /* an OCaml `assert` of the reference would have failed */
#define NBH_ERR_EG (-11)     /* Integrator.Eg_err: energy error bound exceeded */

const char *nbh_last_error(void);

/* ---- Ics (ics.ml) — seeded synthetic inputs, SURVEY.md §8d ------------------------------ */
/* make_plummer (ics.ml:211-227), make_hot_spherical (:229-239), make_cold_spherical
 * (:199-209), the unit-mass box of test/ic_test.ml:47-53; all end with adjust_frame. */
int nbh_make_plummer(int n, uint64_t seed, double *m, double *q, double *p);
int nbh_make_hot_spherical(int n, uint64_t seed, double *m, double *q, double *p);
int nbh_make_cold_spherical(int n, uint64_t seed, double *m, double *q, double *p);
int nbh_make_uniform_box(int n, uint64_t seed, double *m, double *q, double *p);
/* adjust_frame (ics.ml:161-175) */
int nbh_adjust_frame(int n, const double *m, double *q, double *p);
/* add_mass_spectrum with draw_mass_disc (ics.ml:333-343, bin/el_mass_segregation.ml:150-154) */
int nbh_add_mass_disc(int n, uint64_t seed, double heavyfrac, double heavyfac, double *m,
                      double *p);
/* rescale_to_standard_units (ics.ml:304-331); Energy.energy runs on the GPU via ctx. */
int nbh_rescale_to_standard_units(nbk_ctx *ctx, int n, double *m, double *q, double *p,
                                  double eps2);

/* ---- Energy.Make(B) (energy.ml:55-82) ------------------------------------------------------ */
/* total_kinetic_energy, total_potential_energy; energy = *ke + *pe (energy.ml:75-78). */
int nbh_energy(nbk_ctx *ctx, int n, const double *m, const double *q, const double *p,
               double eps2, double *ke, double *pe);
/* Analysis.bodies_to_el_phase_space (analysis.ml:904-919): el[4n] = {E, E/m, |L|, |L|/m}. */
int nbh_bodies_to_el_phase_space(nbk_ctx *ctx, int n, const double *m, const double *q,
                                 const double *p, double eps2, double *el);

/* ---- Omelyan_advancer.OA.advance h bs (omelyan_advancer.ml:96-104) -------------------------- */
int nbh_omelyan_advance(nbk_ctx *ctx, int n, const double *m, double *q, double *p, double *t,
                        double h, double eps2);

/* nsteps of OA.advance h with the bodies resident on the device; energies[nsteps] (may be NULL)
 * = Energy.energy after every step, el[4n] (may be NULL) = bodies_to_el_phase_space of the final
 * state.  Bit-identical to nsteps calls of nbh_omelyan_advance. */
int nbh_omelyan_advance_steps(nbk_ctx *ctx, int n, const double *m, double *q, double *p,
                              double *t, double h, double eps2, int nsteps, double *energies,
                              double *el);

/* ---- Leapfrog (leapfrog.ml:113-165): flat image of Leapfrog.b array ------------------------ */
/* advance bs dt eta: output order is the reference's (sorted by dt_max); id[] follows. */
int nbh_leapfrog_advance(nbk_ctx *ctx, int n, int *id, double *t, double *dt_max, double *m,
                         double *q, double *v, double dt, double eta, long *nkicks);
/* 2 (default): the bodies stay on the device (nbk_lf_*) and every subtree of advance_subsystem
 * whose prefix fits (nbk_lf_subsystem_max bodies; nbh_set_leapfrog_tree_max to lower)
 * runs in ONE launch (nbk_lf_advance_subsystem); advance_const_dt: once every body has
 * dt_max >= dt the remaining steps run on the device (nbk_resident_leapfrog_steps);
 * 1: resident bodies, block recursion on the host - bit-identical to 0;
 * 0: host-buffer path, every step through the host recursion. */
void nbh_set_leapfrog_resident(int mode);
void nbh_set_leapfrog_tree_max(int bodies); /* 0: automatic */
/* resident modes: sort_sub of at least this many rows runs on the device (nbk_lf_sort_sub);
 * default 512; a huge value keeps every sort on the host */
void nbh_set_leapfrog_device_sort_min(int rows);
int nbh_leapfrog_advance_const_dt(nbk_ctx *ctx, int n, int *id, double *t, double *dt_max,
                                  double *m, double *q, double *v, double dt, int nsteps);

/* ---- Hermite.advance bs dt sf (hermite.ml:180-189): flat image of Hermite.b array ----------- */
/* Output in id order (hermite.ml:136); *nsteps is bumped like Base.nsteps (hermite.ml:124). */
/* 0 (default): device-resident stepping loop when n <= nbk_hermite_resident_max(), else the
 * host scheduler with one kernel launch per advance_body; 1 / 2 force either. */
void nbh_set_hermite_mode(int mode);
/* Guard: Hermite.advance fails with NBK_ERR_STATE after this many steps (the reference loops
 * forever when its step-size formula collapses); default 2^40. */
void nbh_set_step_cap(long cap);
int nbh_hermite_advance(nbk_ctx *ctx, int n, int *id, double *t, double *m, double *q, double *p,
                        double *h, double *f, double *df, double dt, double sf, double eps2,
                        long *nsteps);

/* ---- DFloat_hermite.advance (dFloat_hermite.ml:199-209) and DFloat_pushforward.Make(DFloat_hermite)
 * (dFloat_pushforward.ml:39-76, 141-150) ----------------------------------------------------- */
/* Hermite over dual numbers carrying K tangents at once: every field of a body (t, q, p, h, f,
 * df) is a dual; masses, sf, dt, eps are constants.  Values as in nbh_hermite_advance; tangents
 * direction major: t_tan, h_tan [K][n], q_tan, p_tan, f_tan, df_tan [K][3n] (any may be NULL =
 * zero on input, not returned).  The scheduler compares values (the reference's comparisons on
 * diff); both pair loops run on the GPU (nbk_grad_v_dgrad_v_sum_tangent for start_bs,
 * nbk_grad_v_dgrad_v_sum_predicted_tangent for advance_body).  Output rows in id order. */
int nbh_dfloat_hermite_advance(nbk_ctx *ctx, int n, int K, const int *id, double *t, const double *m,
                               double *q, double *p, double *h, double *f, double *df,
                               double *t_tan, double *q_tan, double *p_tan, double *h_tan,
                               double *f_tan, double *df_tan, double dt, double sf, double eps2,
                               long *nsteps);
/* jacobian_advance dt sf bs: fresh Hermite bodies, state packed [q(3n); p(3n)], jac 6n x 6n
 * row-major (jac[r*6n + c] = d out[r] / d in[c]), out6n (may be NULL) the advanced state; one
 * pass over the 6n unit directions.  *nsteps (may be NULL) += advance_body calls. */
int nbh_dfloat_jacobian_advance(nbk_ctx *ctx, int n, const double *m, const double *q,
                                const double *p, double dt, double sf, double eps2, double *jac,
                                double *out6n, long *nsteps);
/* omega_pushforward's scalar from a Jacobian (dFloat_pushforward.ml:141-150): 1 if symplectic. */
double nbh_omega_pushforward(int n, const double *jac);
void nbh_set_dfloat_step_cap(long cap);

/* ---- Advancer.A.advance ?extpot bs dt sf (advancer.ml:422-428) ------------------------------ */
/* state[i*35..] = {t, m, q[3], p[3], h, hmax, t0, dVdq0[3], dVdq1[3], dVdq2[3], p0[3], q0[3],
 * q1[3], q2[3], cs[3]}: the Advancer.A.body record (advancer.ml:27-45); fresh bodies carry NaN
 * after p (A.make, advancer.ml:98-116).  extpot = gradient of an external potential
 * (`?extpot : b -> float array`), or NULL.  Output order: sorted by hmax, id[] follows.
 * stats (may be NULL) receives {interact_bodies calls, pair interactions evaluated}. */
#define NBH_ADV_STRIDE 35
typedef void (*nbh_extpot_fn)(void *arg, double m, const double *q, const double *p, double *gv);
/* 2 (default): the body records stay resident on the device for the whole advance (nbk_adv_*;
 * used when extpot == NULL) and every subtree of advance_range whose prefix fits
 * (nbk_adv_range_max rows; half of that above 8192 bodies, nbh_set_advancer_tree_max to force)
 * runs its whole block recursion on the device (nbk_adv_advance_range);
 * 1: resident records, block recursion on the host - bit-identical to 0;
 * 0: host-buffer path, one nbk_interact_sum per interact_bodies. */
void nbh_set_advancer_resident(int mode);
void nbh_set_advancer_tree_max(int rows); /* 0: automatic */
int nbh_advancer_advance(nbk_ctx *ctx, int n, int *id, double *state, double dt, double sf,
                         double eps2, nbh_extpot_fn extpot, void *extpot_arg, long *stats);
/* The field of test/advance_test.ml:39-46 (Plummer, M = 1, eps = 0.1) and its potential energy. */
void nbh_extpot_plummer_test(void *arg, double m, const double *q, const double *p, double *gv);
double nbh_extpot_plummer_test_energy(int n, const double *m, const double *q);

/* ---- Integrator.Make(B)(A).constant_sf_integrate (integrator.ml:58-84) over Advancer.A ------ */
/* Steps by min(ci, max_dt) with A.advance until t >= t0 + dt; after each step the relative
 * energy error against the initial energy is checked: > max_eg_err returns NBH_ERR_EG (the
 * reference raises Eg_err with the last good state: state/id are left at that state).
 * energy_errors (may be NULL, capacity cap) receives the error after every step; *nerr their
 * number. */
int nbh_constant_sf_integrate(nbk_ctx *ctx, int n, int *id, double *state, double dt, double ci,
                              double sf, double max_eg_err, double eps2, double *energy_errors,
                              int cap, int *nerr);

/* ---- text snapshots: A.print / A.read (advancer.ml:88-96, 118-124), Leapfrog.print / read
 * (leapfrog.ml:44-62) --------------------------------------------------------------------- */
/* The reference's `--- !Particle` records, byte for byte (Printf "%g" is C's): id, t, r, v
 * (= p/m for Advancer.A), m, and for Leapfrog t_max = t + dt_max.  append != 0 appends.  The
 * readers take every record of the file (the reference reads one per call): *n = number found,
 * an error if it exceeds cap or a record does not scan (Scanf.Scan_failure).  Read bodies carry
 * p = v * m (advancer.ml:124) and dt_max = t_max - t (leapfrog.ml:62). */
int nbh_advancer_print(const char *path, int append, int n, const int *id, const double *t,
                       const double *m, const double *q, const double *p);
int nbh_advancer_read(const char *path, int cap, int *n, int *id, double *t, double *m, double *q,
                      double *p);
int nbh_leapfrog_print(const char *path, int append, int n, const int *id, const double *t,
                       const double *dt_max, const double *m, const double *q, const double *v);
int nbh_leapfrog_read(const char *path, int cap, int *n, int *id, double *t, double *dt_max,
                      double *m, double *q, double *v);

/* ---- binary state streams: Marshal.to_channel chan (bs : B.b array) [] (bin/nbody.ml:80,
 * bin/el_mass_segregation.ml:96-108) and Analysis.load_states / load_state (analysis.ml:556-575) -- */
/* Array.length (load_states file): complete values in the stream (a truncated tail is ignored,
 * as the reference's `with _ ->` does). */
int nbh_marshal_count_states(const char *path, int *nstates);
/* load_state index file: *n bodies (<= cap), id[], state[n][35] in Advancer.A.body field order
 * (output_counter dropped).  *kind = 17 for Advancer.A.body, 6 for Leapfrog.b (row = {t, m, q[3],
 * v[3], dt_max}, rest NaN).  Physical sharing in the stream (CODE_SHARED) is resolved. */
int nbh_marshal_read_state(const char *path, int index, int cap, int *n, int *kind, int *id,
                           double *state);
/* Marshal.to_channel of an Advancer.A.body array: appends (append != 0) one value to the stream.
 * No sharing is emitted; doubles little-endian, small header - what ocamlopt writes on x86-64 /
 * arm64 up to sharing. */
int nbh_marshal_write_state(const char *path, int append, int n, const int *id, const double *state);

#ifdef __cplusplus
}
#endif
#endif
