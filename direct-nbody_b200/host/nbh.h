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

#ifdef __cplusplus
}
#endif
#endif
