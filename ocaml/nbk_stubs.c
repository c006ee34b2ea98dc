/* nbk_stubs.c — OCaml <-> libnbk glue: the C side of the `external` declarations in nbk.ml.
 *
 * NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE: the image has no OCaml toolchain (no ocaml,
 * ocamlfind, ocamlbuild, opam or dune), so <caml/...> headers are absent.  These sources are
 * what a maintainer of farr/direct-nbody adds to the reference tree; they are mechanical
 * (unbox arguments, call include/nbk.h, raise Failure on a non-zero code) and every call
 * they make is exercised here through the same C ABI from Python/C++ (tests/).
 *
 * All arrays are Bigarray.Array1 of float64, c_layout (bigarray is already linked for every
 * target of the reference, _tags:5-11), so Caml_ba_data_val gives the double* libnbk wants
 * without copying.  libnbk never calls back into OCaml and never keeps a pointer past the
 * call, so no root registration is needed; the calls are long (an O(N^2) sweep), so the
 * runtime lock could be released around them, but the reference is single-threaded
 *  and nothing is gained.
 */
#include <caml/alloc.h>
#include <caml/bigarray.h>
#include <caml/custom.h>
#include <caml/fail.h>
#include <caml/memory.h>
#include <caml/mlvalues.h>

#include "nbk.h"

#define Ctx_val(v) (*((nbk_ctx **)Data_custom_val(v)))
#define D(v) ((double *)Caml_ba_data_val(v))

static void ctx_finalize(value v) {
  if (Ctx_val(v)) nbk_destroy(Ctx_val(v));
  Ctx_val(v) = NULL;
}

static struct custom_operations ctx_ops = {"org.nbk.ctx",       ctx_finalize,
                                           custom_compare_default, custom_hash_default,
                                           custom_serialize_default, custom_deserialize_default,
                                           custom_compare_ext_default, custom_fixed_length_default};

static void check(nbk_ctx *ctx, int rc) {
  if (rc != NBK_OK) caml_failwith(nbk_last_error(ctx)); /* reference convention: exceptions */
}

CAMLprim value nbk_ml_create(value device) {
  CAMLparam1(device);
  CAMLlocal1(v);
  nbk_ctx *ctx = NULL;
  int rc = nbk_create(&ctx, Int_val(device));
  if (rc != NBK_OK) caml_failwith(nbk_last_error(NULL));
  v = caml_alloc_custom(&ctx_ops, sizeof(nbk_ctx *), 0, 1);
  Ctx_val(v) = ctx;
  CAMLreturn(v);
}

CAMLprim value nbk_ml_create_multi(value ndev) {
  CAMLparam1(ndev);
  CAMLlocal1(v);
  nbk_ctx *ctx = NULL;
  int rc = nbk_create_multi(&ctx, Int_val(ndev), NULL);
  if (rc != NBK_OK) caml_failwith(nbk_last_error(NULL));
  v = caml_alloc_custom(&ctx_ops, sizeof(nbk_ctx *), 0, 1);
  Ctx_val(v) = ctx;
  CAMLreturn(v);
}

/* external grad_v_sum : ctx -> m -> q -> eps2:float -> range(t0,t1,s0,s1) -> gsum -> unit */
CAMLprim value nbk_ml_grad_v_sum(value ctx, value m, value q, value eps2, value rng,
                                 value gsum) {
  CAMLparam5(ctx, m, q, eps2, rng);
  CAMLxparam1(gsum);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx),
        nbk_grad_v_sum(Ctx_val(ctx), n, D(m), D(q), Double_val(eps2), Int_val(Field(rng, 0)),
                       Int_val(Field(rng, 1)), Int_val(Field(rng, 2)), Int_val(Field(rng, 3)),
                       D(gsum)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_grad_v_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_grad_v_sum(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* external grad_v_dgrad_v_sum : ctx -> m -> q -> p -> eps2 -> range -> gsum -> dgsum -> unit */
CAMLprim value nbk_ml_grad_v_dgrad_v_sum(value ctx, value m, value q, value p, value eps2,
                                         value rng, value gsum, value dgsum) {
  CAMLparam5(ctx, m, q, p, eps2);
  CAMLxparam3(rng, gsum, dgsum);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx),
        nbk_grad_v_dgrad_v_sum(Ctx_val(ctx), n, D(m), D(q), D(p), Double_val(eps2),
                               Int_val(Field(rng, 0)), Int_val(Field(rng, 1)),
                               Int_val(Field(rng, 2)), Int_val(Field(rng, 3)), D(gsum),
                               D(dgsum)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_grad_v_dgrad_v_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_grad_v_dgrad_v_sum(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
}

/* external grad_v_dgrad_v_sum_predicted :
     ctx -> t -> m -> q -> p -> f -> df -> nt:float -> eps2 -> range -> gsum -> dgsum -> unit */
CAMLprim value nbk_ml_predicted_bc(value *a, int n) {
  (void)n;
  int nb = (int)Caml_ba_array_val(a[2])->dim[0];
  value rng = a[9];
  check(Ctx_val(a[0]),
        nbk_grad_v_dgrad_v_sum_predicted(Ctx_val(a[0]), nb, D(a[1]), D(a[2]), D(a[3]), D(a[4]),
                                         D(a[5]), D(a[6]), Double_val(a[7]), Double_val(a[8]),
                                         Int_val(Field(rng, 0)), Int_val(Field(rng, 1)),
                                         Int_val(Field(rng, 2)), Int_val(Field(rng, 3)), D(a[10]),
                                         D(a[11])));
  return Val_unit;
}
CAMLprim value nbk_ml_predicted(value a0, value a1, value a2, value a3, value a4, value a5,
                                value a6, value a7, value a8, value a9, value a10, value a11) {
  value a[12] = {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11};
  return nbk_ml_predicted_bc(a, 12);
}

/* external energy : ctx -> m -> q -> p -> eps2 -> float * float   (ke, pe) */
CAMLprim value nbk_ml_energy(value ctx, value m, value q, value p, value eps2) {
  CAMLparam5(ctx, m, q, p, eps2);
  CAMLlocal1(res);
  double ke = 0.0, pe = 0.0;
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx), nbk_energy(Ctx_val(ctx), n, D(m), D(q), D(p), Double_val(eps2), &ke, &pe));
  res = caml_alloc_tuple(2);
  Store_field(res, 0, caml_copy_double(ke));
  Store_field(res, 1, caml_copy_double(pe));
  CAMLreturn(res);
}

/* external v_sum : ctx -> m -> q -> eps2 -> range -> phi -> float (sum of phi) */
CAMLprim value nbk_ml_v_sum(value ctx, value m, value q, value eps2, value rng, value phi) {
  CAMLparam5(ctx, m, q, eps2, rng);
  CAMLxparam1(phi);
  double total = 0.0;
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx),
        nbk_v_sum(Ctx_val(ctx), n, D(m), D(q), Double_val(eps2), Int_val(Field(rng, 0)),
                  Int_val(Field(rng, 1)), Int_val(Field(rng, 2)), Int_val(Field(rng, 3)), D(phi),
                  &total));
  CAMLreturn(caml_copy_double(total));
}
CAMLprim value nbk_ml_v_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_v_sum(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* external kick_sum : ctx -> m -> q -> v -> eta -> dt -> range -> dv -> tmin -> unit */
CAMLprim value nbk_ml_kick_sum_bc(value *a, int n) {
  (void)n;
  int nb = (int)Caml_ba_array_val(a[1])->dim[0];
  value rng = a[6];
  check(Ctx_val(a[0]),
        nbk_kick_sum(Ctx_val(a[0]), nb, D(a[1]), D(a[2]), D(a[3]), Double_val(a[4]),
                     Double_val(a[5]), Int_val(Field(rng, 0)), Int_val(Field(rng, 1)),
                     Int_val(Field(rng, 2)), Int_val(Field(rng, 3)), D(a[7]), D(a[8])));
  return Val_unit;
}
CAMLprim value nbk_ml_kick_sum(value a0, value a1, value a2, value a3, value a4, value a5,
                               value a6, value a7, value a8) {
  value a[9] = {a0, a1, a2, a3, a4, a5, a6, a7, a8};
  return nbk_ml_kick_sum_bc(a, 9);
}

/* external grad_v_dgrad_v_sum_tangent :
     ctx -> m -> q -> p -> k:int -> dq -> dp -> eps2 -> range -> gsum_tan -> dgsum_tan -> unit */
CAMLprim value nbk_ml_tangent_bc(value *a, int n) {
  (void)n;
  int nb = (int)Caml_ba_array_val(a[1])->dim[0];
  value rng = a[8];
  check(Ctx_val(a[0]),
        nbk_grad_v_dgrad_v_sum_tangent(Ctx_val(a[0]), nb, D(a[1]), D(a[2]), D(a[3]), Int_val(a[4]),
                                       D(a[5]), D(a[6]), Double_val(a[7]), Int_val(Field(rng, 0)),
                                       Int_val(Field(rng, 1)), Int_val(Field(rng, 2)),
                                       Int_val(Field(rng, 3)), NULL, NULL, D(a[9]), D(a[10])));
  return Val_unit;
}
CAMLprim value nbk_ml_tangent(value a0, value a1, value a2, value a3, value a4, value a5,
                              value a6, value a7, value a8, value a9, value a10) {
  value a[11] = {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10};
  return nbk_ml_tangent_bc(a, 11);
}

/* ---- bodies resident on the device for Omelyan_advancer.OA (nbk_oa_*) ---------------------- */
/* external oa_upload : ctx -> m -> q -> p -> unit */
CAMLprim value nbk_ml_oa_upload(value ctx, value m, value q, value p) {
  CAMLparam4(ctx, m, q, p);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx), nbk_oa_upload(Ctx_val(ctx), n, D(m), D(q), D(p)));
  CAMLreturn(Val_unit);
}

/* external oa_steps : ctx -> h:float -> eps2:float -> nsteps:int -> unit */
CAMLprim value nbk_ml_oa_steps(value ctx, value h, value eps2, value nsteps) {
  CAMLparam4(ctx, h, eps2, nsteps);
  check(Ctx_val(ctx), nbk_oa_steps(Ctx_val(ctx), Double_val(h), Double_val(eps2), Int_val(nsteps)));
  CAMLreturn(Val_unit);
}

/* external oa_download : ctx -> q -> p -> unit */
CAMLprim value nbk_ml_oa_download(value ctx, value q, value p) {
  CAMLparam3(ctx, q, p);
  check(Ctx_val(ctx), nbk_oa_download(Ctx_val(ctx), D(q), D(p)));
  CAMLreturn(Val_unit);
}

/* external oa_diagnostics : ctx -> eps2:float -> el -> float * float   (ke, pe) */
CAMLprim value nbk_ml_oa_diagnostics(value ctx, value eps2, value el) {
  CAMLparam3(ctx, eps2, el);
  CAMLlocal1(res);
  double ke = 0.0, pe = 0.0;
  check(Ctx_val(ctx), nbk_oa_diagnostics(Ctx_val(ctx), Double_val(eps2), &ke, &pe, D(el)));
  res = caml_alloc_tuple(2);
  Store_field(res, 0, caml_copy_double(ke));
  Store_field(res, 1, caml_copy_double(pe));
  CAMLreturn(res);
}

/* ---- Advancer.A (advancer.ml:207-428) --------------------------------------------------------- */
/* external interact_sum : ctx -> m -> q -> na:int -> eps2:float -> s -> unit
 * bodies [imin, n) re-indexed from 0, the first na of them active (advancer.ml:213-235) */
CAMLprim value nbk_ml_interact_sum(value ctx, value m, value q, value na, value eps2, value s) {
  CAMLparam5(ctx, m, q, na, eps2);
  CAMLxparam1(s);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx), nbk_interact_sum(Ctx_val(ctx), n, Int_val(na), D(m), D(q), Double_val(eps2), D(s)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_interact_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_interact_sum(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* Records resident on the device: records = Bigarray of n*35 doubles in the field order of
 * include/nbk.h (t, m, q, p, h, hmax, t0, dVdq0-2, p0, q0, q1, q2, cs). */
/* external adv_upload : ctx -> records -> unit */
CAMLprim value nbk_ml_adv_upload(value ctx, value rec) {
  CAMLparam2(ctx, rec);
  int n = (int)(Caml_ba_array_val(rec)->dim[0] / 35);
  check(Ctx_val(ctx), nbk_adv_upload(Ctx_val(ctx), n, D(rec)));
  CAMLreturn(Val_unit);
}
/* external adv_download : ctx -> i0:int -> i1:int -> records -> unit */
CAMLprim value nbk_ml_adv_download(value ctx, value i0, value i1, value rec) {
  CAMLparam4(ctx, i0, i1, rec);
  check(Ctx_val(ctx), nbk_adv_download(Ctx_val(ctx), Int_val(i0), Int_val(i1), D(rec)));
  CAMLreturn(Val_unit);
}
/* external adv_permute : ctx -> (int32, c_layout) Array1 -> unit */
CAMLprim value nbk_ml_adv_permute(value ctx, value perm) {
  CAMLparam2(ctx, perm);
  int n = (int)Caml_ba_array_val(perm)->dim[0];
  check(Ctx_val(ctx), nbk_adv_permute(Ctx_val(ctx), n, (const int *)Caml_ba_data_val(perm)));
  CAMLreturn(Val_unit);
}
/* external adv_begin_step : ctx -> imin:int -> imax:int -> hnew:float -> unit */
CAMLprim value nbk_ml_adv_begin_step(value ctx, value imin, value imax, value hnew) {
  CAMLparam4(ctx, imin, imax, hnew);
  check(Ctx_val(ctx), nbk_adv_begin_step(Ctx_val(ctx), Int_val(imin), Int_val(imax), Double_val(hnew)));
  CAMLreturn(Val_unit);
}
/* external adv_interact : ctx -> imin:int -> imax:int -> t:float -> vC:float -> eps2:float -> unit */
CAMLprim value nbk_ml_adv_interact(value ctx, value imin, value imax, value t, value vC, value eps2) {
  CAMLparam5(ctx, imin, imax, t, vC);
  CAMLxparam1(eps2);
  check(Ctx_val(ctx), nbk_adv_interact(Ctx_val(ctx), Int_val(imin), Int_val(imax), Double_val(t),
                                       Double_val(vC), Double_val(eps2)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_adv_interact_bc(value *a, int n) {
  (void)n;
  return nbk_ml_adv_interact(a[0], a[1], a[2], a[3], a[4], a[5]);
}
/* external adv_finish : ctx -> imin:int -> imax:int -> t:float -> records -> unit
 * (records receives the (imax-imin)*35 finished doubles) */
CAMLprim value nbk_ml_adv_finish(value ctx, value imin, value imax, value t, value rec) {
  CAMLparam5(ctx, imin, imax, t, rec);
  check(Ctx_val(ctx), nbk_adv_finish(Ctx_val(ctx), Int_val(imin), Int_val(imax), Double_val(t), D(rec)));
  CAMLreturn(Val_unit);
}
/* external adv_get_t : ctx -> int -> float */
CAMLprim value nbk_ml_adv_get_t(value ctx, value i) {
  CAMLparam2(ctx, i);
  double t = 0.0;
  check(Ctx_val(ctx), nbk_adv_get_t(Ctx_val(ctx), Int_val(i), &t));
  CAMLreturn(caml_copy_double(t));
}
/* external adv_init_first : ctx -> eps2:float -> unit */
CAMLprim value nbk_ml_adv_init_first(value ctx, value eps2) {
  CAMLparam2(ctx, eps2);
  check(Ctx_val(ctx), nbk_adv_init_first(Ctx_val(ctx), Double_val(eps2)));
  CAMLreturn(Val_unit);
}
/* external adv_range_max : unit -> int */
CAMLprim value nbk_ml_adv_range_max(value unit) {
  (void)unit;
  return Val_int(nbk_adv_range_max());
}
/* external adv_advance_range : ctx -> dt:float -> sf:float -> eps2:float -> hmax -> perm -> unit
 * advance_range bs imax dt sf None on the device, imax = dim hmax = dim perm (int32) */
CAMLprim value nbk_ml_adv_advance_range(value ctx, value dt, value sf, value eps2, value hmax,
                                        value perm) {
  CAMLparam5(ctx, dt, sf, eps2, hmax);
  CAMLxparam1(perm);
  int imax = (int)Caml_ba_array_val(hmax)->dim[0];
  check(Ctx_val(ctx), nbk_adv_advance_range(Ctx_val(ctx), imax, Double_val(dt), Double_val(sf),
                                            Double_val(eps2), D(hmax), (int *)Caml_ba_data_val(perm),
                                            NULL));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_adv_advance_range_bc(value *a, int n) {
  (void)n;
  return nbk_ml_adv_advance_range(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* ---- Leapfrog with the bodies resident on the device (leapfrog.ml:106-158) --------------------- */
/* external lf_upload : ctx -> dt_max -> m -> q -> v -> unit */
CAMLprim value nbk_ml_lf_upload(value ctx, value dtm, value m, value q, value v) {
  CAMLparam5(ctx, dtm, m, q, v);
  int n = (int)Caml_ba_array_val(m)->dim[0];
  check(Ctx_val(ctx), nbk_lf_upload(Ctx_val(ctx), n, D(dtm), D(m), D(q), D(v)));
  CAMLreturn(Val_unit);
}
/* external lf_download : ctx -> dt_max -> q -> v -> unit */
CAMLprim value nbk_ml_lf_download(value ctx, value dtm, value q, value v) {
  CAMLparam4(ctx, dtm, q, v);
  check(Ctx_val(ctx), nbk_lf_download(Ctx_val(ctx), D(dtm), D(q), D(v)));
  CAMLreturn(Val_unit);
}
/* external lf_advance_subsystem : ctx -> dt:float -> eta:float -> t -> dt_max -> perm -> unit
 * advance_subsystem dt eta bs iend on the device, iend = dim t = dim dt_max = dim perm (int32) */
CAMLprim value nbk_ml_lf_advance_subsystem(value ctx, value dt, value eta, value t, value dtm,
                                           value perm) {
  CAMLparam5(ctx, dt, eta, t, dtm);
  CAMLxparam1(perm);
  int iend = (int)Caml_ba_array_val(t)->dim[0];
  check(Ctx_val(ctx), nbk_lf_advance_subsystem(Ctx_val(ctx), iend, Double_val(dt), Double_val(eta),
                                               D(t), D(dtm), (int *)Caml_ba_data_val(perm), NULL));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_lf_advance_subsystem_bc(value *a, int n) {
  (void)n;
  return nbk_ml_lf_advance_subsystem(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* ================================================================================================
 * Round 2: the integrator-facing entry points that had no binding yet (missing #4).
 * Same conventions as above; functions of more than five arguments come as a native stub plus a
 * bytecode stub taking argv (OCaml's rule for externals with > 5 parameters).
 * ============================================================================================== */
#define R0(rng) Int_val(Field(rng, 0))
#define R1(rng) Int_val(Field(rng, 1))
#define R2(rng) Int_val(Field(rng, 2))
#define R3(rng) Int_val(Field(rng, 3))
#define I32(v) ((int *)Caml_ba_data_val(v))
#define DIM0(v) ((int)Caml_ba_array_val(v)->dim[0])

static value pair_of_doubles(double a, double b) {
  CAMLparam0();
  CAMLlocal1(res);
  res = caml_alloc_tuple(2);
  Store_field(res, 0, caml_copy_double(a));
  Store_field(res, 1, caml_copy_double(b));
  CAMLreturn(res);
}

/* external device_count : ctx -> int */
CAMLprim value nbk_ml_device_count(value ctx) { return Val_int(nbk_device_count(Ctx_val(ctx))); }

/* external kinetic_energy : ctx -> m -> p -> float        (Energy.total_kinetic_energy only:
 * pe = NULL, so NO potential sweep runs) */
CAMLprim value nbk_ml_kinetic_energy(value ctx, value m, value p) {
  CAMLparam3(ctx, m, p);
  double ke = 0.0;
  check(Ctx_val(ctx), nbk_energy(Ctx_val(ctx), DIM0(m), D(m), NULL, D(p), 0.0, &ke, NULL));
  CAMLreturn(caml_copy_double(ke));
}
/* external potential_energy : ctx -> m -> q -> eps2:float -> float   (ke = NULL) */
CAMLprim value nbk_ml_potential_energy(value ctx, value m, value q, value eps2) {
  CAMLparam4(ctx, m, q, eps2);
  double pe = 0.0;
  check(Ctx_val(ctx), nbk_energy(Ctx_val(ctx), DIM0(m), D(m), D(q), NULL, Double_val(eps2), NULL, &pe));
  CAMLreturn(caml_copy_double(pe));
}

/* external kick_dv_sum : ctx -> m -> q -> dt:float -> range -> dv -> unit
 * (the eta = infinity path of Leapfrog.advance_const_dt: tmin = NULL, no velocities needed) */
CAMLprim value nbk_ml_kick_dv_sum(value ctx, value m, value q, value dt, value rng, value dv) {
  CAMLparam5(ctx, m, q, dt, rng);
  CAMLxparam1(dv);
  check(Ctx_val(ctx), nbk_kick_sum(Ctx_val(ctx), DIM0(m), D(m), D(q), NULL, 0.0, Double_val(dt),
                                   R0(rng), R1(rng), R2(rng), R3(rng), D(dv), NULL));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_kick_dv_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_kick_dv_sum(a[0], a[1], a[2], a[3], a[4], a[5]);
}

/* external kick_block_sum : ctx -> m -> q -> v -> ibegin:int -> eta:float -> dt:float -> dv -> tmin -> unit
 * the whole kick loop of advance_subsystem (leapfrog.ml:132-136) for the bodies [0, iend = dim m) */
CAMLprim value nbk_ml_kick_block_sum(value ctx, value m, value q, value v, value ibegin, value eta,
                                     value dt, value dv, value tmin) {
  CAMLparam5(ctx, m, q, v, ibegin);
  CAMLxparam4(eta, dt, dv, tmin);
  check(Ctx_val(ctx), nbk_kick_block_sum(Ctx_val(ctx), DIM0(m), Int_val(ibegin), D(m), D(q), D(v),
                                         Double_val(eta), Double_val(dt), D(dv), D(tmin)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_kick_block_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_kick_block_sum(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]);
}

/* external grad_v_sum_tangent : ctx -> m -> q -> k:int -> dq -> eps2:float -> range -> gsum_tan -> unit */
CAMLprim value nbk_ml_grad_v_sum_tangent(value ctx, value m, value q, value k, value dq, value eps2,
                                         value rng, value gt) {
  CAMLparam5(ctx, m, q, k, dq);
  CAMLxparam3(eps2, rng, gt);
  check(Ctx_val(ctx), nbk_grad_v_sum_tangent(Ctx_val(ctx), DIM0(m), D(m), D(q), Int_val(k), D(dq),
                                             Double_val(eps2), R0(rng), R1(rng), R2(rng), R3(rng),
                                             NULL, D(gt)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_grad_v_sum_tangent_bc(value *a, int n) {
  (void)n;
  return nbk_ml_grad_v_sum_tangent(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
}

/* external grad_v_dgrad_v_sum_predicted_tangent :
 *   ctx -> state:(t, m, q, p, f, df) -> k:int -> tans:(t_tan, q_tan, p_tan, f_tan, df_tan) ->
 *   nt:float -> nt_tan -> eps2:float -> range -> out:(gsum, dgsum, gsum_tan, dgsum_tan) -> unit
 * DFloat_hermite.advance_body's loop (dFloat_hermite.ml:115-135); the array groups travel as
 * OCaml tuples to keep the external at nine arguments. */
CAMLprim value nbk_ml_predicted_tangent(value ctx, value st, value k, value tans, value nt,
                                        value nt_tan, value eps2, value rng, value out) {
  CAMLparam5(ctx, st, k, tans, nt);
  CAMLxparam4(nt_tan, eps2, rng, out);
  check(Ctx_val(ctx),
        nbk_grad_v_dgrad_v_sum_predicted_tangent(
            Ctx_val(ctx), DIM0(Field(st, 1)), D(Field(st, 0)), D(Field(st, 1)), D(Field(st, 2)),
            D(Field(st, 3)), D(Field(st, 4)), D(Field(st, 5)), Int_val(k), D(Field(tans, 0)),
            D(Field(tans, 1)), D(Field(tans, 2)), D(Field(tans, 3)), D(Field(tans, 4)),
            Double_val(nt), D(nt_tan), Double_val(eps2), R0(rng), R1(rng), R2(rng), R3(rng),
            D(Field(out, 0)), D(Field(out, 1)), D(Field(out, 2)), D(Field(out, 3))));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_predicted_tangent_bc(value *a, int n) {
  (void)n;
  return nbk_ml_predicted_tangent(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]);
}

/* ---- Analysis.binaries / binaries_threshold / tightest_binary (analysis.ml:836-876) ---------- */
/* external binary_scan : ctx -> m -> q -> p -> eps2:float -> threshold:float -> range ->
 *                        emin -> partner:(int32) -> count:(int32) -> unit */
CAMLprim value nbk_ml_binary_scan(value ctx, value m, value q, value p, value eps2, value thr,
                                  value rng, value emin, value partner, value count) {
  CAMLparam5(ctx, m, q, p, eps2);
  CAMLxparam5(thr, rng, emin, partner, count);
  check(Ctx_val(ctx), nbk_binary_scan(Ctx_val(ctx), DIM0(m), D(m), D(q), D(p), Double_val(eps2),
                                      Double_val(thr), R0(rng), R1(rng), R2(rng), R3(rng), D(emin),
                                      I32(partner), I32(count)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_binary_scan_bc(value *a, int n) {
  (void)n;
  return nbk_ml_binary_scan(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]);
}
/* external binary_list : ctx -> m -> q -> p -> eps2:float -> threshold:float ->
 *                        pi:(int32) -> pj:(int32) -> e -> int      (number of pairs found) */
CAMLprim value nbk_ml_binary_list(value ctx, value m, value q, value p, value eps2, value thr,
                                  value pi, value pj, value e) {
  CAMLparam5(ctx, m, q, p, eps2);
  CAMLxparam4(thr, pi, pj, e);
  int npairs = 0;
  check(Ctx_val(ctx), nbk_binary_list(Ctx_val(ctx), DIM0(m), D(m), D(q), D(p), Double_val(eps2),
                                      Double_val(thr), DIM0(pi), I32(pi), I32(pj), D(e), &npairs));
  CAMLreturn(Val_int(npairs));
}
CAMLprim value nbk_ml_binary_list_bc(value *a, int n) {
  (void)n;
  return nbk_ml_binary_list(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8]);
}

/* ---- resident packed bodies (small-active-set callers) ------------------------------------------ */
/* external bodies_upload : ctx -> m -> q -> vel -> is_velocity:bool -> unit */
CAMLprim value nbk_ml_bodies_upload(value ctx, value m, value q, value vel, value isv) {
  CAMLparam5(ctx, m, q, vel, isv);
  check(Ctx_val(ctx), nbk_bodies_upload(Ctx_val(ctx), DIM0(m), D(m), D(q), D(vel), Bool_val(isv)));
  CAMLreturn(Val_unit);
}
/* external bodies_update : ctx -> i0:int -> i1:int -> q -> vel -> is_velocity:bool -> unit */
CAMLprim value nbk_ml_bodies_update(value ctx, value i0, value i1, value q, value vel, value isv) {
  CAMLparam5(ctx, i0, i1, q, vel);
  CAMLxparam1(isv);
  check(Ctx_val(ctx), nbk_bodies_update(Ctx_val(ctx), Int_val(i0), Int_val(i1), D(q), D(vel),
                                        Bool_val(isv)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_bodies_update_bc(value *a, int n) {
  (void)n;
  return nbk_ml_bodies_update(a[0], a[1], a[2], a[3], a[4], a[5]);
}
/* external resident_count : ctx -> int */
CAMLprim value nbk_ml_resident_count(value ctx) { return Val_int(nbk_resident_count(Ctx_val(ctx))); }
/* external resident_grad_v_sum : ctx -> eps2:float -> range -> gsum -> unit */
CAMLprim value nbk_ml_resident_grad_v_sum(value ctx, value eps2, value rng, value g) {
  CAMLparam4(ctx, eps2, rng, g);
  check(Ctx_val(ctx), nbk_resident_grad_v_sum(Ctx_val(ctx), Double_val(eps2), R0(rng), R1(rng),
                                              R2(rng), R3(rng), D(g)));
  CAMLreturn(Val_unit);
}
/* external resident_grad_v_dgrad_v_sum : ctx -> eps2:float -> range -> gsum -> dgsum -> unit */
CAMLprim value nbk_ml_resident_grad_v_dgrad_v_sum(value ctx, value eps2, value rng, value g,
                                                  value dg) {
  CAMLparam5(ctx, eps2, rng, g, dg);
  check(Ctx_val(ctx), nbk_resident_grad_v_dgrad_v_sum(Ctx_val(ctx), Double_val(eps2), R0(rng),
                                                      R1(rng), R2(rng), R3(rng), D(g), D(dg)));
  CAMLreturn(Val_unit);
}
/* external resident_v_sum : ctx -> eps2:float -> range -> phi -> float */
CAMLprim value nbk_ml_resident_v_sum(value ctx, value eps2, value rng, value phi) {
  CAMLparam4(ctx, eps2, rng, phi);
  double total = 0.0;
  check(Ctx_val(ctx), nbk_resident_v_sum(Ctx_val(ctx), Double_val(eps2), R0(rng), R1(rng), R2(rng),
                                         R3(rng), D(phi), &total));
  CAMLreturn(caml_copy_double(total));
}
/* external resident_kick_sum : ctx -> eta:float -> dt:float -> range -> dv -> tmin -> unit */
CAMLprim value nbk_ml_resident_kick_sum(value ctx, value eta, value dt, value rng, value dv,
                                        value tmin) {
  CAMLparam5(ctx, eta, dt, rng, dv);
  CAMLxparam1(tmin);
  check(Ctx_val(ctx), nbk_resident_kick_sum(Ctx_val(ctx), Double_val(eta), Double_val(dt), R0(rng),
                                            R1(rng), R2(rng), R3(rng), D(dv), D(tmin)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_resident_kick_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_resident_kick_sum(a[0], a[1], a[2], a[3], a[4], a[5]);
}
/* external resident_leapfrog_steps : ctx -> dt:float -> nsteps:int -> q -> v -> unit
 * the steady state of Leapfrog.advance_const_dt entirely on the device (leapfrog.ml:160-165) */
CAMLprim value nbk_ml_resident_leapfrog_steps(value ctx, value dt, value nsteps, value q, value v) {
  CAMLparam5(ctx, dt, nsteps, q, v);
  check(Ctx_val(ctx), nbk_resident_leapfrog_steps(Ctx_val(ctx), Double_val(dt), Int_val(nsteps),
                                                  D(q), D(v)));
  CAMLreturn(Val_unit);
}

/* ---- Leapfrog per-phase calls on the resident array (above nbk_lf_subsystem_max bodies) -------- */
/* external lf_permute : ctx -> perm:(int32) -> unit           sort_sub, leapfrog.ml:106-111 */
CAMLprim value nbk_ml_lf_permute(value ctx, value perm) {
  CAMLparam2(ctx, perm);
  check(Ctx_val(ctx), nbk_lf_permute(Ctx_val(ctx), DIM0(perm), I32(perm)));
  CAMLreturn(Val_unit);
}
/* external lf_sort_sub : ctx -> perm:(int32) -> unit   sort_sub bs iend on the device, iend = dim perm;
 * perm receives new row s = old row perm.{s} */
CAMLprim value nbk_ml_lf_sort_sub(value ctx, value perm) {
  CAMLparam2(ctx, perm);
  check(Ctx_val(ctx), nbk_lf_sort_sub(Ctx_val(ctx), DIM0(perm), I32(perm)));
  CAMLreturn(Val_unit);
}
/* external lf_begin : ctx -> ibegin:int -> iend:int -> unit   leapfrog.ml:125-127 */
CAMLprim value nbk_ml_lf_begin(value ctx, value ibegin, value iend) {
  CAMLparam3(ctx, ibegin, iend);
  check(Ctx_val(ctx), nbk_lf_begin(Ctx_val(ctx), Int_val(ibegin), Int_val(iend)));
  CAMLreturn(Val_unit);
}
/* external lf_drift : ctx -> i0:int -> i1:int -> dt:float -> unit   leapfrog.ml:64-68 */
CAMLprim value nbk_ml_lf_drift(value ctx, value i0, value i1, value dt) {
  CAMLparam4(ctx, i0, i1, dt);
  check(Ctx_val(ctx), nbk_lf_drift(Ctx_val(ctx), Int_val(i0), Int_val(i1), Double_val(dt)));
  CAMLreturn(Val_unit);
}
/* external lf_kick_block : ctx -> iend:int -> ibegin:int -> eta:float -> dt:float ->
 *                          with_tscale:bool -> drift_after:float -> dt_max -> unit
 * the kick loop of advance_subsystem applied in place + the drift that follows (leapfrog.ml:132-137);
 * dt_max receives the updated dt_max.{0 .. iend-1} */
CAMLprim value nbk_ml_lf_kick_block(value ctx, value iend, value ibegin, value eta, value dt,
                                    value with_ts, value drift_after, value dtm) {
  CAMLparam5(ctx, iend, ibegin, eta, dt);
  CAMLxparam3(with_ts, drift_after, dtm);
  check(Ctx_val(ctx), nbk_lf_kick_block(Ctx_val(ctx), Int_val(iend), Int_val(ibegin), Double_val(eta),
                                        Double_val(dt), Bool_val(with_ts), 1,
                                        Double_val(drift_after), D(dtm)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_lf_kick_block_bc(value *a, int n) {
  (void)n;
  return nbk_ml_lf_kick_block(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
}
/* external lf_subsystem_max : unit -> int */
CAMLprim value nbk_ml_lf_subsystem_max(value unit) {
  (void)unit;
  return Val_int(nbk_lf_subsystem_max());
}

/* ---- Hermite on device-resident state (hermite.ml:95-130, 171-178) ----------------------------- */
/* external hermite_upload : ctx -> t -> m -> q -> p -> h -> f -> df -> unit */
CAMLprim value nbk_ml_hermite_upload(value ctx, value t, value m, value q, value p, value h,
                                     value f, value df) {
  CAMLparam5(ctx, t, m, q, p);
  CAMLxparam3(h, f, df);
  check(Ctx_val(ctx), nbk_hermite_upload(Ctx_val(ctx), DIM0(m), D(t), D(m), D(q), D(p), D(h), D(f),
                                         D(df)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_hermite_upload_bc(value *a, int n) {
  (void)n;
  return nbk_ml_hermite_upload(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
}
/* external hermite_resident_max : unit -> int */
CAMLprim value nbk_ml_hermite_resident_max(value unit) {
  (void)unit;
  return Val_int(nbk_hermite_resident_max());
}
/* external hermite_advance_resident : ctx -> stop_time:float -> eta:float -> eps2:float ->
 *                                     max_steps:int -> int      (advance_body calls made)
 * Hermite.advance_bodies + finish_bs in ONE persistent kernel (hermite.ml:132-143, 171-178) */
CAMLprim value nbk_ml_hermite_advance_resident(value ctx, value stop, value eta, value eps2,
                                               value max_steps) {
  CAMLparam5(ctx, stop, eta, eps2, max_steps);
  long ns = 0;
  check(Ctx_val(ctx), nbk_hermite_advance_resident(Ctx_val(ctx), Double_val(stop), Double_val(eta),
                                                   Double_val(eps2), Long_val(max_steps), &ns));
  CAMLreturn(Val_long(ns));
}
/* external hermite_download : ctx -> t -> q -> p -> h -> f -> df -> unit */
CAMLprim value nbk_ml_hermite_download(value ctx, value t, value q, value p, value h, value f,
                                       value df) {
  CAMLparam5(ctx, t, q, p, h);
  CAMLxparam2(f, df);
  check(Ctx_val(ctx), nbk_hermite_download(Ctx_val(ctx), D(t), D(q), D(p), D(h), D(f), D(df)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_hermite_download_bc(value *a, int n) {
  (void)n;
  return nbk_ml_hermite_download(a[0], a[1], a[2], a[3], a[4], a[5], a[6]);
}
/* external hermite_body_sum : ctx -> i:int -> nt:float -> eps2:float -> upd:int -> upd_state ->
 *                             gsum -> dgsum -> unit
 * one advance_body force evaluation on the resident state; upd >= 0 first overwrites that body with
 * upd_state (13 doubles: t, q, p, f, df) - the previous step's result rides along */
CAMLprim value nbk_ml_hermite_body_sum(value ctx, value i, value nt, value eps2, value upd,
                                       value upd_state, value g, value dg) {
  CAMLparam5(ctx, i, nt, eps2, upd);
  CAMLxparam3(upd_state, g, dg);
  check(Ctx_val(ctx), nbk_hermite_body_sum(Ctx_val(ctx), Int_val(i), Double_val(nt), Double_val(eps2),
                                           Int_val(upd), D(upd_state), D(g), D(dg)));
  CAMLreturn(Val_unit);
}
CAMLprim value nbk_ml_hermite_body_sum_bc(value *a, int n) {
  (void)n;
  return nbk_ml_hermite_body_sum(a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
}

/* external adv_energy : ctx -> eps2:float -> float * float    Energy.energy of the resident records */
CAMLprim value nbk_ml_adv_energy(value ctx, value eps2) {
  CAMLparam2(ctx, eps2);
  double ke = 0.0, pe = 0.0;
  check(Ctx_val(ctx), nbk_adv_energy(Ctx_val(ctx), Double_val(eps2), &ke, &pe));
  CAMLreturn(pair_of_doubles(ke, pe));
}
/* external set_peer_timeout : ctx -> seconds:float -> unit */
CAMLprim value nbk_ml_set_peer_timeout(value ctx, value s) {
  CAMLparam2(ctx, s);
  check(Ctx_val(ctx), nbk_set_peer_timeout(Ctx_val(ctx), Double_val(s)));
  CAMLreturn(Val_unit);
}

CAMLprim value nbk_ml_set_sym(value ctx, value mode) {
  CAMLparam2(ctx, mode);
  check(Ctx_val(ctx), nbk_set_sym(Ctx_val(ctx), Int_val(mode)));
  CAMLreturn(Val_unit);
}
